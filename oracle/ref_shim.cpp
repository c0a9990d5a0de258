// C-callable shim over the UNMODIFIED reference library.  TEST INFRASTRUCTURE ONLY.
//
// oracle/Makefile compiles /root/reference/lib/{edm,grid,gaussian_grid,edm_bias}.cpp where
// they lie, plus this file, into oracle/_ref/libedm_ref.so (git-ignored).  Nothing in the
// product (electronic-dance-music_b200/) links or loads it; only tests/, bench.py's
// cpu_baseline / --impl reference arm and __graft_entry__.smoke() do, as the checker or the
// timed CPU baseline.  No reference source is copied: this file only *calls* the reference's
// public classes (lib/grid.h:185, lib/gaussian_grid.h:59, lib/edm_bias.h:29).
//
// EDMBias keeps its limiter backlog in private members; the parity tests must see them, so
// the class keyword is widened for this translation unit only (an ABI-neutral trick: member
// layout is unchanged because access specifiers do not reorder these members in g++).
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <new>
#include <string>
#include <vector>
#include <chrono>

#include <sstream>
#include <fstream>
#include <iostream>
#include <iterator>
#include <map>
#include <iomanip>
#include <cstdio>
#define private public
#include "edm_bias.h"
#undef private
#include "gaussian_grid.h"
#include "grid.h"

using namespace EDM;

namespace {

struct GaussHandle {
  int dim;
  GaussGrid* g;
  bool owned;
};
struct GridHandle {
  int dim;
  Grid* g;
};
struct BiasHandle {
  void* storage;  // calloc'd: the reference never initialises overflow_buffer_
  EDMBias* b;
  GaussHandle gauss;
  GridHandle hist;
};

template <int D> DimmedGaussGrid<D>* gg(GaussHandle* h) { return static_cast<DimmedGaussGrid<D>*>(h->g); }
template <int D> DimmedGrid<D>* pg(GridHandle* h) { return static_cast<DimmedGrid<D>*>(h->g); }

template <int D> void gauss_info_t(GaussHandle* h, int* n, double* dx, double* mn, double* mx, int* mini) {
  DimmedGaussGrid<D>* g = gg<D>(h);
  for (int i = 0; i < D; i++) {
    n[i] = g->grid_.grid_number_[i];
    dx[i] = g->grid_.dx_[i];
    mn[i] = g->grid_.min_[i];
    mx[i] = g->grid_.max_[i];
    mini[i] = (int)g->minisize_[i];
  }
}
template <int D> void gauss_arrays_t(GaussHandle* h, double* v, double* d, int set) {
  DimmedGaussGrid<D>* g = gg<D>(h);
  size_t sz = g->grid_.grid_size_;
  if (set) {
    memcpy(g->grid_.grid_, v, sz * sizeof(double));
    memcpy(g->grid_.grid_deriv_, d, sz * D * sizeof(double));
  } else {
    memcpy(v, g->grid_.grid_, sz * sizeof(double));
    memcpy(d, g->grid_.grid_deriv_, sz * D * sizeof(double));
  }
}
template <int D> void gauss_tables_t(GaussHandle* h, int dimi, double* denom, double* deriv) {
  DimmedGaussGrid<D>* g = gg<D>(h);
  memcpy(denom, g->bc_denom_table_[dimi], sizeof(double) * BC_TABLE_SIZE);
  memcpy(deriv, g->bc_denom_deriv_table_[dimi], sizeof(double) * BC_TABLE_SIZE);
}
template <int D> void grid_info_t(GridHandle* h, int* n, double* dx, double* mn, double* mx, int* flags) {
  DimmedGrid<D>* g = pg<D>(h);
  for (int i = 0; i < D; i++) {
    n[i] = g->grid_number_[i];
    dx[i] = g->dx_[i];
    mn[i] = g->min_[i];
    mx[i] = g->max_[i];
    flags[2 + i] = g->b_periodic_[i];
  }
  flags[0] = g->b_derivatives_;
  flags[1] = g->b_interpolate_;
}
template <int D> void grid_arrays_t(GridHandle* h, double* v, double* d, int set) {
  DimmedGrid<D>* g = pg<D>(h);
  size_t sz = g->grid_size_;
  if (set) {
    memcpy(g->grid_, v, sz * sizeof(double));
    if (g->b_derivatives_ && d) memcpy(g->grid_deriv_, d, sz * D * sizeof(double));
  } else {
    memcpy(v, g->grid_, sz * sizeof(double));
    if (g->b_derivatives_ && d) memcpy(d, g->grid_deriv_, sz * D * sizeof(double));
  }
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------ GaussGrid
void* ref_gauss_create(int dim, const double* mn, const double* mx, const double* spacing,
                       const int* periodic, int interp, const double* sigma) {
  GaussHandle* h = new GaussHandle;
  h->dim = dim;
  h->g = make_gauss_grid(dim, mn, mx, spacing, periodic, interp, sigma);
  h->owned = true;
  return h;
}
void ref_gauss_destroy(void* p) {
  GaussHandle* h = (GaussHandle*)p;
  if (h->owned) {
    delete h->g;
    delete h;
  }
}
void ref_gauss_set_boundary(void* p, const double* mn, const double* mx, const int* periodic) {
  ((GaussHandle*)p)->g->set_boundary(mn, mx, periodic);
}
void ref_gauss_info(void* p, int* n, double* dx, double* mn, double* mx, int* mini) {
  GaussHandle* h = (GaussHandle*)p;
  if (h->dim == 1) gauss_info_t<1>(h, n, dx, mn, mx, mini);
  if (h->dim == 2) gauss_info_t<2>(h, n, dx, mn, mx, mini);
  if (h->dim == 3) gauss_info_t<3>(h, n, dx, mn, mx, mini);
}
size_t ref_gauss_size(void* p) { return ((GaussHandle*)p)->g->get_grid_size(); }
void ref_gauss_get_arrays(void* p, double* v, double* d) {
  GaussHandle* h = (GaussHandle*)p;
  if (h->dim == 1) gauss_arrays_t<1>(h, v, d, 0);
  if (h->dim == 2) gauss_arrays_t<2>(h, v, d, 0);
  if (h->dim == 3) gauss_arrays_t<3>(h, v, d, 0);
}
void ref_gauss_set_arrays(void* p, double* v, double* d) {
  GaussHandle* h = (GaussHandle*)p;
  if (h->dim == 1) gauss_arrays_t<1>(h, v, d, 1);
  if (h->dim == 2) gauss_arrays_t<2>(h, v, d, 1);
  if (h->dim == 3) gauss_arrays_t<3>(h, v, d, 1);
}
void ref_gauss_tables(void* p, int dimi, double* denom, double* deriv) {
  GaussHandle* h = (GaussHandle*)p;
  if (h->dim == 1) gauss_tables_t<1>(h, dimi, denom, deriv);
  if (h->dim == 2) gauss_tables_t<2>(h, dimi, denom, deriv);
  if (h->dim == 3) gauss_tables_t<3>(h, dimi, denom, deriv);
}
double ref_gauss_add_value(void* p, const double* x, double height) {
  return ((GaussHandle*)p)->g->add_value(x, height);
}
void ref_gauss_add_values(void* p, long n, const double* x, const double* height, double* bias_added) {
  GaussHandle* h = (GaussHandle*)p;
  for (long i = 0; i < n; i++) {
    double ba = h->g->add_value(x + i * h->dim, height[i]);
    if (bias_added) bias_added[i] = ba;
  }
}
void ref_gauss_eval(void* p, long n, const double* x, double* val, double* der) {
  GaussHandle* h = (GaussHandle*)p;
  for (long i = 0; i < n; i++) val[i] = h->g->get_value_deriv(x + i * h->dim, der + i * h->dim);
}
void ref_gauss_get_value(void* p, long n, const double* x, double* val) {
  GaussHandle* h = (GaussHandle*)p;
  for (long i = 0; i < n; i++) val[i] = h->g->get_value(x + i * h->dim);
}
void ref_gauss_remap(void* p, double* x) {
  GaussHandle* h = (GaussHandle*)p;
  if (h->dim == 1) gg<1>(h)->remap(x);
  if (h->dim == 2) gg<2>(h)->remap(x);
  if (h->dim == 3) gg<3>(h)->remap(x);
}
void ref_gauss_set_interpolation(void* p, int b) { ((GaussHandle*)p)->g->set_interpolation(b); }
void ref_gauss_write(void* p, const char* fn) { ((GaussHandle*)p)->g->write(fn); }

// ------------------------------------------------------------------ plain Grid
void* ref_grid_create(int dim, const double* mn, const double* mx, const double* spacing,
                      const int* periodic, int b_deriv, int b_interp) {
  GridHandle* h = new GridHandle;
  h->dim = dim;
  h->g = make_grid(dim, mn, mx, spacing, periodic, b_deriv, b_interp);
  return h;
}
void* ref_grid_read(int dim, const char* filename, int b_interp) {
  GridHandle* h = new GridHandle;
  h->dim = dim;
  h->g = read_grid(dim, filename, b_interp);
  return h;
}
void ref_grid_destroy(void* p) {
  GridHandle* h = (GridHandle*)p;
  delete h->g;
  delete h;
}
void ref_grid_info(void* p, int* n, double* dx, double* mn, double* mx, int* flags) {
  GridHandle* h = (GridHandle*)p;
  if (h->dim == 1) grid_info_t<1>(h, n, dx, mn, mx, flags);
  if (h->dim == 2) grid_info_t<2>(h, n, dx, mn, mx, flags);
  if (h->dim == 3) grid_info_t<3>(h, n, dx, mn, mx, flags);
}
size_t ref_grid_size(void* p) { return ((GridHandle*)p)->g->get_grid_size(); }
void ref_grid_get_arrays(void* p, double* v, double* d) {
  GridHandle* h = (GridHandle*)p;
  if (h->dim == 1) grid_arrays_t<1>(h, v, d, 0);
  if (h->dim == 2) grid_arrays_t<2>(h, v, d, 0);
  if (h->dim == 3) grid_arrays_t<3>(h, v, d, 0);
}
void ref_grid_set_arrays(void* p, double* v, double* d) {
  GridHandle* h = (GridHandle*)p;
  if (h->dim == 1) grid_arrays_t<1>(h, v, d, 1);
  if (h->dim == 2) grid_arrays_t<2>(h, v, d, 1);
  if (h->dim == 3) grid_arrays_t<3>(h, v, d, 1);
}
void ref_grid_set_interpolation(void* p, int b) { ((GridHandle*)p)->g->set_interpolation(b); }
void ref_grid_eval(void* p, long n, const double* x, double* val, double* der) {
  GridHandle* h = (GridHandle*)p;
  for (long i = 0; i < n; i++) val[i] = h->g->get_value_deriv(x + i * h->dim, der + i * h->dim);
}
void ref_grid_get_value(void* p, long n, const double* x, double* val) {
  GridHandle* h = (GridHandle*)p;
  for (long i = 0; i < n; i++) val[i] = h->g->get_value(x + i * h->dim);
}
void ref_grid_hist_add(void* p, long n, const double* x, const double* v) {
  GridHandle* h = (GridHandle*)p;
  for (long i = 0; i < n; i++) h->g->add_value(x + i * h->dim, v[i]);
}
void ref_grid_write(void* p, const char* fn) { ((GridHandle*)p)->g->write(fn); }
double ref_grid_expected_bias(void* p) { return ((GridHandle*)p)->g->expected_bias(); }

// ------------------------------------------------------------------ EDMBias
void* ref_bias_create(const char* edm_file) {
  BiasHandle* h = new BiasHandle;
  h->storage = calloc(1, sizeof(EDMBias) + 64);
  h->b = new (h->storage) EDMBias(std::string(edm_file));
  h->gauss.g = NULL;
  h->hist.g = NULL;
  return h;
}
void ref_bias_destroy(void* p) {
  BiasHandle* h = (BiasHandle*)p;
  h->b->~EDMBias();
  free(h->storage);
  delete h;
}
void ref_bias_setup(void* p, double temperature, double boltz) { ((BiasHandle*)p)->b->setup(temperature, boltz); }
void ref_bias_subdivide(void* p, const double* sublo, const double* subhi, const double* boxlo,
                        const double* boxhi, const int* periodic, const double* skin) {
  ((BiasHandle*)p)->b->subdivide(sublo, subhi, boxlo, boxhi, periodic, skin);
}
void* ref_bias_gauss(void* p) {
  BiasHandle* h = (BiasHandle*)p;
  h->gauss.dim = (int)h->b->dim_;
  h->gauss.g = h->b->bias_;
  h->gauss.owned = false;
  return &h->gauss;
}
void* ref_bias_hist(void* p) {
  BiasHandle* h = (BiasHandle*)p;
  h->hist.dim = (int)h->b->dim_;
  h->hist.g = h->b->cv_hist_;
  return &h->hist;
}
// params[]: 0 dim, 1 b_tempering, 2 b_targeting, 3 global_tempering, 4 bias_factor,
// 5 boltzmann_factor, 6 hill_prefactor, 7 bias_per_step, 8 hill_density, 9 cum_bias,
// 10 total_volume, 11 expected_target, 12 b_outofbounds, 13 steps
void ref_bias_params(void* p, double* out) {
  EDMBias* b = ((BiasHandle*)p)->b;
  out[0] = b->dim_;
  out[1] = b->b_tempering_;
  out[2] = b->b_targeting_;
  out[3] = b->global_tempering_;
  out[4] = b->bias_factor_;
  out[5] = b->boltzmann_factor_;
  out[6] = b->hill_prefactor_;
  out[7] = b->bias_per_step_;
  out[8] = b->hill_density_;
  out[9] = b->cum_bias_;
  out[10] = b->total_volume_;
  out[11] = b->expected_target_;
  out[12] = b->b_outofbounds_;
  out[13] = (double)b->steps_;
}
void ref_bias_arrays(void* p, double* dx, double* sigma, double* mn, double* mx) {
  EDMBias* b = ((BiasHandle*)p)->b;
  for (unsigned i = 0; i < b->dim_; i++) {
    dx[i] = b->bias_dx_[i];
    sigma[i] = b->bias_sigma_[i];
    mn[i] = b->min_[i];
    mx[i] = b->max_[i];
  }
}
void ref_bias_set_cum_bias(void* p, double v) { ((BiasHandle*)p)->b->cum_bias_ = v; }
// overflow deque: [left, right) slots of (dim+1) doubles, lib/edm_bias.h:175-177
void ref_bias_backlog(void* p, long* left, long* right, double* buffer) {
  EDMBias* b = ((BiasHandle*)p)->b;
  *left = (long)b->overflow_left_i_;
  *right = (long)b->overflow_right_i_;
  if (buffer) memcpy(buffer, b->overflow_buffer_, sizeof(double) * BIAS_BUFFER_DBLS);
}
void ref_bias_set_mask(void* p, const int* mask) { ((BiasHandle*)p)->b->set_mask(mask); }
void ref_bias_flush_log(void* p) { ((BiasHandle*)p)->b->hill_output_.flush(); }

double ref_bias_update_forces(void* p, long n, const double* x, long xstride, double* f, long fstride,
                              int apply_mask) {
  EDMBias* b = ((BiasHandle*)p)->b;
  std::vector<const double*> xr(n);
  std::vector<double*> fr(n);
  for (long i = 0; i < n; i++) {
    xr[i] = x + i * xstride;
    fr[i] = f + i * fstride;
  }
  return b->update_forces((int)n, xr.data(), fr.data(), apply_mask);
}
void ref_bias_add_hills(void* p, long n, const double* x, long xstride, const double* runiform, int apply_mask) {
  EDMBias* b = ((BiasHandle*)p)->b;
  std::vector<const double*> xr(n);
  for (long i = 0; i < n; i++) xr[i] = x + i * xstride;
  b->add_hills((int)n, xr.data(), runiform, apply_mask);
}
void ref_bias_pre_add_hill(void* p, int est) { ((BiasHandle*)p)->b->pre_add_hill(est); }
void ref_bias_add_hill_many(void* p, long n, const double* x, const double* runiform) {
  EDMBias* b = ((BiasHandle*)p)->b;
  for (long i = 0; i < n; i++) b->add_hill(x + i * b->dim_, runiform[i]);
}
void ref_bias_post_add_hill(void* p) { ((BiasHandle*)p)->b->post_add_hill(); }
void ref_bias_write_bias(void* p, const char* fn) { ((BiasHandle*)p)->b->write_bias(fn); }
void ref_bias_write_histogram(void* p) { ((BiasHandle*)p)->b->write_histogram(); }
void ref_bias_clear_histogram(void* p) { ((BiasHandle*)p)->b->clear_histogram(); }

// Restatement of the pair loop of lammps/fix_edm_pair.cpp:177-240 as a driver over the
// reference's own update_force / add_hill, in the LIB-LEVEL order (evaluate every pair, then
// pre/add/post) that SURVEY 3.2 fixes as the parity definition.  pairs: (i, j) half list, both
// atoms local.  shift: optional per-pair periodic image shift of x[j] (3 doubles) or NULL.
// uniforms: 2 per pair (fix_edm_pair.cpp:230-236) or NULL when do_hills == 0.
double ref_pair_step(void* p, long npairs, const int* pi, const int* pj, const double* x, double* f,
                     const double* shift, int do_hills, int est, const double* uniforms, double* r_out) {
  EDMBias* b = ((BiasHandle*)p)->b;
  double energy = 0;
  std::vector<double> rr(npairs);
  for (long k = 0; k < npairs; k++) {
    int i = pi[k], j = pj[k];
    double delx = x[3 * i + 0] - x[3 * j + 0];
    double dely = x[3 * i + 1] - x[3 * j + 1];
    double delz = x[3 * i + 2] - x[3 * j + 2];
    if (shift) {
      delx -= shift[3 * k + 0];
      dely -= shift[3 * k + 1];
      delz -= shift[3 * k + 2];
    }
    double r = sqrt(delx * delx + dely * dely + delz * delz);
    double rinv = 1.0 / r;
    delx *= rinv;
    dely *= rinv;
    delz *= rinv;
    double edm_force[1] = {0};
    energy += b->update_force(&r, edm_force);
    f[3 * i + 0] += delx * edm_force[0];
    f[3 * i + 1] += dely * edm_force[0];
    f[3 * i + 2] += delz * edm_force[0];
    f[3 * j + 0] -= delx * edm_force[0];
    f[3 * j + 1] -= dely * edm_force[0];
    f[3 * j + 2] -= delz * edm_force[0];
    rr[k] = r;
    if (r_out) r_out[k] = r;
  }
  if (do_hills) {
    b->pre_add_hill(est);
    for (long k = 0; k < npairs; k++) {
      b->add_hill(&rr[k], uniforms[2 * k]);
      b->add_hill(&rr[k], uniforms[2 * k + 1]);
    }
    b->post_add_hill();
  }
  return energy;
}

// The same with ghost atoms (lammps/fix_edm_pair.cpp:223, 233): i is local, j may be a ghost (>= nlocal): no force on
// it, one proposal instead of two.
double ref_pair_step_ghost(void* p, long npairs, const int* pi, const int* pj, const double* x, double* f,
                           const double* shift, long nlocal, int do_hills, int est, const double* uniforms,
                           double* r_out, long* ncalls_out) {
  EDMBias* b = ((BiasHandle*)p)->b;
  double energy = 0;
  long ncalls = 0;
  std::vector<double> rr(npairs);
  for (long k = 0; k < npairs; k++) {
    int i = pi[k], j = pj[k];
    double delx = x[3 * i + 0] - x[3 * j + 0];
    double dely = x[3 * i + 1] - x[3 * j + 1];
    double delz = x[3 * i + 2] - x[3 * j + 2];
    if (shift) {
      delx -= shift[3 * k + 0];
      dely -= shift[3 * k + 1];
      delz -= shift[3 * k + 2];
    }
    double r = sqrt(delx * delx + dely * dely + delz * delz);
    double rinv = 1.0 / r;
    delx *= rinv;
    dely *= rinv;
    delz *= rinv;
    double edm_force[1] = {0};
    energy += b->update_force(&r, edm_force);
    f[3 * i + 0] += delx * edm_force[0];
    f[3 * i + 1] += dely * edm_force[0];
    f[3 * i + 2] += delz * edm_force[0];
    if (j < nlocal) {
      f[3 * j + 0] -= delx * edm_force[0];
      f[3 * j + 1] -= dely * edm_force[0];
      f[3 * j + 2] -= delz * edm_force[0];
    }
    rr[k] = r;
    if (r_out) r_out[k] = r;
  }
  if (do_hills) {
    b->pre_add_hill(est);
    for (long k = 0; k < npairs; k++) {
      b->add_hill(&rr[k], uniforms[2 * k]);
      ncalls++;
      if (pj[k] < nlocal) {
        b->add_hill(&rr[k], uniforms[2 * k + 1]);
        ncalls++;
      }
    }
    b->post_add_hill();
  }
  if (ncalls_out) *ncalls_out = ncalls;
  return energy;
}

// ------------------------------------------------------------------ timing helpers (CPU baseline)
// Timed inside C++ so that ctypes/array marshalling stays outside the measured region.
double ref_time_pair_eval(void* p, long npairs, const double* r, int repeats) {
  EDMBias* b = ((BiasHandle*)p)->b;
  double sink = 0;
  auto t0 = std::chrono::steady_clock::now();
  for (int it = 0; it < repeats; it++) {
    for (long k = 0; k < npairs; k++) {
      double fr[1] = {0};
      sink += b->update_force(&r[k], fr) + fr[0];
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  if (sink == 1.2345e-300) std::cerr << "";
  return std::chrono::duration<double>(t1 - t0).count();
}
// FixEDMPair::post_force's pair loop, lammps/fix_edm_pair.cpp:173-247, as a driver over the UNMODIFIED reference
// library in the reference's own order: per pair r = sqrt(del^2), update_force(&r), +-scatter into f[i], f[j] (both
// local), and on a hill step two add_hill calls, between pre_add_hill(last_calls) and post_add_hill.  The uniforms
// come from a counter hash in place of LAMMPS' RanMars (one draw per add_hill, as there).  Returns seconds.
static inline unsigned long long shim_mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
double ref_time_fix_pair(void* p, long npairs, const int* pi, const int* pj, const signed char* img,
                         const double* box, const double* x, double* f, int do_hills, int est,
                         unsigned long long seed, unsigned long long step, double* energy_out, long* ncalls_out) {
  EDMBias* b = ((BiasHandle*)p)->b;
  const unsigned long long key = shim_mix64(seed ^ shim_mix64(step + 0x9E3779B97F4A7C15ULL));
  unsigned long long ctr = 0;
  double edm_energy = 0;
  long ncalls = 0;
  auto t0 = std::chrono::steady_clock::now();
  if (do_hills) b->pre_add_hill(est);
  for (long k = 0; k < npairs; k++) {
    const int i = pi[k], j = pj[k];
    double delx = x[3 * i + 0] - x[3 * j + 0] - img[3 * k + 0] * box[0];
    double dely = x[3 * i + 1] - x[3 * j + 1] - img[3 * k + 1] * box[1];
    double delz = x[3 * i + 2] - x[3 * j + 2] - img[3 * k + 2] * box[2];
    double r = sqrt(delx * delx + dely * dely + delz * delz);
    double rinv = 1.0 / r;
    delx *= rinv;
    dely *= rinv;
    delz *= rinv;
    double edm_force[1] = {0};
    edm_energy += b->update_force(&r, edm_force);
    f[3 * i + 0] += delx * edm_force[0];
    f[3 * i + 1] += dely * edm_force[0];
    f[3 * i + 2] += delz * edm_force[0];
    f[3 * j + 0] -= delx * edm_force[0];
    f[3 * j + 1] -= dely * edm_force[0];
    f[3 * j + 2] -= delz * edm_force[0];
    if (do_hills) {
      for (int w = 0; w < 2; w++) {
        const unsigned long long bits = shim_mix64(key + (ctr++) * 0x9E3779B97F4A7C15ULL);
        b->add_hill(&r, (double)(bits >> 11) * (1.0 / 9007199254740992.0));
        ncalls++;
      }
    }
  }
  if (do_hills) b->post_add_hill();
  auto t1 = std::chrono::steady_clock::now();
  if (energy_out) *energy_out = edm_energy;
  if (ncalls_out) *ncalls_out = ncalls;
  return std::chrono::duration<double>(t1 - t0).count();
}
double ref_time_add_values(void* p, long n, const double* x, const double* h) {
  GaussHandle* g = (GaussHandle*)p;
  double sink = 0;
  auto t0 = std::chrono::steady_clock::now();
  for (long i = 0; i < n; i++) sink += g->g->add_value(x + i * g->dim, h[i]);
  auto t1 = std::chrono::steady_clock::now();
  if (sink == 1.2345e-300) std::cerr << "";
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
