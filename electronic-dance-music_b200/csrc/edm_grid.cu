// Grid / GaussGrid in HBM: geometry, McGDP tables, upload/download, batched evaluation
// (Grid::get_value_deriv) and batched hill deposition (GaussGrid::add_value).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <atomic>
#include <vector>

#include "edm_host.h"

namespace edm {

static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void set_error(const std::string& msg) { g_last_error = msg; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
  g_last_error = buf;
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return EDM_ERR_NO_DEVICE;
  return EDM_ERR_CUDA;
}
int ensure_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    (void)cudaGetLastError();
    set_error("no CUDA device: this library has no CPU fallback");
    return EDM_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= count) {
    set_error("device index out of range");
    return EDM_ERR_ARG;
  }
  EDM_CUDA(cudaSetDevice(device));
  return EDM_OK;
}
int sm_count(int device) {
  static int cached[64];
  if (device < 0 || device >= 64) return 148;
  if (cached[device] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
    cached[device] = n;
  }
  return cached[device];
}
int Scratch::reserve(size_t need) {
  if (need <= bytes) return EDM_OK;
  if (p) EDM_CUDA(cudaFree(p));
  p = nullptr;
  bytes = 0;
  size_t want = need + need / 4 + 256;
  EDM_CUDA(cudaMalloc(&p, want));
  bytes = want;
  return EDM_OK;
}
void Scratch::release() {
  if (p) cudaFree(p);
  p = nullptr;
  bytes = 0;
}

// ------------------------------------------------------------------ host-side geometry

static int host_int_floor(double v) { return (int)((int)v < 0.0 ? -ceil(fabs(v)) : floor(v)); }  // lib/grid.h:17-20

static double host_sigmoid(double x) {  // lib/gaussian_grid.h:16-23
  if (x < 0) return 1;
  if (x > 1) return 0;
  return 2 * x * x * x - 3 * x * x + 1;
}
static double host_sigmoid_dx(double x) {  // lib/gaussian_grid.h:25-32
  if (x < 0) return 0;
  if (x > 1) return 0;
  return 6 * x * x - 6 * x;
}

// DimmedGrid::get_index for one dim, lib/grid.h:264-273
static long long host_get_index(const GridDesc& d, int i, double x) {
  double xi = x;
  if (d.periodic[i]) xi -= (d.max[i] - d.min[i]) * host_int_floor((xi - d.min[i]) / (d.max[i] - d.min[i]));
  return (long long)(size_t)floor((xi - d.min[i]) / d.dx[i]);
}

static void finish_geometry(edm_grid* g) {
  GridDesc& d = g->d;
  d.rec_w = d.dim == 1 ? 2 : 4;
  d.size = 1;
  d.vol_element = 1;
  for (int i = 0; i < d.dim; i++) {
    d.len[i] = d.max[i] - d.min[i];
    d.upper[i] = d.max[i] - d.dx[i];
    d.inv_dx[i] = 1.0 / d.dx[i];
    d.size *= d.n[i];
    d.vol_element *= d.dx[i];  // lib/gaussian_grid.h:201-203
  }
}

static int alloc_records(edm_grid* g) {
  GridDesc& d = g->d;
  size_t bytes = (size_t)d.size * d.rec_w * sizeof(double);
  EDM_CUDA(cudaMalloc(&d.rec, bytes));
  EDM_CUDA(cudaMemset(d.rec, 0, bytes));
  EDM_CUDA(cudaMalloc(&g->d_flags, 4 * sizeof(int)));
  EDM_CUDA(cudaMemset(g->d_flags, 0, 4 * sizeof(int)));
  return EDM_OK;
}

// GaussGrid::set_boundary (lib/gaussian_grid.h:378-435) + the per-point tables derived from it +
// the index pairs of duplicate_boundary (lib/gaussian_grid.h:571-630).  Host libm on purpose: the
// tables come out bit-identical to the reference's.
static int build_boundary(edm_grid* g, const double* mn, const double* mx, const int* per) {
  GridDesc& d = g->d;
  for (int i = 0; i < d.dim; i++) {
    d.bmin[i] = mn[i];
    d.bmax[i] = mx[i];
    d.blen[i] = mx[i] - mn[i];
    d.bper[i] = per[i];
    d.sqrtpi_sigma[i] = sqrt(M_PI) * d.sigma[i];
  }
  for (int i = 0; i < d.dim; i++) {
    std::vector<double> tab((size_t)d.n[i] * kPtabW, 0.0);
    if (!d.bper[i]) {
      g->bc_denom[i].assign(kBcTableSize, 0.0);
      g->bc_deriv[i].assign(kBcTableSize, 0.0);
      const double sg = d.sigma[i], lo = d.bmin[i], hi = d.bmax[i];
      for (size_t j = 0; j < (size_t)kBcTableSize; j++) {
        double s = j * (hi - lo) / (kBcTableSize - 1) + lo;
        double tmp1 = sqrt(M_PI) * sg / 2. * (erf((s - lo) / sg) + erf((hi - s) / sg));
        double den = tmp1;
        double tmp2 = sqrt(M_PI) * sg / 2. * erf((hi - lo) / sg);
        den += (tmp2 - tmp1) * host_sigmoid((s - lo) / (kBcMar * sg));
        den += (tmp2 - tmp1) * host_sigmoid((hi - s) / (kBcMar * sg));
        g->bc_denom[i][j] = den;
        double tmp3 = 1. * (exp(-pow(s - lo, 2) / pow(sg, 2)) - exp(-pow(hi - s, 2) / pow(sg, 2)));
        double dd = tmp3;
        dd += (tmp2 - tmp1) * host_sigmoid_dx((s - lo) / (kBcMar * sg)) / (kBcMar * sg) -
              tmp3 * host_sigmoid((s - lo) / (kBcMar * sg));
        dd += -(tmp2 - tmp1) * host_sigmoid_dx((hi - s) / (kBcMar * sg)) / (kBcMar * sg) -
              tmp3 * host_sigmoid((hi - s) / (kBcMar * sg));
        g->bc_deriv[i][j] = dd;
      }
    }
    for (int k = 0; k < d.n[i]; k++) {
      double* row = &tab[(size_t)k * kPtabW];
      double xx = d.min[i] + d.dx[i] * (size_t)k;  // lib/gaussian_grid.h:270
      row[0] = xx;
      row[1] = 1.0;
      if (!d.bper[i]) {
        if (xx < d.bmin[i] || xx > d.bmax[i]) {  // lib/gaussian_grid.h:273-276
          row[1] = 0.0;
          continue;
        }
        const double sg = d.sigma[i];
        size_t bc_index = (size_t)((kBcTableSize - 1) * (xx - d.bmin[i]) / (d.bmax[i] - d.bmin[i]));  // T12
        if (bc_index >= (size_t)kBcTableSize) bc_index = kBcTableSize - 1;
        row[2] = host_sigmoid((xx - d.bmin[i]) / (sg * kBcMar));
        row[3] = host_sigmoid((d.bmax[i] - xx) / (sg * kBcMar));
        row[4] = host_sigmoid_dx((xx - d.bmin[i]) / (sg * kBcMar)) / (kBcMar * sg);
        row[5] = -host_sigmoid_dx((d.bmax[i] - xx) / (sg * kBcMar)) / (kBcMar * sg);
        row[6] = g->bc_denom[i][bc_index];
        row[7] = g->bc_deriv[i][bc_index];
      }
    }
    if (!g->d_ptab[i]) EDM_CUDA(cudaMalloc(&g->d_ptab[i], tab.size() * sizeof(double)));
    EDM_CUDA(cudaMemcpy(g->d_ptab[i], tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    d.ptab[i] = g->d_ptab[i];
  }
  // duplicate_boundary index pairs
  long long min_i[3], max_i[3];
  for (int i = 0; i < d.dim; i++) {
    min_i[i] = host_get_index(d, i, d.bmin[i]);
    max_i[i] = host_get_index(d, i, d.bmax[i]);
    while (min_i[i] * d.dx[i] + d.min[i] < d.bmin[i]) min_i[i] += 1;
    while (max_i[i] * d.dx[i] + d.min[i] > d.bmax[i] || max_i[i] == d.n[i]) max_i[i] -= 1;
  }
  std::vector<long long> pairs;
  int combos = 1;
  for (int i = 0; i < d.dim; i++) combos *= 4;
  for (int c = 0; c < combos; c++) {
    int flag = 0, t = c;
    long long outer[3], bound[3];
    for (int j = 0; j < d.dim; j++) {
      int off = t % 4;
      t /= 4;
      switch (off) {
        case 0:
          flag |= d.bper[j];
          flag |= (min_i[j] == 0);
          outer[j] = min_i[j] - 1;
          bound[j] = min_i[j];
          break;
        case 1:
          outer[j] = bound[j] = min_i[j];
          break;
        case 2:
          outer[j] = bound[j] = max_i[j];
          break;
        default:
          flag |= d.bper[j];
          flag |= (max_i[j] == d.n[j] - 1);
          outer[j] = max_i[j] + 1;
          bound[j] = max_i[j];
          break;
      }
    }
    if (flag) continue;
    long long lo = 0, lb = 0, stride = 1;
    bool ok = true;
    for (int j = 0; j < d.dim; j++) {
      if (outer[j] < 0 || outer[j] >= d.n[j] || bound[j] < 0 || bound[j] >= d.n[j]) ok = false;
      lo += outer[j] * stride;
      lb += bound[j] * stride;
      stride *= d.n[j];
    }
    if (ok && lo != lb) {
      pairs.push_back(lo);
      pairs.push_back(lb);
    }
  }
  if (g->d_dup) {
    cudaFree(g->d_dup);
    g->d_dup = nullptr;
  }
  d.n_dup = (int)(pairs.size() / 2);
  d.dup_pairs = nullptr;
  if (d.n_dup) {
    EDM_CUDA(cudaMalloc(&g->d_dup, pairs.size() * sizeof(long long)));
    EDM_CUDA(cudaMemcpy(g->d_dup, pairs.data(), pairs.size() * sizeof(long long), cudaMemcpyHostToDevice));
    d.dup_pairs = g->d_dup;
  }
  return EDM_OK;
}

// ------------------------------------------------------------------ kernels: layout conversion

__global__ void interleave_kernel(GridDesc g, const double* values, const double* derivs) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < g.size; p += stride) {
    double* r = g.rec + p * g.rec_w;
    r[0] = values[p];
    for (int d = 0; d < g.dim; d++) r[1 + d] = derivs ? derivs[p * g.dim + d] : 0.0;
    for (int d = 1 + g.dim; d < g.rec_w; d++) r[d] = 0.0;
  }
}
__global__ void deinterleave_kernel(GridDesc g, double* values, double* derivs) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < g.size; p += stride) {
    const double* r = g.rec + p * g.rec_w;
    values[p] = r[0];
    if (derivs)
      for (int d = 0; d < g.dim; d++) derivs[p * g.dim + d] = r[1 + d];
  }
}

// ------------------------------------------------------------------ kernels: evaluation

// mode 0: get_value_deriv, mode 1: get_value
template <int DIM>
__global__ void __launch_bounds__(256) eval_kernel(GridDesc g, long n, const double* __restrict__ x, long xs,
                                                   double* __restrict__ val, double* __restrict__ der, int mode) {
  long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double xi[DIM], dr[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) xi[d] = x[i * xs + d];
    double v;
    if (mode == 0)
      v = d_eval_point<DIM>(g, xi, dr, g.b_interp != 0);
    else
      v = d_get_value<DIM>(g, xi);
    val[i] = v;
    if (der && mode == 0)
#pragma unroll
      for (int d = 0; d < DIM; d++) der[i * DIM + d] = dr[d];
  }
}

// DimmedGrid::add_value, lib/grid.h:370-385: one thread, in list order (bumps may collide)
template <int DIM> __global__ void hist_add_kernel(GridDesc g, long n, const double* x, long xs, const double* v) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (long i = 0; i < n; i++) {
    bool inside = true;
    long long lin = 0, pstride = 1;
    for (int d = 0; d < DIM; d++) {
      double xd = x[i * xs + d];
      if (!g.periodic[d] && (xd < g.min[d] || xd >= g.upper[d])) inside = false;
      if (g.periodic[d]) xd = d_wrap(xd, g.min[d], g.len[d]);
      long long idx = (long long)floor(__ddiv_rn(__dsub_rn(xd, g.min[d]), g.dx[d]));
      idx = idx < 0 ? 0 : (idx > g.n[d] - 1 ? g.n[d] - 1 : idx);
      lin += idx * pstride;
      pstride *= g.n[d];
    }
    if (inside) g.rec[lin * g.rec_w] += v[i];
  }
}

// Grid::add, lib/grid.h:275-290
template <int DIM> __global__ void grid_add_kernel(GridDesc g, GridDesc other, double scale, double offset) {
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < g.size; p += stride) {
    double x[DIM], der[DIM];
    long long t = p;
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      long long idx = (d < DIM - 1) ? t % g.n[d] : t;
      t = (t - idx) / g.n[d];
      x[d] = __dadd_rn(g.min[d], __dmul_rn(g.dx[d], (double)idx));
    }
    double v = d_eval_point<DIM>(other, x, der, other.b_interp != 0);
    double* r = g.rec + p * g.rec_w;
    r[0] += scale * v + offset;
#pragma unroll
    for (int d = 0; d < DIM; d++) r[1 + d] += scale * der[d];
  }
}

__global__ void minmax_kernel(GridDesc g, double* out) {  // lib/grid.h:292-309; tiny, setup-time only
  __shared__ double smin[256], smax[256];
  double mn = g.rec[0], mx = g.rec[0];
  for (long long p = threadIdx.x; p < g.size; p += blockDim.x) {
    double v = g.rec[p * g.rec_w];
    mn = fmin(mn, v);
    mx = fmax(mx, v);
  }
  smin[threadIdx.x] = mn;
  smax[threadIdx.x] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < blockDim.x; i++) {
      mn = fmin(mn, smin[i]);
      mx = fmax(mx, smax[i]);
    }
    out[0] = mn;
    out[1] = mx;
  }
}

template <int DIM> __global__ void remap_kernel(GridDesc g, double* x) {
  double v[DIM];
  for (int d = 0; d < DIM; d++) v[d] = x[d];
  d_remap<DIM>(g, v);
  for (int d = 0; d < DIM; d++) x[d] = v[d];
}

// ------------------------------------------------------------------ kernels: deposit

// duplicate_boundary, lib/gaussian_grid.h:571-630 (T13: values only)
__global__ void dup_boundary_kernel(GridDesc g, int* flags, const int* __restrict__ gate, int want) {
  pdl_trigger();
  pdl_wait();
  if (gate && *gate < want) return;
  if (flags[0] == 0) return;
  for (int k = threadIdx.x; k < g.n_dup; k += blockDim.x)
    g.rec[g.dup_pairs[2 * k] * g.rec_w] = g.rec[g.dup_pairs[2 * k + 1] * g.rec_w];
  __syncthreads();
  if (threadIdx.x == 0) flags[0] = 0;
}

// Hill-parallel deposit (2-D/3-D, and 1-D grids the owner kernel does not take): one CTA per hill,
// threads over the window, fp64 RED atomics into the records because hills of one batch may
// overlap.  Per-point summation order is then not the list order; the difference is rounding only.
template <int DIM>
__global__ void __launch_bounds__(256) deposit_hills_kernel(GridDesc g, long n, const double* __restrict__ centres,
                                                            const double* __restrict__ heights,
                                                            double* __restrict__ ba_out, int* flags) {
  __shared__ double red[33];
  __shared__ AxisEntry s_axis[DIM > 1 ? DIM * kAxisMax : 1];
  for (long hill = blockIdx.x; hill < n; hill += gridDim.x) {
    double x0[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) x0[d] = centres[hill * DIM + d];
    bool dirty;
    const double tot = cta_window_pass<DIM, kPassAtomic>(g, x0, heights[hill], red, dirty, s_axis);
    if (threadIdx.x == 0 && ba_out) ba_out[hill] = tot;
    if (dirty) flags[0] = 1;
  }
}

// ---- 1-D owner-computes deposit (the hills/s hot kernel for the pair-RDF config) ----
//
// A warp owns 32 consecutive grid points (one per lane) and keeps its point's constants in
// registers; the hill list of its chunk is scanned 32 hills at a time with a ballot, and every
// overlapping hill is applied in list order, so each grid point sums its hills exactly in the
// order the reference would have deposited them and no atomics are needed.  Per-hill integrals
// (the add_value return) are warp-reduced and written to one slot per (hill, warp).

struct Hill1D {
  double x, h, t1, t3;
  int xi;
  int ok;
};

__global__ void deposit1d_prepare_kernel(GridDesc g, long n, const int* __restrict__ n_dev,
                                         const double* __restrict__ centres, const double* __restrict__ heights,
                                         Hill1D* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = (*n_dev < n) ? *n_dev : n;
  if (i >= n) return;
  HillGeom<1> hg;
  double x0[1] = {centres[i]};
  bool ok = d_hill_prepare<1>(g, x0, hg);
  Hill1D o;
  o.x = hg.x[0];
  o.h = heights[i];
  o.t1 = ok ? hg.t1[0] : 0.0;
  o.t3 = ok ? hg.t3[0] : 0.0;
  o.xi = ok ? hg.xi[0] : 0;
  o.ok = ok ? 1 : 0;
  out[i] = o;
}

__device__ __forceinline__ int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <bool GPER>
__global__ void __launch_bounds__(128) deposit1d_owner_kernel(GridDesc g, long n, const int* __restrict__ n_dev,
                                                              const Hill1D* __restrict__ hills, long chunk, int nslot,
                                                              double* __restrict__ partial,
                                                              double* __restrict__ ba_slots, int* flags) {
  pdl_trigger();
  pdl_wait();
  if (n_dev) n = (*n_dev < n) ? *n_dev : n;
  const int lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  const int nwarps = (g.n[0] + 31) >> 5;
  const int w = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
  if (w >= nwarps) return;
  const int c = blockIdx.y;
  const long h0 = (long)c * chunk;
  const long h1 = (h0 + chunk < n) ? h0 + chunk : n;
  const int npts = g.n[0];
  const int m = g.minisize[0];
  const int ms = g.supp[0];  // hills further than this from the warp's 32 points cannot reach them
  const int p0 = w * 32;
  const int p = p0 + lane;
  const bool have = p < npts;
  const bool bper = g.bper[0] != 0;
  const bool all_overlap = GPER && (2 * m + 1 + 32 >= npts);

  // everything that depends on the grid point only lives in registers for the whole launch
  double xx = 0, uL = 0, uU = 0, t6 = 0, t7 = 0, Z = 1, Zd = 0;
  bool valid = false;
  if (have) {
    const double* row = g.ptab[0] + (long long)p * kPtabW;
    xx = row[0];
    valid = row[1] != 0.0;
    uL = row[2];
    uU = row[3];
    t6 = row[4];
    t7 = row[5];
    Z = row[6];
    Zd = row[7];
  }
  const double sigma = g.sigma[0], inv_sigma = 1.0 / sigma, len = g.len[0], vol = g.vol_element;
  const double invZ = bper ? 1.0 / g.sqrtpi_sigma[0] : 1.0 / Z;
  const double invZ2 = invZ * invZ;
  const double ZdinvZ2 = Zd * invZ2;
  const bool wall = !bper && (uL != 0.0 || uU != 0.0 || t6 != 0.0 || t7 != 0.0);  // within 2 sigma~ of a wall
  // chunk 0 starts from the stored value so that a single-chunk launch reproduces the reference's
  // rounding sequence ((g + t1) + t2) + ... exactly
  double acc_v = 0.0, acc_d = 0.0;
  if (c == 0 && have) {
    double2 r = *reinterpret_cast<const double2*>(g.rec + (long long)p * 2);
    acc_v = r.x;
    acc_d = r.y;
  }
  bool dirty = false;

  for (long base = h0; base < h1; base += 32) {
    long mine = base + lane;
    bool ov = false;
    if (mine < h1) {
      const int ok = hills[mine].ok;
      const int xi = hills[mine].xi;
      if (ok) {
        if (GPER) {
          int d = (p0 - xi) % npts;
          if (d < 0) d += npts;
          ov = all_overlap || (d <= ms) || (d >= npts - ms - 31);
        } else {
          ov = (xi - ms <= p0 + 31) && (xi + ms >= p0);
        }
      }
    }
    unsigned mask = __ballot_sync(0xffffffffu, ov);
    while (mask) {
      const int b = __ffs(mask) - 1;
      mask &= mask - 1;
      const Hill1D hl = hills[base + b];  // same address across the warp: one broadcast load
      int mult = 0;
      if (valid) {
        if (GPER) {
          int o0 = (p - hl.xi) % npts;
          if (o0 < 0) o0 += npts;
          // window offsets o0 - k*npts inside [-m, m]: a window wider than the grid revisits points
          const int khi = floor_div(o0 + m, npts), klo = -floor_div(m - o0, npts);
          mult = khi - klo + 1;
          if (mult < 0) mult = 0;
        } else {
          const int o = p - hl.xi;
          mult = (o >= -m && o <= m) ? 1 : 0;
        }
      }
      // Straight-line body: lanes outside the window or the support add an exact 0.0 (the divergence
      // bookkeeping of a nested if cost more issue slots than the arithmetic it skipped).
      double term_ba = 0.0;
      {
        double v = __dsub_rn(xx, hl.x);
        if (GPER) v = __dsub_rn(v, __dmul_rn(d_round(__ddiv_rn(v, len)), len));
        // reciprocal first; the exactly rounded quotient only where it could change the support test
        double dp = v * inv_sigma;
        double dp2 = dp * dp;
        if (fabs(dp2 - kGaussSupport) < 1e-12) {
          dp = __ddiv_rn(v, sigma);
          dp2 = __dmul_rn(dp, dp);
        }
        const bool in = mult > 0 && dp2 < kGaussSupport;
        const double E = d_exp_support(in ? -dp2 : 0.0);
        const double t5 = -2.0 * dp * inv_sigma;
        double etot, F;
        if (wall) {  // McGDP + zero-force terms, lib/gaussian_grid.h:310-337
          double corr = (hl.t1 - E) * uL + (hl.t3 - E) * uU;
          F = t5 * E + (hl.t1 - E) * t6 - t5 * E * uL + (hl.t3 - E) * t7 - t5 * E * uU;
          F = (F * Z - Zd * (E + corr)) * invZ2;
          corr *= invZ;
          etot = E * invZ + corr;
          dirty |= in && (corr * corr > 0.0);
        } else if (!bper) {  // interior of a non-periodic boundary: the correction terms vanish
          etot = E * invZ;
          F = E * (t5 * invZ - ZdinvZ2);
        } else {  // periodic boundary, lib/gaussian_grid.h:340,352
          etot = E * invZ;
          F = t5 * etot;
        }
        const double add = in ? hl.h * etot : 0.0;
        const double addd = in ? hl.h * F : 0.0;
        if (GPER) {
          for (int k = 0; k < mult; k++) {
            acc_v += add;
            acc_d += addd;
            term_ba += add * vol;
          }
        } else {
          acc_v += add;
          acc_d += addd;
          term_ba = add * vol;
        }
      }
      const double tot = warp_sum(term_ba);
      if (lane == 0) {
        int slot;
        if (GPER) {
          int s = (hl.xi - m) % npts;
          if (s < 0) s += npts;
          slot = (w - (s >> 5) + nwarps) % nwarps;
        } else {
          slot = w - floor_div(hl.xi - m, 32);
        }
        if (slot >= 0 && slot < nslot) ba_slots[(base + b) * (long)nslot + slot] = tot;
      }
    }
  }
  if (have) {
    double2 o = make_double2(acc_v, acc_d);
    *reinterpret_cast<double2*>(partial + ((long long)c * npts + p) * 2) = o;
  }
  if (dirty) flags[0] = 1;
}

__global__ void deposit1d_commit_kernel(GridDesc g, int nchunks, const double* __restrict__ partial,
                                        const int* __restrict__ flag, int want) {
  pdl_trigger();
  pdl_wait();
  if (flag && *flag < want) return;
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= g.n[0]) return;
  double v = 0.0, dv = 0.0;
  for (int c = 0; c < nchunks; c++) {
    double2 r = *reinterpret_cast<const double2*>(partial + ((long long)c * g.n[0] + p) * 2);
    if (c == 0) {
      v = r.x;
      dv = r.y;
    } else {
      v += r.x;
      dv += r.y;
    }
  }
  *reinterpret_cast<double2*>(g.rec + (long long)p * 2) = make_double2(v, dv);
}

__global__ void deposit1d_ba_kernel(long n, const int* __restrict__ n_dev, int nslot,
                                    const double* __restrict__ ba_slots, double* __restrict__ ba) {
  pdl_trigger();
  pdl_wait();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = (*n_dev < n) ? *n_dev : n;
  if (i >= n) return;
  double t = 0.0;
  for (int s = 0; s < nslot; s++) t += ba_slots[i * (long)nslot + s];
  ba[i] = t;
}

}  // namespace edm

using namespace edm;

namespace edm {

static int deposit1d_nslot(const GridDesc& d) {
  const int nwarps = (d.n[0] + 31) / 32;
  int nslot = (2 * d.minisize[0] + 1 + 31) / 32 + 2;
  return nslot > nwarps ? nwarps : nslot;
}

bool deposit1d_eligible(const edm_grid* g) { return g->d.dim == 1 && g->d.is_gauss && g->d.minisize[0] < g->d.n[0]; }

// one chunk: every grid point applies the hills in list order starting from its stored value, which
// is exactly the reference's per-point rounding sequence
int deposit1d_stage(edm_grid* g, const double* centres, const double* heights, double* ba, const int* n_dev,
                    long n_max, cudaStream_t st) {
  const GridDesc& d = g->d;
  const int npts = d.n[0], nwarps = (npts + 31) / 32, nslot = deposit1d_nslot(d);
  size_t b_h = ((size_t)n_max * sizeof(Hill1D) + 255) / 256 * 256;
  size_t b_p = ((size_t)npts * 2 * sizeof(double) + 255) / 256 * 256;
  size_t b_s = (size_t)n_max * nslot * sizeof(double);
  EDM_TRY(g->work.reserve(b_h + b_p + b_s));
  char* basep = g->work.as<char>();
  Hill1D* hl = reinterpret_cast<Hill1D*>(basep);
  double* partial = reinterpret_cast<double*>(basep + b_h);
  double* slots = reinterpret_cast<double*>(basep + b_h + b_p);
  if (ba) EDM_CUDA(cudaMemsetAsync(slots, 0, b_s, st));
  EDM_CUDA(launch_pdl(deposit1d_prepare_kernel, dim3((unsigned)((n_max + 255) / 256)), dim3(256), 0, st, d, n_max, n_dev,
                      centres, heights, hl));
  long chunk = (n_max + 31) / 32 * 32;
  dim3 grid((nwarps + 3) / 4, 1);
  if (d.periodic[0])
    EDM_CUDA(launch_pdl(deposit1d_owner_kernel<true>, grid, dim3(128), 0, st, d, n_max, n_dev, (const Hill1D*)hl, chunk, nslot,
                        partial, slots, g->d_flags));
  else
    EDM_CUDA(launch_pdl(deposit1d_owner_kernel<false>, grid, dim3(128), 0, st, d, n_max, n_dev, (const Hill1D*)hl, chunk, nslot,
                        partial, slots, g->d_flags));
  if (ba) deposit1d_ba_kernel<<<(unsigned)((n_max + 255) / 256), 256, 0, st>>>(n_max, n_dev, nslot, slots, ba);
  g->stage_partial = partial;
  count_launches(ba ? 3 : 2);
  EDM_CUDA(cudaGetLastError());
  return EDM_OK;
}

int deposit1d_commit_if(edm_grid* g, const int* flag, int want, cudaStream_t st) {
  const GridDesc& d = g->d;
  const int npts = d.n[0];
  EDM_CUDA(launch_pdl(deposit1d_commit_kernel, dim3((npts + 255) / 256), dim3(256), 0, st, d, 1, (const double*)g->stage_partial,
                      flag, want));
  count_launches(1);
  if (d.n_dup) {
    EDM_CUDA(launch_pdl(dup_boundary_kernel, dim3(1), dim3(64), 0, st, d, g->d_flags, flag, want));
    count_launches(1);
  }
  EDM_CUDA(cudaGetLastError());
  return EDM_OK;
}

int edm_grid_dup_boundary_if(edm_grid* g, const int* gate, int want, cudaStream_t st) {
  if (!g->d.n_dup) return EDM_OK;
  dup_boundary_kernel<<<1, 64, 0, st>>>(g->d, g->d_flags, gate, want);
  EDM_CUDA(cudaGetLastError());
  return EDM_OK;
}

}  // namespace edm

// ------------------------------------------------------------------ C ABI: grids

extern "C" {

const char* edm_last_error(void) { return g_last_error.c_str(); }

int edm_device_count(int* count) {
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    c = 0;
  }
  if (count) *count = c;
  return EDM_OK;
}

double edm_uniform_pair(uint64_t seed, uint64_t step, uint64_t pairkey, int which) {
  return pair_uniform_from_bits(pair_bits(uniform_key(seed, step), pairkey), which);
}

int edm_host_pin(void* ptr, size_t bytes) {
  EDM_REQUIRE(ptr && bytes, "NULL argument");
  EDM_TRY(ensure_device(0));
  EDM_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return EDM_OK;
}

int edm_host_unpin(void* ptr) {
  EDM_REQUIRE(ptr != nullptr, "NULL argument");
  EDM_CUDA(cudaHostUnregister(ptr));
  return EDM_OK;
}

int edm_launch_count(long long* count) {
  if (count) *count = g_launches.load();
  return EDM_OK;
}

double edm_uniform(uint64_t seed, uint64_t step, uint64_t counter) {
  return uniform_from_key(uniform_key(seed, step), counter);
}

static int grid_new(edm_grid_t** out, int device, int dim) {
  EDM_REQUIRE(out != nullptr, "out is NULL");
  EDM_REQUIRE(dim >= 1 && dim <= 3, "dimension must be 1, 2 or 3");
  EDM_TRY(ensure_device(device));
  edm_grid* g = new edm_grid();
  memset(&g->d, 0, sizeof(g->d));
  g->device = device;
  g->d.dim = dim;
  *out = g;
  return EDM_OK;
}

int edm_grid_create(edm_grid_t** out, int device, int dim, const double* mn, const double* mx,
                    const double* spacing, const int* periodic, int b_derivatives, int b_interpolate) {
  EDM_TRY(grid_new(out, device, dim));
  edm_grid* g = *out;
  GridDesc& d = g->d;
  d.b_deriv = b_derivatives;
  d.b_interp = b_interpolate;
  for (int i = 0; i < dim; i++) {  // lib/grid.h:199-211 (T3)
    d.min[i] = mn[i];
    d.max[i] = mx[i];
    d.periodic[i] = periodic[i];
    d.n[i] = (int)ceil((d.max[i] - d.min[i]) / spacing[i]);
    d.dx[i] = (d.max[i] - d.min[i]) / d.n[i];
    d.n[i] = d.periodic[i] ? d.n[i] : d.n[i] + 1;
    if (!d.periodic[i]) d.max[i] += d.dx[i];
  }
  finish_geometry(g);
  int rc = alloc_records(g);
  if (rc != EDM_OK) {
    delete g;
    *out = nullptr;
  }
  return rc;
}

int edm_grid_create_from_header(edm_grid_t** out, int device, int dim, const int* bins, const double* mn,
                                const double* mx, const int* periodic, int b_derivatives, int b_interpolate) {
  EDM_TRY(grid_new(out, device, dim));
  edm_grid* g = *out;
  GridDesc& d = g->d;
  d.b_deriv = b_derivatives;
  d.b_interp = b_interpolate;
  for (int i = 0; i < dim; i++) {  // lib/grid.h:800-806
    d.min[i] = mn[i];
    d.max[i] = mx[i];
    d.periodic[i] = periodic[i];
    d.n[i] = bins[i];
    d.dx[i] = (d.max[i] - d.min[i]) / d.n[i];
    if (!d.periodic[i]) {
      d.max[i] += d.dx[i];
      d.n[i] += 1;
    }
  }
  finish_geometry(g);
  int rc = alloc_records(g);
  if (rc != EDM_OK) {
    delete g;
    *out = nullptr;
  }
  return rc;
}

int edm_gauss_create(edm_grid_t** out, int device, int dim, const double* mn, const double* mx,
                     const double* spacing, const int* periodic, int b_interpolate, const double* sigma) {
  EDM_TRY(edm_grid_create(out, device, dim, mn, mx, spacing, periodic, 1, b_interpolate));
  edm_grid* g = *out;
  GridDesc& d = g->d;
  d.is_gauss = 1;
  d.dup_possible = 0;
  for (int i = 0; i < dim; i++) {
    g->sigma_user[i] = sigma[i];
    d.sigma[i] = sigma[i] * sqrt(2.);  // lib/gaussian_grid.h:75 (T8)
  }
  // update_minigrid, lib/gaussian_grid.h:559-569
  for (int i = 0; i < dim; i++) {
    double dist = sqrt(2 * kGaussSupport) * d.sigma[i];
    d.minisize[i] = host_int_floor(dist / d.dx[i]);
    if (d.periodic[i] && 2 * d.minisize[i] + 1 > d.n[i]) d.dup_possible = 1;
  }
  // The window is 5.66 sigma wide each way but only points with sum dp^2 < 8 receive anything
  // (lib/gaussian_grid.h:299): |xx - x| < sqrt(8) sigma~ in every dim.  A point more than `supp`
  // cells from the centre cell lies at least (supp - 1) dx > sqrt(8) sigma~ away, so loops may stop
  // at supp and still visit every point the reference changes.  Not when a periodic window revisits
  // points (each revisit adds again).
  for (int i = 0; i < dim; i++) {
    int s = host_int_floor(sqrt(kGaussSupport) * d.sigma[i] / d.dx[i]) + 2;
    d.supp[i] = (d.dup_possible || s > d.minisize[i]) ? d.minisize[i] : s;
  }
  // the ctor calls set_boundary with the un-extended box (lib/gaussian_grid.h:78, T9)
  int rc = build_boundary(g, mn, mx, periodic);
  if (rc != EDM_OK) {
    edm_grid_destroy(g);
    *out = nullptr;
  }
  return rc;
}

int edm_grid_destroy(edm_grid_t* g) {
  if (!g) return EDM_OK;
  cudaSetDevice(g->device);
  if (g->d.rec) cudaFree(g->d.rec);
  for (int i = 0; i < 3; i++)
    if (g->d_ptab[i]) cudaFree(g->d_ptab[i]);
  if (g->d_dup) cudaFree(g->d_dup);
  if (g->d_flags) cudaFree(g->d_flags);
  g->io.release();
  g->work.release();
  delete g;
  return EDM_OK;
}

int edm_grid_set_boundary(edm_grid_t* g, const double* mn, const double* mx, const int* periodic) {
  EDM_REQUIRE(g && g->d.is_gauss, "set_boundary needs a GaussGrid");
  EDM_TRY(ensure_device(g->device));
  EDM_CUDA(cudaDeviceSynchronize());
  return build_boundary(g, mn, mx, periodic);
}

int edm_grid_geometry(const edm_grid_t* g, int* dim, int* n, double* dx, double* mn, double* mx, int* periodic,
                      int* minisize, size_t* size) {
  EDM_REQUIRE(g != nullptr, "grid is NULL");
  const GridDesc& d = g->d;
  if (dim) *dim = d.dim;
  for (int i = 0; i < d.dim; i++) {
    if (n) n[i] = d.n[i];
    if (dx) dx[i] = d.dx[i];
    if (mn) mn[i] = d.min[i];
    if (mx) mx[i] = d.max[i];
    if (periodic) periodic[i] = d.periodic[i];
    if (minisize) minisize[i] = d.is_gauss ? d.minisize[i] : 0;
  }
  if (size) *size = (size_t)d.size;
  return EDM_OK;
}

int edm_grid_flags(const edm_grid_t* g, int* b_derivatives, int* b_interpolate, int* is_gauss) {
  EDM_REQUIRE(g != nullptr, "grid is NULL");
  if (b_derivatives) *b_derivatives = g->d.b_deriv;
  if (b_interpolate) *b_interpolate = g->d.b_interp;
  if (is_gauss) *is_gauss = g->d.is_gauss;
  return EDM_OK;
}

int edm_grid_boundary(const edm_grid_t* g, double* bmin, double* bmax, int* bperiodic, double* sigma_sqrt2) {
  EDM_REQUIRE(g && g->d.is_gauss, "boundary needs a GaussGrid");
  for (int i = 0; i < g->d.dim; i++) {
    if (bmin) bmin[i] = g->d.bmin[i];
    if (bmax) bmax[i] = g->d.bmax[i];
    if (bperiodic) bperiodic[i] = g->d.bper[i];
    if (sigma_sqrt2) sigma_sqrt2[i] = g->d.sigma[i];
  }
  return EDM_OK;
}

int edm_grid_set_interpolation(edm_grid_t* g, int b) {
  EDM_REQUIRE(g != nullptr, "grid is NULL");
  g->d.b_interp = b;
  return EDM_OK;
}

static int launch_blocks(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  int dev = 0;
  cudaGetDevice(&dev);  // callers have selected the grid's device (ensure_device)
  long long cap = (long long)sm_count(dev) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

int edm_grid_upload(edm_grid_t* g, const double* values, const double* derivs) {
  EDM_REQUIRE(g && values, "grid/values is NULL");
  EDM_TRY(ensure_device(g->device));
  const GridDesc& d = g->d;
  size_t nv = (size_t)d.size, nd = derivs ? nv * d.dim : 0;
  EDM_TRY(g->io.reserve((nv + nd) * sizeof(double)));
  double* dv = g->io.as<double>();
  double* dd = derivs ? dv + nv : nullptr;
  EDM_CUDA(cudaMemcpy(dv, values, nv * sizeof(double), cudaMemcpyHostToDevice));
  if (derivs) EDM_CUDA(cudaMemcpy(dd, derivs, nd * sizeof(double), cudaMemcpyHostToDevice));
  interleave_kernel<<<launch_blocks(d.size, 256), 256>>>(d, dv, dd);
  EDM_CUDA(cudaGetLastError());
  EDM_CUDA(cudaDeviceSynchronize());
  return EDM_OK;
}

int edm_grid_download(const edm_grid_t* gc, double* values, double* derivs) {
  edm_grid* g = const_cast<edm_grid*>(gc);
  EDM_REQUIRE(g && values, "grid/values is NULL");
  EDM_TRY(ensure_device(g->device));
  const GridDesc& d = g->d;
  size_t nv = (size_t)d.size, nd = derivs ? nv * d.dim : 0;
  EDM_TRY(g->io.reserve((nv + nd) * sizeof(double)));
  double* dv = g->io.as<double>();
  double* dd = derivs ? dv + nv : nullptr;
  deinterleave_kernel<<<launch_blocks(d.size, 256), 256>>>(d, dv, dd);
  EDM_CUDA(cudaGetLastError());
  EDM_CUDA(cudaMemcpy(values, dv, nv * sizeof(double), cudaMemcpyDeviceToHost));
  if (derivs) EDM_CUDA(cudaMemcpy(derivs, dd, nd * sizeof(double), cudaMemcpyDeviceToHost));
  return EDM_OK;
}

int edm_grid_clear(edm_grid_t* g) {
  EDM_REQUIRE(g != nullptr, "grid is NULL");
  EDM_TRY(ensure_device(g->device));
  EDM_CUDA(cudaMemset(g->d.rec, 0, (size_t)g->d.size * g->d.rec_w * sizeof(double)));
  return EDM_OK;
}

static int eval_launch(const edm_grid* g, long n, const double* x, long xs, double* val, double* der, int mode,
                       cudaStream_t st) {
  if (n <= 0) return EDM_OK;
  count_launches(1);
  GridDesc d = g->d;
  if (mode == 2) {  // the plain grid inside a GaussGrid: no boundary test, no remap (lib/grid.h:390-446 only)
    d.is_gauss = 0;
    mode = 0;
  }
  int blocks = launch_blocks(n, 256);
  switch (d.dim) {
    case 1: eval_kernel<1><<<blocks, 256, 0, st>>>(d, n, x, xs, val, der, mode); break;
    case 2: eval_kernel<2><<<blocks, 256, 0, st>>>(d, n, x, xs, val, der, mode); break;
    default: eval_kernel<3><<<blocks, 256, 0, st>>>(d, n, x, xs, val, der, mode); break;
  }
  EDM_CUDA(cudaGetLastError());
  return EDM_OK;
}

int edm_grid_eval_dev(const edm_grid_t* g, long n, const double* x, long xstride, double* value, double* der,
                      void* stream) {
  EDM_REQUIRE(g && x && value, "NULL argument");
  EDM_REQUIRE(xstride >= g->d.dim, "xstride < dim");
  EDM_TRY(ensure_device(g->device));
  return eval_launch(g, n, x, xstride, value, der, 0, (cudaStream_t)stream);
}

static int eval_host(const edm_grid* gc, long n, const double* x, long xs, double* value, double* der, int mode) {
  edm_grid* g = const_cast<edm_grid*>(gc);
  EDM_REQUIRE(g && x && value, "NULL argument");
  EDM_REQUIRE(xs >= g->d.dim, "xstride < dim");
  EDM_TRY(ensure_device(g->device));
  if (n <= 0) return EDM_OK;
  int D = g->d.dim;
  size_t nx = (size_t)n * xs, nd = der ? (size_t)n * D : 0;
  EDM_TRY(g->io.reserve((nx + n + nd) * sizeof(double)));
  double* dx = g->io.as<double>();
  double* dv = dx + nx;
  double* dd = der ? dv + n : nullptr;
  EDM_CUDA(cudaMemcpy(dx, x, nx * sizeof(double), cudaMemcpyHostToDevice));
  EDM_TRY(eval_launch(g, n, dx, xs, dv, dd, mode, 0));
  EDM_CUDA(cudaMemcpy(value, dv, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  if (der) EDM_CUDA(cudaMemcpy(der, dd, nd * sizeof(double), cudaMemcpyDeviceToHost));
  return EDM_OK;
}

int edm_grid_eval(const edm_grid_t* g, long n, const double* x, long xstride, double* value, double* der) {
  return eval_host(g, n, x, xstride, value, der, 0);
}
int edm_grid_get_value(const edm_grid_t* g, long n, const double* x, long xstride, double* value) {
  return eval_host(g, n, x, xstride, value, nullptr, 1);
}
int edm_grid_eval_plain(const edm_grid_t* g, long n, const double* x, long xstride, double* value, double* der) {
  return eval_host(g, n, x, xstride, value, der, 2);
}

int edm_grid_hist_add(edm_grid_t* g, long n, const double* x, long xs, const double* v) {
  EDM_REQUIRE(g && x && v, "NULL argument");
  if (g->d.b_interp) {  // lib/grid.h:371-373 aborts; the C ABI reports it
    set_error("Cannot add_value when using derivatives");
    return EDM_ERR_STATE;
  }
  EDM_TRY(ensure_device(g->device));
  if (n <= 0) return EDM_OK;
  size_t nx = (size_t)n * xs;
  EDM_TRY(g->io.reserve((nx + n) * sizeof(double)));
  double* dx = g->io.as<double>();
  double* dv = dx + nx;
  EDM_CUDA(cudaMemcpy(dx, x, nx * sizeof(double), cudaMemcpyHostToDevice));
  EDM_CUDA(cudaMemcpy(dv, v, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  switch (g->d.dim) {
    case 1: hist_add_kernel<1><<<1, 32>>>(g->d, n, dx, xs, dv); break;
    case 2: hist_add_kernel<2><<<1, 32>>>(g->d, n, dx, xs, dv); break;
    default: hist_add_kernel<3><<<1, 32>>>(g->d, n, dx, xs, dv); break;
  }
  EDM_CUDA(cudaGetLastError());
  EDM_CUDA(cudaDeviceSynchronize());
  return EDM_OK;
}

int edm_grid_add(edm_grid_t* g, const edm_grid_t* other, double scale, double offset) {
  EDM_REQUIRE(g && other && g->d.dim == other->d.dim, "grids missing or of different dimension");
  EDM_TRY(ensure_device(g->device));
  int blocks = launch_blocks(g->d.size, 256);
  switch (g->d.dim) {
    case 1: grid_add_kernel<1><<<blocks, 256>>>(g->d, other->d, scale, offset); break;
    case 2: grid_add_kernel<2><<<blocks, 256>>>(g->d, other->d, scale, offset); break;
    default: grid_add_kernel<3><<<blocks, 256>>>(g->d, other->d, scale, offset); break;
  }
  EDM_CUDA(cudaGetLastError());
  EDM_CUDA(cudaDeviceSynchronize());
  return EDM_OK;
}

int edm_grid_minmax(const edm_grid_t* gc, double* min_value, double* max_value) {
  edm_grid* g = const_cast<edm_grid*>(gc);
  EDM_REQUIRE(g != nullptr, "grid is NULL");
  EDM_TRY(ensure_device(g->device));
  EDM_TRY(g->io.reserve(2 * sizeof(double)));
  minmax_kernel<<<1, 256>>>(g->d, g->io.as<double>());
  EDM_CUDA(cudaGetLastError());
  double out[2];
  EDM_CUDA(cudaMemcpy(out, g->io.p, sizeof(out), cudaMemcpyDeviceToHost));
  if (min_value) *min_value = out[0];
  if (max_value) *max_value = out[1];
  return EDM_OK;
}

int edm_grid_remap(const edm_grid_t* gc, double* x) {
  edm_grid* g = const_cast<edm_grid*>(gc);
  EDM_REQUIRE(g && g->d.is_gauss && x, "remap needs a GaussGrid and a point");
  EDM_TRY(ensure_device(g->device));
  EDM_TRY(g->io.reserve(3 * sizeof(double)));
  EDM_CUDA(cudaMemcpy(g->io.p, x, g->d.dim * sizeof(double), cudaMemcpyHostToDevice));
  switch (g->d.dim) {
    case 1: remap_kernel<1><<<1, 1>>>(g->d, g->io.as<double>()); break;
    case 2: remap_kernel<2><<<1, 1>>>(g->d, g->io.as<double>()); break;
    default: remap_kernel<3><<<1, 1>>>(g->d, g->io.as<double>()); break;
  }
  EDM_CUDA(cudaGetLastError());
  EDM_CUDA(cudaMemcpy(x, g->io.p, g->d.dim * sizeof(double), cudaMemcpyDeviceToHost));
  return EDM_OK;
}

// ------------------------------------------------------------------ C ABI: batched deposit

static int deposit_1d_owner(edm_grid* g, long n, const double* centres, const double* heights, double* ba,
                            cudaStream_t st) {
  const GridDesc& d = g->d;
  const int npts = d.n[0];
  const int nwarps = (npts + 31) / 32;
  const int m = d.minisize[0];
  int nslot = (2 * m + 1 + 31) / 32 + 2;
  if (nslot > nwarps) nslot = nwarps;
  const long kSub = 1L << 18;
  for (long off = 0; off < n; off += kSub) {
    long cnt = (n - off < kSub) ? n - off : kSub;
    int nchunks = (int)((cnt + 1023) / 1024);
    if (nchunks > 24) nchunks = 24;
    if (nchunks < 1) nchunks = 1;
    long chunk = (cnt + nchunks - 1) / nchunks;
    chunk = (chunk + 31) / 32 * 32;
    nchunks = (int)((cnt + chunk - 1) / chunk);
    size_t b_h = (size_t)cnt * sizeof(Hill1D);
    size_t b_p = (size_t)nchunks * npts * 2 * sizeof(double);
    size_t b_s = (size_t)cnt * nslot * sizeof(double);
    b_h = (b_h + 255) / 256 * 256;
    b_p = (b_p + 255) / 256 * 256;
    EDM_TRY(g->work.reserve(b_h + b_p + b_s));
    char* basep = g->work.as<char>();
    Hill1D* hl = reinterpret_cast<Hill1D*>(basep);
    double* partial = reinterpret_cast<double*>(basep + b_h);
    double* slots = reinterpret_cast<double*>(basep + b_h + b_p);
    EDM_CUDA(cudaMemsetAsync(slots, 0, b_s, st));
    deposit1d_prepare_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(d, cnt, nullptr, centres + off, heights + off, hl);
    dim3 grid((nwarps + 3) / 4, nchunks);
    if (d.periodic[0])
      deposit1d_owner_kernel<true><<<grid, 128, 0, st>>>(d, cnt, nullptr, hl, chunk, nslot, partial, slots, g->d_flags);
    else
      deposit1d_owner_kernel<false><<<grid, 128, 0, st>>>(d, cnt, nullptr, hl, chunk, nslot, partial, slots, g->d_flags);
    deposit1d_commit_kernel<<<(npts + 255) / 256, 256, 0, st>>>(d, nchunks, partial, nullptr, 0);
    if (ba) deposit1d_ba_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(cnt, nullptr, nslot, slots, ba + off);
    count_launches(ba ? 4 : 3);
    EDM_CUDA(cudaGetLastError());
  }
  if (d.n_dup) dup_boundary_kernel<<<1, 64, 0, st>>>(d, g->d_flags, nullptr, 0);
  EDM_CUDA(cudaGetLastError());
  return EDM_OK;
}

int edm_gauss_deposit_dev(edm_grid_t* g, long n, const double* centres, const double* heights, double* bias_added,
                          void* stream) {
  EDM_REQUIRE(g && g->d.is_gauss, "deposit needs a GaussGrid");
  EDM_REQUIRE(n == 0 || (centres && heights), "NULL argument");
  EDM_TRY(ensure_device(g->device));
  if (n <= 0) return EDM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const GridDesc& d = g->d;
  if (d.dim == 1 && d.minisize[0] < d.n[0]) return deposit_1d_owner(g, n, centres, heights, bias_added, st);
  const long nsm8 = 8L * sm_count(g->device);
  int blocks = (int)(n < nsm8 ? n : nsm8);
  switch (d.dim) {
    case 1: deposit_hills_kernel<1><<<blocks, 256, 0, st>>>(d, n, centres, heights, bias_added, g->d_flags); break;
    case 2: deposit_hills_kernel<2><<<blocks, 256, 0, st>>>(d, n, centres, heights, bias_added, g->d_flags); break;
    default: deposit_hills_kernel<3><<<blocks, 256, 0, st>>>(d, n, centres, heights, bias_added, g->d_flags); break;
  }
  EDM_CUDA(cudaGetLastError());
  if (d.n_dup) dup_boundary_kernel<<<1, 64, 0, st>>>(d, g->d_flags, nullptr, 0);
  EDM_CUDA(cudaGetLastError());
  return EDM_OK;
}

int edm_gauss_deposit(edm_grid_t* g, long n, const double* centres, const double* heights, double* bias_added) {
  EDM_REQUIRE(g && g->d.is_gauss, "deposit needs a GaussGrid");
  EDM_TRY(ensure_device(g->device));
  if (n <= 0) return EDM_OK;
  int D = g->d.dim;
  size_t nc = (size_t)n * D;
  EDM_TRY(g->io.reserve((nc + 2 * (size_t)n) * sizeof(double)));
  double* dc = g->io.as<double>();
  double* dh = dc + nc;
  double* db = dh + n;
  EDM_CUDA(cudaMemcpy(dc, centres, nc * sizeof(double), cudaMemcpyHostToDevice));
  EDM_CUDA(cudaMemcpy(dh, heights, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  EDM_TRY(edm_gauss_deposit_dev(g, n, dc, dh, db, nullptr));
  EDM_CUDA(cudaDeviceSynchronize());
  if (bias_added) EDM_CUDA(cudaMemcpy(bias_added, db, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  return EDM_OK;
}

}  // extern "C"
