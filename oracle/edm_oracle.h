/* CPU restatement of the EDM per-timestep bias engine.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain C11 restatement of the reference algorithm (whitead/electronic-dance-music, lib/ and
 * lammps/fix_edm_pair.cpp) for the hot path named in BASELINE.json.  It exists to CHECK the
 * CUDA path; nothing under electronic-dance-music_b200/ may include, link or load it.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this port (a) bit-for-bit against the
 * unmodified reference compiled into oracle/_ref/libedm_ref.so, (b) against the reference's own
 * known answers (tests/edm_test.cpp:123,445,886; python-example/EDM.ipynb:103) and (c) against
 * fixtures under tests/golden/ generated from the compiled reference.
 *
 * Built -O2 -ffp-contract=off so every expression keeps the reference's
 * left-to-right IEEE evaluation.  Each function cites the reference lines it follows.
 */
#ifndef EDM_ORACLE_H
#define EDM_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_GAUSS_SUPPORT 8.0    /* lib/gaussian_grid.h:10 */
#define ORC_BC_TABLE_SIZE 65536  /* lib/gaussian_grid.h:11 */
#define ORC_BC_MAR 2.0           /* lib/gaussian_grid.h:12 */
#define ORC_BIAS_CLAMP 1.0       /* lib/edm_bias.h:14 */
#define ORC_BUFFER_SLOTS 2048    /* lib/edm_bias.h:15 */
#define ORC_BUFFER_DBLS 8192     /* lib/edm_bias.h:16 */

typedef struct orc_grid {  /* lib/grid.h:876-885 */
  int dim;
  size_t size;
  int b_derivatives, b_interpolate;
  double* grid;
  double* deriv;
  double dx[3], min[3], max[3];
  int n[3];
  int periodic[3];
} orc_grid;

typedef struct orc_gauss {  /* lib/gaussian_grid.h:544-555 */
  orc_grid g;
  double sigma[3];
  double bmin[3], bmax[3];
  int bper[3];
  size_t minisize[3];
  size_t minisize_total;
  double* bc_denom[3];
  double* bc_deriv[3];
  int dirty;
} orc_gauss;

typedef struct orc_hill_event {  /* one HILLS line, lib/edm_bias.cpp:586-599 */
  long long steps;
  int type; /* 'h','u','b','v' */
  int hills_added;
  double pos[3];
  double height;
  double bias_added;
  double cum_over_vol;
} orc_hill_event;

typedef struct orc_bias orc_bias;

/* ---- plain grid ---- */
void* orc_grid_create(int dim, const double* mn, const double* mx, const double* spacing,
                      const int* periodic, int b_deriv, int b_interp);
void orc_grid_destroy(void* g);
void orc_grid_info(void* g, int* n, double* dx, double* mn, double* mx, int* flags);
size_t orc_grid_size(void* g);
void orc_grid_get_arrays(void* g, double* v, double* d);
void orc_grid_set_arrays(void* g, double* v, double* d);
void orc_grid_set_interpolation(void* g, int b);
void orc_grid_eval(void* g, long n, const double* x, double* val, double* der);
void orc_grid_get_value(void* g, long n, const double* x, double* val);
void orc_grid_hist_add(void* g, long n, const double* x, const double* v);
double orc_grid_expected_bias(void* g);

/* ---- gaussian grid ---- */
void* orc_gauss_create(int dim, const double* mn, const double* mx, const double* spacing,
                       const int* periodic, int interp, const double* sigma);
void orc_gauss_destroy(void* g);
void orc_gauss_set_boundary(void* g, const double* mn, const double* mx, const int* periodic);
void orc_gauss_info(void* g, int* n, double* dx, double* mn, double* mx, int* mini);
size_t orc_gauss_size(void* g);
void orc_gauss_get_arrays(void* g, double* v, double* d);
void orc_gauss_set_arrays(void* g, double* v, double* d);
void orc_gauss_tables(void* g, int dimi, double* denom, double* deriv);
double orc_gauss_add_value(void* g, const double* x, double height);
void orc_gauss_add_values(void* g, long n, const double* x, const double* height, double* bias_added);
void orc_gauss_eval(void* g, long n, const double* x, double* val, double* der);
void orc_gauss_get_value(void* g, long n, const double* x, double* val);
void orc_gauss_remap(void* g, double* x);
void orc_gauss_set_interpolation(void* g, int b);

/* ---- EDMBias step logic ---- */
void* orc_bias_create(int dim, int b_tempering, double global_tempering, double bias_factor,
                      double hill_prefactor, double bias_per_step, double hill_density,
                      const double* dx, const double* sigma, const double* mn, const double* mx);
void orc_bias_destroy(void* b);
void orc_bias_set_target(void* b, void* target_grid, double expected_target);
void orc_bias_setup(void* b, double temperature, double boltz);
void orc_bias_subdivide(void* b, const double* sublo, const double* subhi, const double* boxlo,
                        const double* boxhi, const int* periodic, const double* skin);
void* orc_bias_gauss(void* b);
void* orc_bias_hist(void* b);
void orc_bias_params(void* b, double* out14);
void orc_bias_set_cum_bias(void* b, double v);
void orc_bias_backlog(void* b, long* left, long* right, double* buffer);
void orc_bias_set_mask(void* b, const int* mask);
double orc_bias_update_forces(void* b, long n, const double* x, long xstride, double* f, long fstride,
                              int apply_mask);
void orc_bias_add_hills(void* b, long n, const double* x, long xstride, const double* runiform, int apply_mask);
void orc_bias_pre_add_hill(void* b, int est);
void orc_bias_add_hill_many(void* b, long n, const double* x, const double* runiform);
void orc_bias_post_add_hill(void* b);
long orc_bias_log_size(void* b);
void orc_bias_log_copy(void* b, orc_hill_event* out);
void orc_bias_log_clear(void* b);
void orc_bias_log_enable(void* b, int on);

/* ---- pair-distance CV driver (lammps/fix_edm_pair.cpp:177-240, lib-level order) ---- */
double orc_pair_step(void* b, long npairs, const int* pi, const int* pj, const double* x, double* f,
                     const double* shift, int do_hills, int est, const double* uniforms, double* r_out);
/* Half neighbour list (i < j, sorted by i then j) of all pairs with minimum-image distance
 * < cutoff in a periodic orthorhombic box [0,L)^3; stands in for the LAMMPS NeighList the
 * reference consumes.  Returns the pair count; fills pi/pj/shift up to cap entries. */
long orc_build_half_list(long natoms, const double* x, const double* box, double cutoff, long cap, int* pi,
                         int* pj, double* shift);

/* Counter-based uniform in [0,1): the stand-in for LAMMPS RanMars on both the CUDA path and
 * this oracle (the RNG is outside the parity boundary: add_hill takes runiform as an argument,
 * lib/edm_bias.cpp:528).  Same function as edm_uniform() in electronic-dance-music_b200/csrc. */
double orc_uniform(unsigned long long seed, unsigned long long step, unsigned long long counter);
double orc_uniform_pair(unsigned long long seed, unsigned long long step, unsigned long long pairkey, int which);
void orc_uniform_pair_fill(unsigned long long seed, unsigned long long step, long n,
                           const unsigned long long* pairkeys, double* out);
void orc_uniform_fill(unsigned long long seed, unsigned long long step, unsigned long long first, long n,
                      double* out);

/* timing helpers for bench.py's cpu_baseline (loops timed in C, marshalling excluded) */
double orc_time_pair_eval(void* b, long npairs, const double* r, int repeats);
double orc_time_add_values(void* g, long n, const double* x, const double* h);

#ifdef __cplusplus
}
#endif
#endif
