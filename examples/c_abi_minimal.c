/* Minimal C host for the B200 bias engine: what a non-C++ caller (cgo, JNI, Fortran bind(C)) does.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_minimal.c -Lelectronic-dance-music_b200/lib -ledm_b200 \
 *       -Wl,-rpath,$PWD/electronic-dance-music_b200/lib -lm -o c_abi_minimal && ./c_abi_minimal
 *
 * One add_hill at 0.25 with the notebook's input (python-example/input.edm of the reference), then the bias energy
 * and derivative at 0.24: python-example/EDM.ipynb:103 gives (1.1002417338159258, -0.6144025830861709). */
#include <math.h>
#include <stdio.h>

#include "edm_b200.h"

#define CHECK(call)                                                     \
  do {                                                                  \
    int rc_ = (call);                                                   \
    if (rc_ != EDM_OK) {                                                \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, edm_last_error()); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

int main(void) {
  int ndev = 0;
  CHECK(edm_device_count(&ndev));
  if (ndev == 0) {
    fprintf(stderr, "no CUDA device: this library has no CPU path\n");
    return 2;
  }
  /* what EDMBias(input.edm), setup(1, 1), subdivide([0],[10]) build: a bias grid over [0,10] with walls at
   * [0,1] (McGovern-De Pablo hills), a CV histogram with spacing sigma, and the step parameters */
  const double mn[1] = {0.0}, mx[1] = {10.0}, dx[1] = {0.01}, sigma[1] = {0.5};
  const double bmin[1] = {0.0}, bmax[1] = {1.0};
  const int per[1] = {0};
  edm_grid_t *bias = NULL, *hist = NULL;
  CHECK(edm_gauss_create(&bias, 0, 1, mn, mx, dx, per, 1, sigma));
  CHECK(edm_grid_set_boundary(bias, bmin, bmax, per));
  CHECK(edm_grid_create(&hist, 0, 1, mn, mx, sigma, per, 0, 0));
  edm_bias_params_t prm;
  prm.dim = 1;
  prm.b_tempering = 0;
  prm.b_targeting = 0;
  prm.global_tempering = 0.0;
  prm.bias_factor = 0.0;
  prm.boltzmann_factor = 1.0;
  prm.hill_prefactor = 1.0;
  prm.bias_per_step = 1.0;   /* defaults to hill_prefactor, lib/edm_bias.cpp:1030-1033 */
  prm.hill_density = -1.0;   /* absent: every candidate deposits */
  prm.expected_target = 0.0;
  prm.total_volume = 1.0;
  edm_bias_t* b = NULL;
  CHECK(edm_bias_create(&b, bias, hist, NULL, &prm));
  /* EDMBias.add_hill([0.25]) of the notebook = pre_add_hill(1); add_hill(x, u); post_add_hill().  The hill
   * integrates to 1.0101 > bias_per_step, so the limiter takes the excess back with an undo hill. */
  const double centre[1] = {0.25}, u[1] = {0.0};
  CHECK(edm_bias_pre_add_hill(b, 1));
  CHECK(edm_bias_add_hill_batch(b, 1, centre, u));
  CHECK(edm_bias_post_add_hill(b));
  edm_bias_state_t st;
  CHECK(edm_bias_state(b, &st));
  const double x[1] = {0.24};
  double v = 0.0, dv = 0.0;
  CHECK(edm_grid_eval(bias, 1, x, 1, &v, &dv));
  printf("hills deposited %d  cum_bias %.17g  V(0.24) %.17g  dV/dx %.17g\n", st.hills_added, st.cum_bias, v, dv);
  const int ok = fabs(v - 1.1002417338159258) <= 1e-10 * 1.1002417338159258 &&
                 fabs(dv + 0.6144025830861709) <= 1e-10 * 0.6144025830861709;
  CHECK(edm_bias_destroy(b));
  CHECK(edm_grid_destroy(hist));
  CHECK(edm_grid_destroy(bias));
  puts(ok ? "matches the reference's notebook vector" : "MISMATCH");
  return ok ? 0 : 1;
}
