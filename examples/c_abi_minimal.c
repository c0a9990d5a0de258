/* Minimal C host for the B200 bias engine: what a non-C++ caller (cgo, JNI, Fortran bind(C)) does.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_minimal.c -Lelectronic-dance-music_b200/lib -ledm_b200 \
 *       -Wl,-rpath,$PWD/electronic-dance-music_b200/lib -lm -o c_abi_minimal && ./c_abi_minimal
 *
 * One hill at 0.25 on the notebook's grid (python-example/input.edm of the reference), then the bias energy and
 * derivative at 0.24: python-example/EDM.ipynb:103 gives (1.1002417338159258, -0.6144025830861709). */
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
  /* EDMBias::subdivide on box [0,10], bias box [0,1] non-periodic: grid [0,10], boundary [0,1] */
  const double mn[1] = {0.0}, mx[1] = {10.0}, dx[1] = {0.01}, sigma[1] = {0.5};
  const double bmin[1] = {0.0}, bmax[1] = {1.0};
  const int per[1] = {0};
  edm_grid_t* bias = NULL;
  CHECK(edm_gauss_create(&bias, 0, 1, mn, mx, dx, per, 1, sigma));
  CHECK(edm_grid_set_boundary(bias, bmin, bmax, per));
  const double centre[1] = {0.25}, height[1] = {1.0};
  double added = 0.0;
  CHECK(edm_gauss_deposit(bias, 1, centre, height, &added));
  const double x[1] = {0.24};
  double v = 0.0, dv = 0.0;
  CHECK(edm_grid_eval(bias, 1, x, 1, &v, &dv));
  printf("bias_added %.17g  V(0.24) %.17g  dV/dx %.17g\n", added, v, dv);
  const int ok = fabs(v - 1.1002417338159258) <= 1e-10 * 1.1002417338159258 &&
                 fabs(dv + 0.6144025830861709) <= 1e-10 * 0.6144025830861709;
  CHECK(edm_grid_destroy(bias));
  puts(ok ? "matches the reference's notebook vector" : "MISMATCH");
  return ok ? 0 : 1;
}
