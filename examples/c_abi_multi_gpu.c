/* One process per GPU, plain C: the multi-rank recipe of INTEGRATION.md section 4, end to end.
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_multi_gpu.c -Lelectronic-dance-music_b200/lib -ledm_b200 \
 *       -Wl,-rpath,$PWD/electronic-dance-music_b200/lib -lm -o c_abi_multi_gpu
 *   for r in 0 1; do ./c_abi_multi_gpu $r 2 /tmp/edm_rendezvous & done; wait
 *
 * Every rank owns its own atoms (a 2-D coordinate CV with local well-tempering, hill_density 200), keeps a replica of
 * the bias grid, and calls the single-rank entry point edm_bias_step_coords on HOST buffers; because a communicator is
 * attached (edm_bias_set_comm), the library exchanges the accepted hills itself before the limiter -- over NVLink peer
 * windows on one node, ncclAllGather otherwise -- and every rank deposits the same hills.  What the reference does with
 * MPI_COMM_WORLD inside EDMBias (lib/edm_bias.cpp:565-583, 614-706).  The id of the communicator travels through a file
 * here (edm_comm_init_file); a LAMMPS build ships it with MPI_Bcast (lammps/fix_edm.cpp of this repo).
 *
 * Prints one line per rank: "rank R/N transport T steps S hills H cum_bias C checksum K" -- cum_bias and the checksum
 * of the bias over a fixed probe set must be IDENTICAL (every digit) on all ranks. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "edm_b200.h"

#define CHECK(call)                                                     \
  do {                                                                  \
    int rc_ = (call);                                                   \
    if (rc_ != EDM_OK) {                                                \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, edm_last_error()); \
      return 1;                                                         \
    }                                                                   \
  } while (0)

static unsigned long long rng_state;
static double uniform01(void) { /* splitmix64: any generator will do, the atoms are this rank's own */
  unsigned long long z = (rng_state += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char** argv) {
  const int rank = argc > 1 ? atoi(argv[1]) : 0;
  const int nranks = argc > 2 ? atoi(argv[2]) : 1;
  const char* rendezvous = argc > 3 ? argv[3] : "/tmp/edm_b200_rendezvous";
  int ndev = 0;
  CHECK(edm_device_count(&ndev));
  if (ndev == 0) {
    fprintf(stderr, "no CUDA device: this library has no CPU path\n");
    return 2;
  }
  const int device = rank % ndev;

  /* the bias of a 2-D coordinate CV on a periodic 16 x 16 box, as an edm file would set it up */
  const double mn[2] = {0.0, 0.0}, mx[2] = {16.0, 16.0}, dx[2] = {0.03125, 0.03125}, sigma[2] = {0.0625, 0.0625};
  const int per[2] = {1, 1};
  edm_grid_t *bias = NULL, *hist = NULL;
  CHECK(edm_gauss_create(&bias, device, 2, mn, mx, dx, per, 1, sigma));
  CHECK(edm_grid_set_boundary(bias, mn, mx, per));
  CHECK(edm_grid_create(&hist, device, 2, mn, mx, sigma, per, 0, 0));
  edm_bias_params_t prm;
  prm.dim = 2;
  prm.b_tempering = 1;
  prm.b_targeting = 0;
  prm.global_tempering = -1.0; /* local well-tempering: a hill's height reads the bias under it */
  prm.bias_factor = 5.0;
  prm.boltzmann_factor = 300.0 * 0.0019872;
  prm.hill_prefactor = 0.02;
  prm.bias_per_step = 1000.0;
  prm.hill_density = 200.0;
  prm.expected_target = 0.0;
  prm.total_volume = 256.0;
  edm_bias_t* b = NULL;
  CHECK(edm_bias_create(&b, bias, hist, NULL, &prm));

  /* the communicator: collective over all ranks; afterwards every whole-round call of `b` exchanges first */
  edm_comm_t* comm = NULL;
  CHECK(edm_comm_init_file(&comm, rendezvous, nranks, rank, device, 120.0));
  CHECK(edm_bias_set_comm(b, comm, 0));
  int peer = 0;
  CHECK(edm_comm_peer_windows(comm, &peer));

  const long n = 20000;
  const int steps = 5;
  double* x = (double*)malloc(sizeof(double) * 2 * n);
  double* f = (double*)calloc(2 * n, sizeof(double));
  rng_state = 1234 + 977ULL * (unsigned long long)rank;
  double energy = 0.0;
  for (int step = 0; step < steps; step++) {
    for (long i = 0; i < 2 * n; i++) x[i] = 16.0 * uniform01();
    /* fix edm's post_force: f -= dV/dx for this rank's atoms, then add_hills over them; the selection threshold uses
     * the job-wide atom count (n * nranks), the accepted hills of all ranks are committed in rank order */
    CHECK(edm_bias_step_coords(b, n, x, 2, f, 2, NULL, -1, 1, NULL, 42 + (unsigned long long)rank, step, &energy));
  }
  edm_bias_state_t st;
  CHECK(edm_bias_state(b, &st));
  /* checksum of the replica: the bias at a fixed probe set (the same on every rank) */
  enum { NP = 4096 };
  static double probe[2 * NP], val[NP], der[2 * NP];
  rng_state = 99;
  for (int i = 0; i < 2 * NP; i++) probe[i] = 16.0 * uniform01();
  CHECK(edm_grid_eval(bias, NP, probe, 2, val, der));
  double checksum = 0.0;
  for (int i = 0; i < NP; i++) checksum += val[i] * (1.0 + (double)(i % 7)) + der[2 * i] - der[2 * i + 1];
  printf("rank %d/%d transport %s steps %lld hills %d cum_bias %.17g checksum %.17g\n", rank, nranks,
         nranks == 1 ? "none" : (peer ? "nvlink-peer-windows" : "nccl-allgather"), st.steps, st.hills_added, st.cum_bias,
         checksum);
  free(x);
  free(f);
  CHECK(edm_bias_set_comm(b, NULL, 0));
  CHECK(edm_comm_destroy(comm));
  CHECK(edm_bias_destroy(b));
  CHECK(edm_grid_destroy(hist));
  CHECK(edm_grid_destroy(bias));
  return st.hills_added > 0 ? 0 : 1;
}
