// EDM::EDMBias — the reference's per-timestep driver (lib/edm_bias.h:29-225) with its state in
// B200 HBM.  Same constructor (an edm input file), same call sequence (setup -> subdivide ->
// update_forces / add_hills or the pre/add/post triple), same public data members; every
// evaluation, selection, limiter and deposit step runs in the CUDA library behind
// include/edm_b200.h.  There is no MPI here: one process drives one GPU and the grid is replicated.
// With a communicator attached (set_comm) post_add_hill / add_hills exchange the accepted hills of all
// ranks through one NCCL all-gather inside the library — the stand-in for flush_buffers and
// update_height's reduction (lib/edm_bias.cpp:565-583, 614-706, 922-931) — and mpi_rank_/mpi_size_
// mirror the communicator.
#ifndef EDM_B200_EDM_BIAS_H_
#define EDM_B200_EDM_BIAS_H_

#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "edm.h"
#include "gaussian_grid.h"
#include "grid.h"

#define BIAS_CLAMP 1.0
#define BIAS_BUFFER_SIZE 2048
#define BIAS_BUFFER_DBLS 8192
#define NO_COMM_PARTNER -1
#define INTERPOLATE 1

#define NEIGH_HILL 'n'
#define BUFF_HILL 'b'
#define BUFF_UNDO_HILL 'v'
#define ADD_HILL 'h'
#define ADD_UNDO_HILL 'u'
#define BUFF_ZERO_HILL 'z'

namespace EDM {

class EDMBias {
 public:
  EDMBias(const std::string& input_filename);
  ~EDMBias();

  void subdivide(const double sublo[3], const double subhi[3], const double boxlo[3], const double boxhi[3],
                 const int b_periodic[3], const double skin[3]);
  void setup(double temperature, double boltzmann_constant);
  int read_input(const std::string& input_filename);

  double update_forces(int nlocal, const double* const* positions, double** forces, int apply_mask) const;
  double update_forces(int nlocal, const double* const* positions, double** forces) const;
  double update_force(const double* positions, double* forces) const;
  void set_mask(const int* mask);

  void add_hills(int nlocal, const double* const* positions, const double* runiform);
  void add_hills(int nlocal, const double* const* positions, const double* runiform, int apply_mask);
  void pre_add_hill(int est_hill_count);
  void add_hill(const double* position, double runiform);
  void post_add_hill();

  void write_bias(const std::string& output) const;
  void write_histogram() const;
  void clear_histogram();
  void write_lammps_table(const std::string& output) const;

  // B200 additions -----------------------------------------------------------------------------
  // fix edm_pair's whole post_force loop on the device (lammps/fix_edm_pair.cpp:177-240): positions
  // and forces are n x 3, box is a periodic orthorhombic cell; returns the bias energy and leaves the
  // number of hill proposals in *ncalls (the caller's next last_calls).
  double pair_step(long natoms, const double* x, double* f, const int* type, int itype, int jtype,
                   const double box[3], double cutoff, int do_hills, long long est_hill_count,
                   unsigned long long seed, unsigned long long step, long long* ncalls);
  // fix edm's whole post_force (lammps/fix_edm.cpp:134-162) as one call: update_forces and, if
  // do_hills, add_hills over the same atoms; the coordinates cross PCIe once and the atoms stream
  // through the GPU in chunks.  Same results as update_forces(...) followed by add_hills(...).
  double update_forces_add_hills(int nlocal, const double* const* positions, double** forces,
                                 const double* runiform, int apply_mask, int do_hills);
  edm_bias_t* device_bias() const { return dev_; }
  // Multi-GPU: one EDMBias per process/GPU, grid replicated.  After set_comm every hill round
  // (add_hills, post_add_hill, update_forces_add_hills, pair_step) gathers the accepted hills of all
  // ranks and commits them in rank-major order on every replica; est_hill_count stays this rank's
  // count, as in the reference's MPI build (lib/edm_bias.cpp:175-180).  May be called before or
  // after subdivide; the communicator is borrowed.
  void set_comm(edm_comm_t* comm, long block_capacity = 0);
  // after a hill round launched directly through the C ABI: HILLS lines, cum_bias_, host mirrors
  void after_device_round() {
    drain_hill_log();
    refresh_state();
  }

  // public data, lib/edm_bias.h:118-157
  int b_tempering_;
  int b_targeting_;
  int mpi_rank_;
  int mpi_size_;
  unsigned int dim_;
  double global_tempering_;
  double bias_factor_;
  double boltzmann_factor_;
  double temperature_;
  double hill_prefactor_;
  double bias_per_step_;
  double hill_density_;
  double cum_bias_;
  double total_volume_;
  double expected_target_;
  int b_outofbounds_;
  double* bias_dx_;
  double* bias_sigma_;
  double* min_;
  double* max_;
  int* b_periodic_boundary_;
  Grid* target_;
  Grid* initial_bias_;
  GaussGrid* bias_;
  const int* mask_;
  std::ofstream hill_output_;

 private:
  EDMBias(const EDMBias& that);
  void create_device_state();
  void drain_hill_log();   // device log -> HILLS lines, lib/edm_bias.cpp:586-599
  void refresh_state();    // cum_bias_ etc. from the device
  const double* pack_rows(int n, const double* const* rows, int width, std::vector<double>& scratch, long* stride) const;
  std::string clean_string(const std::string& input, int append_rank);

  edm_bias_t* dev_;
  edm_comm_t* comm_;
  long comm_cap_;
  Grid* cv_hist_;
  std::string hist_output_;
  int est_hill_count_;
  int in_round_;
  long long steps_;
  std::vector<double> pending_x_, pending_u_;  // add_hill calls of the open round, flushed in batches
  mutable std::vector<double> scratch_x_, scratch_f_;
};

}  // namespace EDM
#endif  // EDM_B200_EDM_BIAS_H_
