#ifdef FIX_CLASS

FixStyle(edm,FixEDM)

#else

#ifndef LMP_FIX_EDM_H
#define LMP_FIX_EDM_H

#include "fix.h"
#include <edm/edm_bias.h>

namespace LAMMPS_NS {

// fix ID group edm T input.edm hill_stride write_stride bias_file seed
// Coordinate-CV entry point (reference: lammps/fix_edm.{h,cpp}).  Each step the atoms' own
// coordinates are the CV: one batched force update on the GPU, and every `stride` steps one hill
// round (selection, limiter, deposit) on the GPU.
class FixEDM : public Fix {
 public:
  FixEDM(class LAMMPS*, int, char**);
  ~FixEDM();
  int setmask();
  void init();
  void setup(int);
  void min_setup(int);
  void post_force(int);
  void post_force_respa(int, int, int);
  void min_post_force(int);
  double compute_scalar();

 private:
  EDM::EDMBias* bias;
  class RanMars* random;
  double* random_numbers;
  int random_capacity;
  char bias_file[512];
  double temperature, edm_energy;
  int stride, write_stride;
  unsigned int seed;
  // atom->x / atom->f page-locked for the host<->device copies while EDM_B200_PIN is set; re-pinned when
  // LAMMPS reallocates them
  void pin_atom_arrays();
  void unpin_atom_arrays();
  // more than one MPI rank: one GPU per rank, hills exchanged through the library's NCCL all-gather
  void setup_exchange();
  edm_comm_t* comm_ = 0;
  void* pinned_x_ = 0;
  void* pinned_f_ = 0;
  long pinned_nmax_ = 0;
};

}  // namespace LAMMPS_NS

#endif
#endif
