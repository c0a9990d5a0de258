#ifdef FIX_CLASS

FixStyle(edm_pair,FixEDMPair)

#else

#ifndef LMP_FIX_EDM_PAIR_H
#define LMP_FIX_EDM_PAIR_H

#include "fix.h"
#include <edm/edm_bias.h>
#include "neigh_list.h"

#include <vector>

namespace LAMMPS_NS {

// fix ID group edm_pair T input.edm hill_stride write_stride bias_file seed itype jtype
// Pair-distance CV entry point (reference: lammps/fix_edm_pair.{h,cpp}).  The half neighbour list
// LAMMPS builds is flattened to CSR and the whole loop — r, bias energy and force at r, force
// scatter, two hill proposals per local pair — runs as one kernel (edm_pair_step_list).
class FixEDMPair : public Fix {
 public:
  FixEDMPair(class LAMMPS*, int, char**);
  ~FixEDMPair();
  int setmask();
  void init();
  void setup(int);
  void min_setup(int);
  void post_force(int);
  void post_force_respa(int, int, int);
  void min_post_force(int);
  void init_list(int, class NeighList*);
  double compute_scalar();

 private:
  EDM::EDMBias* bias;
  class NeighList* list;
  char bias_file[512];
  char lammps_table_file[520];
  double temperature, edm_energy;
  int stride, write_stride;
  unsigned int seed;
  long long last_calls;
  int ipair, jpair;
  std::vector<long> first_;
  std::vector<int> jlist_;
  bool list_on_device_ = false;  // the flattened list sits on the GPU until LAMMPS rebuilds it (neighbor->ago == 0)
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
