// fix edm_pair — pair-distance collective variable, B200 build.
// Reference entry point: lammps/fix_edm_pair.cpp:139-256 (post_force).  Differences, both stated in
// (1) all pairs are evaluated against the start-of-step bias and the hills are deposited
// afterwards (the reference interleaves them pair by pair); (2) uniforms come from the device's
// counter-based generator keyed by (seed + rank, timestep, position in the list) instead of RanMars.
#include "fix_edm_pair.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "atom.h"
#include "error.h"
#include "force.h"
#include "neigh_request.h"
#include "neighbor.h"
#include "pair.h"
#include "respa.h"
#include "update.h"

using namespace LAMMPS_NS;
using namespace FixConst;

FixEDMPair::FixEDMPair(LAMMPS* lmp, int narg, char** arg)
    : Fix(lmp, narg, arg), bias(NULL), list(NULL), edm_energy(0), last_calls(0) {
  if (narg < 11) error->all(FLERR, "Illegal fix edm_pair command");
  if (!atom->tag_enable) error->all(FLERR, "fix edm_pair requires atom tags");
  int me = 0;
  MPI_Comm_rank(world, &me);
  temperature = atof(arg[3]);
  stride = atoi(arg[5]);
  write_stride = atoi(arg[6]);
  strncpy(bias_file, arg[7], sizeof(bias_file) - 1);
  bias_file[sizeof(bias_file) - 1] = '\0';
  snprintf(lammps_table_file, sizeof(lammps_table_file), "%s.ltab", bias_file);
  seed = (unsigned int)atoi(arg[8]) + (unsigned int)me;
  if (stride < 0) error->all(FLERR, "Illegal stride given to edm_pair command");
  if (write_stride < 0) error->all(FLERR, "Illegal write bias stride given to edm_pair command");
  ipair = atoi(arg[9]);
  jpair = atoi(arg[10]);
  if (!ipair || !jpair) error->all(FLERR, "Illegeal EDM command, invalid types");
  thermo_energy = 1;
  int ndev = 0;
  if (!getenv("EDM_B200_DEVICE") && edm_device_count(&ndev) == EDM_OK && ndev > 0)
    EDM::set_default_device(me % ndev);  // one GPU per MPI rank of the node
  bias = new EDM::EDMBias(arg[4]);
  if (bias->dim_ != 1) error->all(FLERR, "Pairwise distance must be 1 dimension in EDM input file");
}

FixEDMPair::~FixEDMPair() {
  unpin_atom_arrays();
  delete bias;
  if (comm_) edm_comm_destroy(comm_);
}

void FixEDMPair::pin_atom_arrays() {
  if (!getenv("EDM_B200_PIN")) return;
  if (pinned_x_ == (void*)atom->x[0] && pinned_f_ == (void*)atom->f[0] && pinned_nmax_ == atom->nmax) return;
  unpin_atom_arrays();
  const size_t bytes = (size_t)atom->nmax * 3 * sizeof(double);
  if (edm_host_pin(atom->x[0], bytes) == EDM_OK) pinned_x_ = atom->x[0];
  if (edm_host_pin(atom->f[0], bytes) == EDM_OK) pinned_f_ = atom->f[0];
  pinned_nmax_ = atom->nmax;
}

void FixEDMPair::unpin_atom_arrays() {
  if (pinned_x_) edm_host_unpin(pinned_x_);
  if (pinned_f_) edm_host_unpin(pinned_f_);
  pinned_x_ = pinned_f_ = 0;
  pinned_nmax_ = 0;
}

// The reference's MPI build broadcasts every rank's hills inside post_add_hill (lib/edm_bias.cpp:565-583,
// 614-706).  Here each rank drives one GPU with a replica of the whole grid; the ranks agree on an NCCL
// communicator once (the 128-byte id travels over LAMMPS' own MPI world) and EDMBias then all-gathers the
// accepted hills inside every hill round, so all replicas deposit the same hills in the same order.
void FixEDMPair::setup_exchange() {
  int me = 0, nprocs = 1;
  MPI_Comm_rank(world, &me);
  MPI_Comm_size(world, &nprocs);
  if (nprocs <= 1 || comm_) return;
  unsigned char id[EDM_COMM_ID_BYTES];
  memset(id, 0, sizeof(id));
  if (me == 0) EDM::edm_check(edm_comm_unique_id(id), "fix_edm_pair.cpp:setup_exchange");
  MPI_Bcast(id, EDM_COMM_ID_BYTES, MPI_BYTE, 0, world);
  EDM::edm_check(edm_comm_init_rank(&comm_, id, nprocs, me, EDM::default_device()), "fix_edm_pair.cpp:setup_exchange");
  bias->set_comm(comm_);
}

int FixEDMPair::setmask() { return POST_FORCE | THERMO_ENERGY | POST_FORCE_RESPA | MIN_POST_FORCE; }

void FixEDMPair::init() {
  if (strcmp(update->integrate_style, "respa") == 0) nlevels_respa = ((Respa*)update->integrate)->nlevels;
  bias->setup(temperature, force->boltz);
  setup_exchange();
  // every rank biases the same whole-range 1-D grid [-skin, cut + 2 skin] (fix_edm_pair.cpp:96-104)
  double skin[3] = {neighbor->skin, 0, 0};
  double lo[3] = {0, 0, 0}, hi[3] = {force->pair->cutforce + neighbor->skin, 0, 0};
  int p[3] = {0, 0, 0};
  bias->subdivide(lo, hi, lo, hi, p, skin);
  last_calls = atom->nmax;
  int irequest = neighbor->request((void*)this);
  neighbor->requests[irequest]->pair = 0;
  neighbor->requests[irequest]->fix = 1;
  edm_energy = 0;
}

void FixEDMPair::init_list(int, NeighList* ptr) { list = ptr; }

void FixEDMPair::setup(int vflag) {
  if (strcmp(update->integrate_style, "verlet") == 0) {
    post_force(vflag);
  } else {
    ((Respa*)update->integrate)->copy_flevel_f(nlevels_respa - 1);
    post_force_respa(vflag, nlevels_respa - 1, 0);
    ((Respa*)update->integrate)->copy_f_flevel(nlevels_respa - 1);
  }
}

void FixEDMPair::min_setup(int vflag) { post_force(vflag); }

void FixEDMPair::post_force(int) {
  if (force->newton_pair)
    error->all(FLERR, "fix edm_pair requires 'newton off' to be declared in the lammps input script");
  pin_atom_arrays();
  const int inum = list->inum;
  // NeighList -> CSR (the NEIGHMASK strip of fix_edm_pair.cpp:196 happens here), uploaded only on the steps
  // LAMMPS rebuilt the list; in between the device keeps it and only positions and forces travel
  if (neighbor->ago == 0 || !list_on_device_) {
    first_.resize((size_t)inum + 1);
    first_[0] = 0;
    for (int ii = 0; ii < inum; ii++) first_[ii + 1] = first_[ii] + list->numneigh[list->ilist[ii]];
    jlist_.resize((size_t)first_[inum]);
    for (int ii = 0; ii < inum; ii++) {
      const int i = list->ilist[ii];
      const int* src = list->firstneigh[i];
      int* dst = &jlist_[(size_t)first_[ii]];
      for (int jj = 0; jj < list->numneigh[i]; jj++) dst[jj] = src[jj] & NEIGHMASK;
    }
    EDM::edm_check(edm_pair_list_set(bias->device_bias(), inum, list->ilist, first_.data(), jlist_.data()),
                   "fix_edm_pair.cpp:post_force");
    list_on_device_ = true;
  }
  const bool hills = stride > 0 && update->ntimestep % stride == 0;
  const long nall = atom->nlocal + atom->nghost;
  edm_pair_result_t res;
  EDM::edm_check(edm_pair_step_listed(bias->device_bias(), nall, atom->nlocal, atom->x[0], atom->f[0], atom->type, ipair,
                                      jpair, hills ? 1 : 0, last_calls, NULL, seed, (uint64_t)update->ntimestep, &res),
                 "fix_edm_pair.cpp:post_force");
  edm_energy = res.energy;
  if (hills) {
    last_calls = res.n_calls;  // fix_edm_pair.cpp:245
    bias->after_device_round();
  }
  if (write_stride > 0 && update->ntimestep % write_stride == 0) {
    bias->write_bias(bias_file);
    bias->write_lammps_table(lammps_table_file);
    bias->write_histogram();
    bias->clear_histogram();
  }
}

void FixEDMPair::post_force_respa(int vflag, int ilevel, int) {
  if (ilevel == nlevels_respa - 1) post_force(vflag);
}

void FixEDMPair::min_post_force(int vflag) { post_force(vflag); }

double FixEDMPair::compute_scalar() { return edm_energy; }
