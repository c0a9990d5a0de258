// fix edm — coordinate collective variables, B200 build.
// Reference entry point: lammps/fix_edm.cpp:134-162 (post_force).  The per-step work is one call
// into EDM::EDMBias: the force update and the hill round share one pipelined pass over the atoms.
#include "fix_edm.h"

#include <cstdlib>
#include <cstring>

#include "atom.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "memory.h"
#include "neighbor.h"
#include "random_mars.h"
#include "respa.h"
#include "update.h"

using namespace LAMMPS_NS;
using namespace FixConst;

FixEDM::FixEDM(LAMMPS* lmp, int narg, char** arg)
    : Fix(lmp, narg, arg), bias(NULL), random(NULL), random_numbers(NULL), random_capacity(0), edm_energy(0) {
  if (narg < 9) error->all(FLERR, "Illegal fix EDM command");
  if (!atom->tag_enable) error->all(FLERR, "fix EDM requires atom tags");
  int me = 0;
  MPI_Comm_rank(world, &me);
  temperature = atof(arg[3]);
  stride = atoi(arg[5]);
  write_stride = atoi(arg[6]);
  strncpy(bias_file, arg[7], sizeof(bias_file) - 1);
  bias_file[sizeof(bias_file) - 1] = '\0';
  seed = atoi(arg[8]);
  if (stride < 0) error->all(FLERR, "Illegal stride given to EDM command");
  if (write_stride < 0) error->all(FLERR, "Illegal write bias stride given to EDM command");
  int ndev = 0;
  if (!getenv("EDM_B200_DEVICE") && edm_device_count(&ndev) == EDM_OK && ndev > 0)
    EDM::set_default_device(me % ndev);  // one GPU per MPI rank of the node
  bias = new EDM::EDMBias(arg[4]);
  thermo_energy = 1;
  random = new RanMars(lmp, seed + me);
}

FixEDM::~FixEDM() {
  unpin_atom_arrays();
  delete bias;
  if (comm_) edm_comm_destroy(comm_);
  delete random;
  free(random_numbers);
}

void FixEDM::pin_atom_arrays() {
  if (!getenv("EDM_B200_PIN")) return;
  if (pinned_x_ == (void*)atom->x[0] && pinned_f_ == (void*)atom->f[0] && pinned_nmax_ == atom->nmax) return;
  unpin_atom_arrays();
  const size_t bytes = (size_t)atom->nmax * 3 * sizeof(double);
  if (edm_host_pin(atom->x[0], bytes) == EDM_OK) pinned_x_ = atom->x[0];
  if (edm_host_pin(atom->f[0], bytes) == EDM_OK) pinned_f_ = atom->f[0];
  pinned_nmax_ = atom->nmax;
}

void FixEDM::unpin_atom_arrays() {
  if (pinned_x_) edm_host_unpin(pinned_x_);
  if (pinned_f_) edm_host_unpin(pinned_f_);
  pinned_x_ = pinned_f_ = 0;
  pinned_nmax_ = 0;
}

// The reference's MPI build broadcasts every rank's hills inside post_add_hill (lib/edm_bias.cpp:565-583,
// 614-706).  Here each rank drives one GPU with a replica of the whole grid; the ranks agree on an NCCL
// communicator once (the 128-byte id travels over LAMMPS' own MPI world) and EDMBias then all-gathers the
// accepted hills inside every hill round, so all replicas deposit the same hills in the same order.
void FixEDM::setup_exchange() {
  int me = 0, nprocs = 1;
  MPI_Comm_rank(world, &me);
  MPI_Comm_size(world, &nprocs);
  if (nprocs <= 1 || comm_) return;
  unsigned char id[EDM_COMM_ID_BYTES];
  memset(id, 0, sizeof(id));
  if (me == 0) EDM::edm_check(edm_comm_unique_id(id), "fix_edm.cpp:setup_exchange");
  MPI_Bcast(id, EDM_COMM_ID_BYTES, MPI_BYTE, 0, world);
  EDM::edm_check(edm_comm_init_rank(&comm_, id, nprocs, me, EDM::default_device()), "fix_edm.cpp:setup_exchange");
  bias->set_comm(comm_);
}

int FixEDM::setmask() { return POST_FORCE | THERMO_ENERGY | POST_FORCE_RESPA | MIN_POST_FORCE; }

void FixEDM::init() {
  if (strcmp(update->integrate_style, "respa") == 0) nlevels_respa = ((Respa*)update->integrate)->nlevels;
  bias->setup(temperature, force->boltz);
  setup_exchange();
  double skin[3] = {neighbor->skin, neighbor->skin, neighbor->skin};
  // the bias grid is replicated on every GPU: the "sub-box" handed over is the whole box
  bias->subdivide(domain->boxlo, domain->boxhi, domain->boxlo, domain->boxhi, domain->periodicity, skin);
  bias->set_mask(atom->mask);
  edm_energy = 0;
}

void FixEDM::setup(int vflag) {
  if (strcmp(update->integrate_style, "verlet") == 0) {
    post_force(vflag);
  } else {
    ((Respa*)update->integrate)->copy_flevel_f(nlevels_respa - 1);
    post_force_respa(vflag, nlevels_respa - 1, 0);
    ((Respa*)update->integrate)->copy_f_flevel(nlevels_respa - 1);
  }
}

void FixEDM::min_setup(int vflag) { post_force(vflag); }

void FixEDM::post_force(int) {
  const int n = atom->nlocal;
  pin_atom_arrays();
  bias->set_mask(atom->mask);
  const int do_hills = (stride > 0 && update->ntimestep % stride == 0) ? 1 : 0;
  if (do_hills) {
    if (random_capacity < n) {  // one uniform per local atom, drawn in atom order (fix_edm.cpp:149-151)
      random_numbers = (double*)realloc(random_numbers, sizeof(double) * (size_t)atom->nmax);
      random_capacity = atom->nmax;
    }
    for (int i = 0; i < n; i++) random_numbers[i] = random->uniform();
  }
  // update_forces, then add_hills on the same coordinates (fix_edm.cpp:140, 153): one upload, one
  // pipelined pass over the atoms
  edm_energy = bias->update_forces_add_hills(n, atom->x, atom->f, random_numbers, groupbit, do_hills);
  if (write_stride > 0 && update->ntimestep % write_stride == 0) {
    bias->write_bias(bias_file);
    bias->write_histogram();
    bias->clear_histogram();
  }
}

void FixEDM::post_force_respa(int vflag, int ilevel, int) {
  if (ilevel == nlevels_respa - 1) post_force(vflag);
}

void FixEDM::min_post_force(int vflag) { post_force(vflag); }

double FixEDM::compute_scalar() { return edm_energy; }
