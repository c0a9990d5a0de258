// Minimal stand-in for the LAMMPS headers the EDM fixes include (fix.h, atom.h, update.h, force.h,
// neighbor.h, neigh_list.h, neigh_request.h, domain.h, random_mars.h, respa.h, error.h, memory.h,
// pair.h, group.h, lmptype.h).  LAMMPS is not in this image; this header exists so that the fixes can
// be COMPILED (and driven by a tiny fake MD loop in the tests).  It declares only what the fixes use.
#ifndef EDM_B200_LAMMPS_MOCK_H
#define EDM_B200_LAMMPS_MOCK_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#ifndef MPI_VERSION
typedef int MPI_Comm;
inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return 0; }
inline int MPI_Comm_size(MPI_Comm, int* s) { *s = 1; return 0; }
typedef int MPI_Datatype;
#define MPI_BYTE 1
inline int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return 0; }  // one rank: nothing to send
#endif

#define FLERR __FILE__, __LINE__
#define NEIGHMASK 0x3FFFFFFF

namespace LAMMPS_NS {

typedef long long bigint;

namespace FixConst {
enum { POST_FORCE = 1 << 0, THERMO_ENERGY = 1 << 1, POST_FORCE_RESPA = 1 << 2, MIN_POST_FORCE = 1 << 3 };
}

class Error {
 public:
  void all(const char* file, int line, const char* msg) {
    fprintf(stderr, "ERROR: %s (%s:%d)\n", msg, file, line);
    exit(1);
  }
};

class Memory {
 public:
  template <typename T> T* create(T*& p, int n, const char*) {
    p = (T*)malloc(sizeof(T) * (size_t)(n > 0 ? n : 1));
    return p;
  }
  template <typename T> void destroy(T*& p) {
    free(p);
    p = NULL;
  }
};

class Atom {
 public:
  int nlocal, nghost, nmax, tag_enable;
  double** x;
  double** f;
  int* type;
  int* mask;
};

class Integrate {
 public:
  virtual ~Integrate() {}
};
class Respa : public Integrate {
 public:
  int nlevels;
  void copy_flevel_f(int) {}
  void copy_f_flevel(int) {}
};

class Update {
 public:
  bigint ntimestep;
  const char* integrate_style;
  Integrate* integrate;
};

class Pair {
 public:
  double cutforce;
};

class Force {
 public:
  double boltz;
  int newton_pair;
  Pair* pair;
};

class NeighRequest {
 public:
  int pair, fix, half, full;
};

class NeighList {
 public:
  int inum;
  int* ilist;
  int* numneigh;
  int** firstneigh;
};

class Neighbor {
 public:
  double skin;
  int ago;  // steps since the lists were last rebuilt (neighbor.h of LAMMPS)
  NeighRequest** requests;
  int nrequest;
  Neighbor() : skin(0), ago(0), requests(NULL), nrequest(0) {}
  int request(void*) {
    requests = (NeighRequest**)realloc(requests, sizeof(NeighRequest*) * (size_t)(nrequest + 1));
    requests[nrequest] = new NeighRequest();
    return nrequest++;
  }
};

class Domain {
 public:
  double boxlo[3], boxhi[3], sublo[3], subhi[3], prd[3];
  int periodicity[3];
};

class LAMMPS {
 public:
  Atom* atom;
  Update* update;
  Force* force;
  Neighbor* neighbor;
  Domain* domain;
  Error* error;
  Memory* memory;
  MPI_Comm world;
};

class Pointers {
 public:
  explicit Pointers(LAMMPS* l)
      : lmp(l), atom(l->atom), update(l->update), force(l->force), neighbor(l->neighbor), domain(l->domain),
        error(l->error), memory(l->memory), world(l->world) {}
  virtual ~Pointers() {}

 protected:
  LAMMPS* lmp;
  Atom*& atom;
  Update*& update;
  Force*& force;
  Neighbor*& neighbor;
  Domain*& domain;
  Error*& error;
  Memory*& memory;
  MPI_Comm& world;
};

class RanMars : protected Pointers {  // deterministic stand-in, NOT Marsaglia's generator
 public:
  RanMars(LAMMPS* l, int seed) : Pointers(l), state_(0x9E3779B97F4A7C15ULL ^ (unsigned long long)seed) {}
  double uniform() {
    state_ += 0x9E3779B97F4A7C15ULL;
    unsigned long long z = state_;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
  }

 private:
  unsigned long long state_;
};

class Fix : protected Pointers {
 public:
  Fix(LAMMPS* l, int, char**) : Pointers(l), groupbit(1), thermo_energy(0), nlevels_respa(0) {}
  virtual ~Fix() {}
  virtual int setmask() = 0;
  virtual void init() {}
  virtual void setup(int) {}
  virtual void min_setup(int) {}
  virtual void post_force(int) {}
  virtual void post_force_respa(int, int, int) {}
  virtual void min_post_force(int) {}
  virtual void init_list(int, NeighList*) {}
  virtual double compute_scalar() { return 0.0; }
  int groupbit;
  int thermo_energy;

 protected:
  int nlevels_respa;
};

}  // namespace LAMMPS_NS
#endif
