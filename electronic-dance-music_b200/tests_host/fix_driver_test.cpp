// A tiny fake MD loop around the two device-native fixes (lammps/fix_edm.cpp, lammps/fix_edm_pair.cpp),
// built against the mock LAMMPS headers: the fixes are constructed from a fix command line, initialised and
// stepped the way LAMMPS would, and every step's energy and forces are checked against a recomputation that
// shares none of their kernels — the batched grid evaluation (edm_grid_eval) of the bias as it stood before
// the step, accumulated on the CPU.  Exercises: CSR flattening of the NeighList, the list kept on the device
// between rebuilds (neighbor->ago), ghost partners, group masks, hill rounds every `stride` steps, writers.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#define private public  // the test reads the fixes' bias
#include "../lammps/fix_edm.h"
#include "../lammps/fix_edm_pair.h"
#undef private

using namespace LAMMPS_NS;

static int g_failed = 0, g_checks = 0;
#define REQUIRE(c)                                                   \
  do {                                                               \
    g_checks++;                                                      \
    if (!(c)) {                                                      \
      g_failed++;                                                    \
      printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c);          \
    }                                                                \
  } while (0)

struct FakeLammps {
  LAMMPS lmp;
  Atom atom;
  Update update;
  Force force;
  Pair pair;
  Neighbor neighbor;
  Domain domain;
  Error error;
  Memory memory;
  Integrate integrate;
  std::vector<double> xs, fs;
  std::vector<double*> xr, fr;
  std::vector<int> type, mask;
  FakeLammps(int nall, int nlocal) {
    lmp.atom = &atom;
    lmp.update = &update;
    lmp.force = &force;
    lmp.neighbor = &neighbor;
    lmp.domain = &domain;
    lmp.error = &error;
    lmp.memory = &memory;
    lmp.world = 0;
    xs.assign((size_t)3 * nall, 0.0);
    fs.assign((size_t)3 * nall, 0.0);
    xr.resize(nall);
    fr.resize(nall);
    for (int i = 0; i < nall; i++) {
      xr[i] = &xs[3 * (size_t)i];
      fr[i] = &fs[3 * (size_t)i];
    }
    type.assign(nall, 1);
    mask.assign(nall, 1);
    atom.nlocal = nlocal;
    atom.nghost = nall - nlocal;
    atom.nmax = nall;
    atom.tag_enable = 1;
    atom.x = &xr[0];
    atom.f = &fr[0];
    atom.type = &type[0];
    atom.mask = &mask[0];
    update.ntimestep = 0;
    update.integrate_style = "verlet";
    update.integrate = &integrate;
    force.boltz = 0.0019872;
    force.newton_pair = 0;
    force.pair = &pair;
    pair.cutforce = 5.0;
    neighbor.skin = 1.0;
    for (int d = 0; d < 3; d++) {
      domain.boxlo[d] = domain.sublo[d] = 0.0;
      domain.boxhi[d] = domain.subhi[d] = 16.0;
      domain.prd[d] = 16.0;
      domain.periodicity[d] = 1;
    }
  }
};

static double urand() { return (double)rand() / ((double)RAND_MAX + 1.0); }

static void write_file(const std::string& fn, const std::string& text) {
  std::ofstream o(fn.c_str());
  o << text;
}

// ---- fix edm_pair on a caller-built half list with ghosts
static void pair_fix() {
  write_file("pair.edm",
             "tempering 1\nglobal_tempering 0.0001\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 0.1\nhill_density 250\n"
             "dimension 1\nbox_low 1.68\nbox_high 5.0\nbias_spacing 0.00025\nbias_sigma 0.025\nhills_filename HILLS_PAIRFIX\n"
             "histogram_filename HIST_PAIRFIX\n");
  const int nall = 4000, nlocal = 3000;
  FakeLammps L(nall, nlocal);
  for (int i = 0; i < nall; i++) {
    for (int d = 0; d < 3; d++) L.xs[3 * (size_t)i + d] = 30.0 * urand();
    L.type[i] = 1 + (i % 2);  // two species: the fix biases 1-2 pairs only
  }
  // The fix starts from est_hill_count = atom->nmax (lammps/fix_edm_pair.cpp:105).  With nmax = nall the first round
  // would accept ~10 000 hills and, like the reference, die on "The bias overflow buffer is full" (checked by
  // hand); a run with headroom in nmax gets through its first round.
  L.atom.nmax = 400000;
  const char* argv[] = {"1", "all", "edm_pair", "300.0", "pair.edm", "2", "4", "BIAS_PAIRFIX", "17", "1", "2"};
  FixEDMPair fix(&L.lmp, 11, (char**)argv);
  fix.init();
  NeighList list;
  std::vector<int> ilist(nlocal), numneigh(nall, 0);
  std::vector<std::vector<int> > neigh(nall);
  std::vector<int*> firstneigh(nall);
  for (int i = 0; i < nlocal; i++) ilist[i] = nlocal - 1 - i;  // LAMMPS does not promise an ascending ilist
  list.inum = nlocal;
  list.ilist = &ilist[0];
  list.numneigh = &numneigh[0];
  list.firstneigh = &firstneigh[0];
  fix.init_list(0, &list);
  const double rlist = L.pair.cutforce + L.neighbor.skin;
  std::vector<double> r, val, der, fref((size_t)3 * nall);
  std::vector<int> pi, pj;
  for (int step = 0; step < 7; step++) {
    L.update.ntimestep = step;
    L.neighbor.ago = step % 3;
    if (L.neighbor.ago == 0) {  // rebuild: half list (newton off: local-local pairs once, local-ghost from the local row)
      for (int i = 0; i < nall; i++) neigh[i].clear();
      for (int i = 0; i < nlocal; i++)
        for (int j = i + 1; j < nall; j++) {
          double d2 = 0;
          for (int d = 0; d < 3; d++) {
            const double t = L.xs[3 * (size_t)i + d] - L.xs[3 * (size_t)j + d];
            d2 += t * t;
          }
          if (d2 < rlist * rlist) neigh[i].push_back(j | (1 << 30));  // a special-bond bit the fix must strip
        }
      for (int i = 0; i < nall; i++) {
        numneigh[i] = (int)neigh[i].size();
        firstneigh[i] = neigh[i].empty() ? NULL : &neigh[i][0];
      }
    }
    // reference: every listed 1-2 pair at its CURRENT distance against the bias as it stands now
    pi.clear();
    pj.clear();
    r.clear();
    for (int ii = 0; ii < nlocal; ii++) {
      const int i = ilist[ii];
      for (size_t k = 0; k < neigh[i].size(); k++) {
        const int j = neigh[i][k] & NEIGHMASK;
        if (L.type[i] == L.type[j]) continue;
        double d2 = 0;
        for (int d = 0; d < 3; d++) {
          const double t = L.xs[3 * (size_t)i + d] - L.xs[3 * (size_t)j + d];
          d2 += t * t;
        }
        pi.push_back(i);
        pj.push_back(j);
        r.push_back(sqrt(d2));
      }
    }
    val.assign(r.size(), 0.0);
    der.assign(r.size(), 0.0);
    REQUIRE(edm_grid_eval(fix.bias->bias_->device_grid(), (long)r.size(), &r[0], 1, &val[0], &der[0]) == EDM_OK);
    double eref = 0;
    for (size_t a = 0; a < fref.size(); a++) fref[a] = 0.01 * (double)(a % 7);  // forces other styles left there
    L.fs = fref;
    for (size_t k = 0; k < r.size(); k++) {
      eref += val[k];
      for (int d = 0; d < 3; d++) {
        const double p = (L.xs[3 * (size_t)pi[k] + d] - L.xs[3 * (size_t)pj[k] + d]) / r[k] * (-der[k]);
        fref[3 * (size_t)pi[k] + d] += p;
        if (pj[k] < nlocal) fref[3 * (size_t)pj[k] + d] -= p;
      }
    }
    if (step == 0) fix.setup(0); else fix.post_force(0);
    const double e = fix.compute_scalar();
    REQUIRE(fabs(e - eref) <= 1e-10 * fabs(eref) + 1e-300);
    double worst = 0, scale = 0;
    for (size_t a = 0; a < fref.size(); a++) {
      worst = fmax(worst, fabs(L.fs[a] - fref[a]));
      scale = fmax(scale, fabs(fref[a]));
    }
    REQUIRE(worst <= 1e-10 * scale);
    if (step >= 2) REQUIRE(eref > 0);  // hills were deposited on steps 0 and 2 (stride 2)
    for (int i = 0; i < nall; i++)  // the atoms drift; the list is rebuilt only every third step
      for (int d = 0; d < 3; d++) L.xs[3 * (size_t)i + d] += 0.05 * (urand() - 0.5);
  }
  REQUIRE(fix.bias->cum_bias_ > 0);
  REQUIRE(fix.last_calls > 100000);  // the first round ran on the reference's crude estimate atom->nmax, later ones on this
  std::ifstream hills("HILLS_PAIRFIX_0"), biasf("BIAS_PAIRFIX"), table("BIAS_PAIRFIX.ltab");
  std::string line;
  REQUIRE((bool)std::getline(hills, line) && line.find(" h ") != std::string::npos);
  REQUIRE((bool)std::getline(biasf, line) && line.substr(0, 2) == "#!");
  REQUIRE(table.good());
}

// ---- fix edm on 2-D coordinates with a group mask
static void coord_fix() {
  write_file("coord.edm",
             "tempering 1\nglobal_tempering -1\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 1000\nhill_density 80\n"
             "dimension 2\nbox_low 0 0\nbox_high 16 16\nbias_spacing 0.0625 0.0625\nbias_sigma 0.25 0.25\n"
             "hills_filename HILLS_COORDFIX\nhistogram_filename HIST_COORDFIX\n");
  const int n = 6000;
  FakeLammps L(n, n);
  for (int i = 0; i < n; i++) L.mask[i] = 1 | ((i % 3 == 0) ? 4 : 0);
  const char* argv[] = {"2", "grp", "edm", "300.0", "coord.edm", "1", "3", "BIAS_COORDFIX", "5"};
  FixEDM fix(&L.lmp, 9, (char**)argv);
  fix.groupbit = 4;  // only every third atom is in the fix group
  fix.init();
  std::vector<double> pts, val, der, fref((size_t)3 * n);
  for (int step = 0; step < 5; step++) {
    L.update.ntimestep = step;
    for (int i = 0; i < n; i++)
      for (int d = 0; d < 3; d++) L.xs[3 * (size_t)i + d] = 16.0 * urand();
    pts.clear();
    for (int i = 0; i < n; i++)
      if (L.mask[i] & 4) {
        pts.push_back(L.xs[3 * (size_t)i]);
        pts.push_back(L.xs[3 * (size_t)i + 1]);
      }
    const long m = (long)pts.size() / 2;
    val.assign(m, 0.0);
    der.assign(2 * m, 0.0);
    REQUIRE(edm_grid_eval(fix.bias->bias_->device_grid(), m, &pts[0], 2, &val[0], &der[0]) == EDM_OK);
    double eref = 0;
    for (size_t a = 0; a < fref.size(); a++) fref[a] = 0.5 - 0.001 * (double)(a % 11);
    L.fs = fref;
    long k = 0;
    for (int i = 0; i < n; i++)
      if (L.mask[i] & 4) {
        eref += val[k];
        fref[3 * (size_t)i] -= der[2 * k];
        fref[3 * (size_t)i + 1] -= der[2 * k + 1];
        k++;
      }
    if (step == 0) fix.setup(0); else fix.post_force(0);
    REQUIRE(fabs(fix.compute_scalar() - eref) <= 1e-10 * fabs(eref) + 1e-300);
    double worst = 0, scale = 0;
    for (size_t a = 0; a < fref.size(); a++) {
      worst = fmax(worst, fabs(L.fs[a] - fref[a]));
      scale = fmax(scale, fabs(fref[a]));
    }
    REQUIRE(worst <= 1e-10 * scale);
    if (step) REQUIRE(eref > 0);
  }
  REQUIRE(fix.bias->cum_bias_ > 0);
  std::ifstream hills("HILLS_COORDFIX_0"), biasf("BIAS_COORDFIX");
  std::string line;
  REQUIRE((bool)std::getline(hills, line) && line.find(" h ") != std::string::npos);
  REQUIRE(biasf.good());
}

int main() {
  srand(4242);
  pair_fix();
  coord_fix();
  printf("%d checks, %d failed\n", g_checks, g_failed);
  return g_failed ? 1 : 0;
}
