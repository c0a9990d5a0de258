// One driver, two builds: the on-disk formats either side of the hot path.
//
//   - against the UNMODIFIED reference library (the checker's Makefile target `text`, headers from /root/reference/lib, the
//     single-rank stub mpi.h): produces the golden texts committed under tests/golden/text/
//     (tests/golden/make_text_golden.py);
//   - against this repo's EDM:: mirror over the CUDA library (build.py): produces the same files from the GPU
//     state, and tests/test_gpu_text_io.py compares them token by token with the goldens.
//
// It only uses the API the two share (lib/edm_bias.h:36-115, lib/grid.h:146-180, lib/gaussian_grid.h:47-55):
// EDMBias(file), setup, subdivide, add_hills, write_bias, write_histogram, write_lammps_table,
// bias_->multi_write / lammps_multi_write, read_grid, Grid::write.  Files written into the working directory:
//   BIAS  HIST  HILLS_0  LTAB  MULTI  [LMULTI]  for each bias case (prefix = case name), and for the restart case
//   a second generation of them after re-reading BIAS through `initial_bias_filename`;
//   FIX1.out FIX2.out FIX3.out = the reference's own grid fixtures read and written back.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <new>
#include <string>
#include <vector>

#include "edm_bias.h"

using namespace EDM;

struct Case {
  const char* name;
  const char* edm;  // without file names
  int dim;
  double sublo[3], subhi[3], boxlo[3], boxhi[3];
  int periodic[3];
  double skin[3];
  double T, kB;
  int steps, n;
  double xlo, xhi;  // candidate positions are drawn from [xlo, xhi) in every dimension
};

static const Case kCases[] = {
    // fix edm_pair geometry: grid [-skin, cut + 2 skin] around walls at 1.68 / 5.0, McGDP hills, limiter active,
    // so the hill log carries h / u / b / v lines and the LAMMPS table has its zero-filled head
    {"rdf1d",
     "tempering 1\nglobal_tempering 0.0001\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 0.004\n"
     "hill_density 40\ndimension 1\nbox_low 1.68\nbox_high 5.0\nbias_spacing 0.05\nbias_sigma 0.2\n",
     1, {0, 0, 0}, {6, 0, 0}, {0, 0, 0}, {6, 0, 0}, {0, 0, 0}, {1, 0, 0}, 300.0, 0.0019872, 5, 400, 0.5, 6.5},
    // periodic 2-D box, local well-tempering
    {"coord2d",
     "tempering 1\nglobal_tempering -1\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 1000\nhill_density 30\n"
     "dimension 2\nbox_low 0 0\nbox_high 4 3\nbias_spacing 0.25 0.2\nbias_sigma 0.4 0.3\n",
     2, {0, 0, 0}, {4, 3, 0}, {0, 0, 0}, {4, 3, 0}, {1, 1, 0}, {0, 0, 0}, 300.0, 0.0019872, 3, 300, 0.0, 3.0},
    // 3-D, mixed periodicity: walls in y
    {"coord3d",
     "tempering 0\nhill_prefactor 0.05\nbias_per_step 1000\nhill_density 20\ndimension 3\nbox_low 0 0 0\n"
     "box_high 2 2 2\nbias_spacing 0.25 0.25 0.25\nbias_sigma 0.3 0.3 0.3\n",
     3, {0, 0, 0}, {2, 2, 2}, {0, 0, 0}, {2, 2, 2}, {1, 0, 1}, {0, 0, 0}, 1.0, 1.0, 2, 200, 0.0, 2.0},
};

// splitmix64: the same stream in both builds, no library generator involved
static unsigned long long g_state;
static double next_uniform() {
  g_state += 0x9E3779B97F4A7C15ULL;
  unsigned long long z = g_state;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

static EDMBias* make_bias(const std::string& file) {
  // zero-filled storage: the reference never initialises its overflow buffer
  void* mem = calloc(1, sizeof(EDMBias));
  return new (mem) EDMBias(file);
}

static void run_steps(EDMBias* b, const Case& c, int first_step, int steps) {
  std::vector<double> x((size_t)c.n * 3), u(c.n);
  std::vector<const double*> rows(c.n);
  for (int s = first_step; s < first_step + steps; s++) {
    g_state = 0x1234567ULL + 7919ULL * (unsigned long long)s + 104729ULL * (unsigned long long)c.dim;
    for (int i = 0; i < c.n; i++) {
      for (int d = 0; d < 3; d++) x[(size_t)3 * i + d] = c.xlo + (c.xhi - c.xlo) * next_uniform();
      u[i] = next_uniform();
      rows[i] = &x[(size_t)3 * i];
    }
    b->add_hills(c.n, rows.data(), u.data());
  }
}

static void write_all(EDMBias* b, const Case& c, const std::string& tag) {
  b->write_bias(tag + "_BIAS");
  b->write_histogram();
  b->write_lammps_table(tag + "_LTAB");
  b->bias_->multi_write(tag + "_MULTI");
  if (c.dim == 1) b->bias_->lammps_multi_write(tag + "_LMULTI");
  b->hill_output_.flush();
}

static void run_case(const Case& c) {
  const std::string name(c.name);
  {
    std::ofstream o((name + ".edm").c_str());
    o << c.edm << "hills_filename " << name << "_HILLS\nhistogram_filename " << name << "_HIST\n";
  }
  EDMBias* b = make_bias(name + ".edm");
  b->setup(c.T, c.kB);
  b->subdivide(c.sublo, c.subhi, c.boxlo, c.boxhi, c.periodic, c.skin);
  run_steps(b, c, 0, c.steps);
  write_all(b, c, name);
  // restart (lib/edm_bias.cpp:166-167, 1066-1072): a second EDMBias starts from the bias file just written,
  // runs on, and writes its own generation of files
  {
    std::ofstream o((name + "_restart.edm").c_str());
    o << c.edm << "initial_bias_filename " << name << "_BIAS\nhills_filename " << name << "_restart_HILLS\n"
      << "histogram_filename " << name << "_restart_HIST\n";
  }
  EDMBias* r = make_bias(name + "_restart.edm");
  r->setup(c.T, c.kB);
  r->subdivide(c.sublo, c.subhi, c.boxlo, c.boxhi, c.periodic, c.skin);
  run_steps(r, c, c.steps, 2);
  write_all(r, c, name + "_restart");
  printf("case %s: cum_bias %.17g after %d steps, %.17g after the restart\n", c.name, b->cum_bias_, c.steps, r->cum_bias_);
}

int main(int argc, char** argv) {
  // fixtures of the reference's own test-suite (tests/1.grid, 2.grid, 3.grid): read them, write them back
  const std::string fixtures = argc > 1 ? argv[1] : "";
  if (!fixtures.empty()) {
    for (int d = 1; d <= 3; d++) {
      char in[1024], out[64];
      snprintf(in, sizeof(in), "%s/%d.grid", fixtures.c_str(), d);
      snprintf(out, sizeof(out), "FIX%d.out", d);
      Grid* g = read_grid(d, in, 0);
      g->write(out);
      printf("fixture %d.grid: %zu points\n", d, g->get_grid_size());
    }
  }
  for (size_t i = 0; i < sizeof(kCases) / sizeof(kCases[0]); i++) run_case(kCases[i]);
  printf("TEXT_IO_DRIVER_OK\n");
  return 0;
}
