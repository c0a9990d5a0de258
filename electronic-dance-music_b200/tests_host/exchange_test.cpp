// The hill exchange from C++, no Python and no torch anywhere: N replicas of one EDM::EDMBias, one per
// device, driven by one host thread each over communicators from edm_comm_init_all (ncclCommInitAll).
// Every rank proposes its own shard of candidates through the reference's pre_add_hill / add_hill /
// post_add_hill triple; post_add_hill all-gathers the accepted hills inside the library
// (lib/edm_bias.cpp:565-583, 614-706).  Checks:
//   - all replicas end bit-identical (grid, derivatives, cum_bias_, backlog),
//   - and bit-identical to ONE single-rank EDMBias fed the rank-major concatenation of the shards with
//     the job-wide est_hill_count (the parity oracle for P GPUs; that single-rank path is
//     itself checked against the CPU oracle by tests/test_gpu_parity.py).
// With one visible device it still runs: a 1-rank communicator exercises pack -> commit.
// Exit code 0 = all checks passed.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../edm/edm_bias.h"

static int g_failed = 0, g_checks = 0;
#define REQUIRE(c)                                          \
  do {                                                      \
    g_checks++;                                             \
    if (!(c)) {                                             \
      g_failed++;                                           \
      printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
    }                                                       \
  } while (0)

struct Config {
  const char* name;
  const char* text;
  int dim;
  double lo, hi;       // bias box per dimension
  double xlo, xhi;     // where candidates are drawn
  int periodic;
  int n_per_rank, steps;
};

static std::string write_edm(const std::string& dir, const Config& c, int who) {
  const std::string f = dir + "/" + c.name + "_" + std::to_string(who) + ".edm";
  std::ofstream o(f.c_str());
  o << c.text << "hills_filename " << dir << "/HILLS_" << c.name << "_" << who << "\nhistogram_filename " << dir
    << "/HIST_" << c.name << "_" << who << "\n";
  return f;
}

static EDM::EDMBias* make_bias(const std::string& file, const Config& c, int device) {
  EDM::set_default_device(device);
  EDM::EDMBias* b = new EDM::EDMBias(file);
  b->setup(300.0, 0.0019872);
  double lo[3] = {c.lo, c.lo, c.lo}, hi[3] = {c.hi, c.hi, c.hi}, skin[3] = {0, 0, 0};
  int per[3] = {c.periodic, c.periodic, c.periodic};
  b->subdivide(lo, hi, lo, hi, per, skin);
  return b;
}

static void shard_inputs(const Config& c, int rank, int step, std::vector<double>& x, std::vector<double>& u) {
  std::mt19937_64 gen(1234u + 7919u * (unsigned)rank + 104729u * (unsigned)step);
  std::uniform_real_distribution<double> pos(c.xlo, c.xhi), uni(0.0, 1.0);
  x.resize((size_t)c.n_per_rank * c.dim);
  u.resize(c.n_per_rank);
  for (int i = 0; i < c.n_per_rank; i++) {
    for (int d = 0; d < c.dim; d++) x[(size_t)i * c.dim + d] = pos(gen);
    u[i] = uni(gen);
  }
}

struct Snapshot {
  std::vector<double> v, d, backlog;
  long left, right;
  double cum;
};

static Snapshot snapshot(EDM::EDMBias* b) {
  Snapshot s;
  edm_grid_t* g = b->bias_->device_grid();
  size_t size = 0;
  int dim = 0;
  edm_grid_geometry(g, &dim, NULL, NULL, NULL, NULL, NULL, NULL, &size);
  s.v.resize(size);
  s.d.resize(size * dim);
  EDM::edm_check(edm_grid_download(g, s.v.data(), s.d.data()), "snapshot");
  s.backlog.resize(EDM_BUFFER_DBLS);
  EDM::edm_check(edm_bias_backlog_get(b->device_bias(), &s.left, &s.right, s.backlog.data()), "snapshot");
  s.cum = b->cum_bias_;
  return s;
}

static bool same(const Snapshot& a, const Snapshot& b) {
  return a.v.size() == b.v.size() && memcmp(a.v.data(), b.v.data(), a.v.size() * 8) == 0 &&
         memcmp(a.d.data(), b.d.data(), a.d.size() * 8) == 0 && a.left == b.left && a.right == b.right &&
         memcmp(a.backlog.data(), b.backlog.data(), a.backlog.size() * 8) == 0 &&
         memcmp(&a.cum, &b.cum, 8) == 0;
}

static void run_config(const Config& c, int nranks, const std::string& dir) {
  printf("config %s: %d rank(s), %d candidates per rank, %d steps\n", c.name, nranks, c.n_per_rank, c.steps);
  std::vector<edm_comm_t*> comms(nranks);
  EDM::edm_check(edm_comm_init_all(comms.data(), nranks, NULL), "edm_comm_init_all");
  std::vector<EDM::EDMBias*> rep(nranks);
  for (int r = 0; r < nranks; r++) {
    rep[r] = make_bias(write_edm(dir, c, r), c, r);
    rep[r]->set_comm(comms[r]);
    REQUIRE(rep[r]->mpi_size_ == nranks && rep[r]->mpi_rank_ == r);
  }
  EDM::EDMBias* single = make_bias(write_edm(dir, c, 99), c, 0);

  for (int step = 0; step < c.steps; step++) {
    // one host thread per rank, as one process per GPU would run it; post_add_hill meets in the all-gather
    std::vector<std::thread> th;
    for (int r = 0; r < nranks; r++) {
      th.emplace_back([&, r]() {
        std::vector<double> x, u;
        shard_inputs(c, r, step, x, u);
        EDM::EDMBias* b = rep[r];
        b->pre_add_hill(c.n_per_rank);  // this rank's count, as the reference's caller passes it
        for (int i = 0; i < c.n_per_rank; i++) b->add_hill(&x[(size_t)i * c.dim], u[i]);
        b->post_add_hill();
      });
    }
    for (auto& t : th) t.join();
    // the single-rank run over the rank-major concatenation
    single->pre_add_hill(c.n_per_rank * nranks);
    for (int r = 0; r < nranks; r++) {
      std::vector<double> x, u;
      shard_inputs(c, r, step, x, u);
      for (int i = 0; i < c.n_per_rank; i++) single->add_hill(&x[(size_t)i * c.dim], u[i]);
    }
    single->post_add_hill();
  }
  Snapshot s0 = snapshot(single);
  REQUIRE(s0.cum > 0.0);
  for (int r = 0; r < nranks; r++) {
    Snapshot sr = snapshot(rep[r]);
    REQUIRE(same(sr, s0));
    REQUIRE(edm_bias_check(rep[r]->device_bias()) == EDM_OK);
  }
  // an exchange block that is too small must be reported, never truncated silently
  if (nranks >= 1) {
    rep[0]->set_comm(comms[0], 2);  // 2 records per rank
    for (int r = 1; r < nranks; r++) rep[r]->set_comm(comms[r], 2);
    std::vector<int> rc(nranks, 0);
    std::vector<std::thread> th;
    for (int r = 0; r < nranks; r++) {
      th.emplace_back([&, r]() {
        std::vector<double> x, u;
        shard_inputs(c, r, 1000, x, u);
        edm_bias_t* b = rep[r]->device_bias();
        edm_bias_pre_add_hill(b, c.n_per_rank);
        edm_bias_add_hill_batch(b, c.n_per_rank, x.data(), u.data());
        rc[r] = edm_bias_post_add_hill(b);
      });
    }
    for (auto& t : th) t.join();
    int overflowed = 0;
    for (int r = 0; r < nranks; r++) overflowed += (rc[r] == EDM_ERR_CAPACITY);
    REQUIRE(overflowed > 0);
  }
  for (int r = 0; r < nranks; r++) {
    delete rep[r];
    edm_comm_destroy(comms[r]);
  }
  delete single;
}

int main(int argc, char** argv) {
  const std::string dir = argc > 1 ? argv[1] : "/tmp";
  int ndev = 0;
  if (edm_device_count(&ndev) != EDM_OK || ndev == 0) {
    printf("exchange_test: no CUDA device (there is no CPU fallback)\n");
    return 2;
  }
  int want = argc > 2 ? atoi(argv[2]) : ndev;
  if (want > ndev) want = ndev;
  if (want > 8) want = 8;
  int ver = 0;
  if (want > 1) {
    EDM::edm_check(edm_comm_nccl_version(&ver), "edm_comm_nccl_version");
    printf("NCCL %d, %d devices\n", ver, want);
  }
  const Config configs[] = {
      {"pair_rdf_limiter",
       "tempering 1\nglobal_tempering 0.0001\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 0.004\n"
       "hill_density 250\ndimension 1\nbox_low 1.68\nbox_high 5.0\nbias_spacing 0.00025\nbias_sigma 0.025\n",
       1, 1.68, 5.0, 0.5, 5.5, 0, 20000, 6},
      {"coord_2d_local_tempering",
       "tempering 1\nglobal_tempering -1\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 1000\n"
       "hill_density 200\ndimension 2\nbox_low 0 0\nbox_high 16 16\nbias_spacing 0.03125 0.03125\n"
       "bias_sigma 0.0625 0.0625\n",
       2, 0.0, 16.0, -1.0, 17.0, 1, 10000, 4},
      {"coord_3d",
       "tempering 0\nhill_prefactor 0.02\nbias_per_step 1000\nhill_density 120\ndimension 3\nbox_low 0 0 0\n"
       "box_high 8 8 8\nbias_spacing 0.125 0.125 0.125\nbias_sigma 0.25 0.25 0.25\n",
       3, 0.0, 8.0, -0.5, 8.5, 1, 6000, 3},
  };
  for (const Config& c : configs) run_config(c, want, dir);
  printf("exchange_test: %d checks, %d failed (%d rank(s))\n", g_checks, g_failed, want);
  if (g_failed == 0) printf("EXCHANGE_TEST_OK ranks=%d\n", want);
  return g_failed ? 1 : 0;
}
