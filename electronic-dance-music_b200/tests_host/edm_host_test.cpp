// The reference's unit tests (tests/edm_test.cpp), restated against the B200 build of the same C++
// API.  Same case names, same checks and tolerances; Boost.Test is not in this image, so a few
// macros stand in for it.  Run on a GPU box: every get_value / add_value below is a kernel launch.
//
//   edm_host_test <dir with 1.grid 2.grid 3.grid sanity.edm read_test.edm or "-"> [case-substring]
//
// With "-" the file-based cases are skipped (the fixtures live in the reference tree).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../edm/edm_bias.h"
#include "../edm/gaussian_grid.h"
#include "../edm/grid.h"

using namespace EDM;

#define EPSILON 1e-10
static int g_failed = 0, g_checks = 0;
static const char* g_case = "";
#define REQUIRE(cond)                                                           \
  do {                                                                          \
    g_checks++;                                                                 \
    if (!(cond)) {                                                              \
      g_failed++;                                                               \
      printf("FAIL %s:%d [%s] %s\n", __FILE__, __LINE__, g_case, #cond);        \
    }                                                                           \
  } while (0)
#define REQUIRE_EQUAL(a, b) REQUIRE((a) == (b))

static std::string g_src;
static const char* g_filter = NULL;
static bool wanted(const char* name) {
  g_case = name;
  if (g_filter && !strstr(name, g_filter)) return false;
  printf("case %s\n", name);
  return true;
}
static bool have_files() { return g_src != "-"; }

static void grid_1d_sanity() {  // edm_test.cpp:25-59
  double min[] = {0}, max[] = {10}, bin_spacing[] = {1};
  int periodic[] = {0};
  DimmedGrid<1> g(min, max, bin_spacing, periodic, 0, 0);
  REQUIRE_EQUAL(g.grid_number_[0], 11);
  REQUIRE_EQUAL(g.grid_size_, (size_t)11);
  size_t array[] = {5}, temp[1];
  g.one2multi(g.multi2one(array), temp);
  REQUIRE_EQUAL(array[0], temp[0]);
  for (int i = 0; i < 10; i++) g.grid_[i] = i;
  double x[] = {3.5};
  REQUIRE(pow(g.get_value(x) - 3, 2) < 0.000001);
  x[0] = 0;
  g.get_value(x);
  x[0] = 10;
  g.get_value(x);
}

static void grid_3d_sanity() {  // edm_test.cpp:61-107
  double min[] = {-2, -5, -3}, max[] = {125, 63, 78}, bin_spacing[] = {1.27, 1.36, 0.643};
  int periodic[] = {0, 1, 1};
  DimmedGrid<3> g(min, max, bin_spacing, periodic, 0, 0);
  REQUIRE_EQUAL(g.grid_number_[0], 101);
  REQUIRE_EQUAL(g.grid_number_[1], 50);
  REQUIRE_EQUAL(g.grid_number_[2], 126);
  size_t array[3], temp[3];
  for (int i = 0; i < g.grid_number_[0]; i++)
    for (int j = 0; j < g.grid_number_[1]; j++)
      for (int k = 0; k < g.grid_number_[2]; k++) {
        array[0] = i;
        array[1] = j;
        array[2] = k;
        g.one2multi(g.multi2one(array), temp);
        REQUIRE(array[0] == temp[0] && array[1] == temp[1] && array[2] == temp[2]);
        g.grid_[g.multi2one(array)] = g.multi2one(array);
      }
  // batched instead of one launch per point: same points as the reference's triple loop
  std::vector<double> pts, want;
  for (int i = 0; i < g.grid_number_[0] - 1; i++)  // the extra non-periodic point is outside in_grid
    for (int j = 0; j < g.grid_number_[1]; j += 7)
      for (int k = 0; k < g.grid_number_[2]; k += 5) {
        pts.push_back(i * g.dx_[0] + g.min_[0] + EPSILON);
        pts.push_back(j * g.dx_[1] + g.min_[1] + EPSILON);
        pts.push_back(k * g.dx_[2] + g.min_[2] + EPSILON);
        array[0] = i;
        array[1] = j;
        array[2] = k;
        want.push_back((double)g.multi2one(array));
      }
  std::vector<double> val(want.size());
  edm_check(edm_grid_get_value(g.device_grid(), (long)want.size(), pts.data(), 3, val.data()), "test");
  bool ok = true;
  for (size_t p = 0; p < want.size(); p++) ok = ok && pow(val[p] - want[p], 2) < 0.0000001;
  REQUIRE(ok);
}

static void grid_reads() {  // edm_test.cpp:109-138
  if (!have_files()) return;
  DimmedGrid<1> g1(g_src + "/1.grid");
  REQUIRE_EQUAL(g1.min_[0], 0);
  REQUIRE_EQUAL(g1.max_[0], 2.5 + g1.dx_[0]);
  REQUIRE_EQUAL(g1.grid_number_[0], 101);
  DimmedGrid<3> g(g_src + "/3.grid");
  REQUIRE_EQUAL(g.min_[2], 0);
  REQUIRE_EQUAL(g.max_[2], 2.5 + g.dx_[2]);
  REQUIRE_EQUAL(g.grid_number_[2], 11);
  double temp[] = {0.75, 0, 1.00};
  REQUIRE(pow(g.get_value(temp) - 1.260095, 2) < EPSILON);
  g.b_interpolate_ = 1;
  g.set_interpolation(1);
  double temp2[] = {0.76, 0, 1.00};
  REQUIRE(g.get_value(temp2) > g.get_value(temp));
  temp2[0] = 0.75;
  temp2[2] = 0.99;
  REQUIRE(g.get_value(temp2) < g.get_value(temp));
}

static void grid_read_write_consistency() {  // edm_test.cpp:142-180
  if (!have_files()) return;
  for (int i = 1; i <= 3; i++) {
    std::stringstream fn;
    fn << i << ".grid";
    std::string input = g_src + "/" + fn.str(), output = fn.str() + ".test";
    Grid* g = i == 1 ? (Grid*)new DimmedGrid<1>(input) : i == 2 ? (Grid*)new DimmedGrid<2>(input)
                                                                : (Grid*)new DimmedGrid<3>(input);
    g->write(output);
    size_t n = g->get_grid_size();
    std::vector<double> ref(g->get_grid(), g->get_grid() + n);
    g->read(output);
    REQUIRE_EQUAL(g->get_grid_size(), n);
    bool ok = true;
    for (size_t j = 0; j < n; j++) ok = ok && pow(ref[j] - g->get_grid()[j], 2) < EPSILON;
    REQUIRE(ok);
    delete g;
  }
}

static void interpolation_1d() {  // edm_test.cpp:182-218
  double min[] = {0}, max[] = {10}, bin_spacing[] = {1};
  int periodic[] = {0};
  DimmedGrid<1> g(min, max, bin_spacing, periodic, 1, 1);
  for (int i = 0; i < 11; i++) {
    g.grid_[i] = log((double)i);
    g.grid_deriv_[i] = 1. / i;
  }
  double array[] = {5.3}, der[1];
  double fhat = g.get_value_deriv(array, der);
  REQUIRE(fhat > log(5.) && fhat < log(6.));
  REQUIRE(der[0] < 1. / 5 && der[0] > 1. / 6.);
  REQUIRE(pow(fhat - log(5.3), 2) < 0.1);
  REQUIRE(pow(der[0] - 1. / 5.3, 2) < 0.1);
}

static void interp_1d_periodic() {  // edm_test.cpp:220-250
  double min[] = {-M_PI}, max[] = {M_PI}, bin_spacing[] = {M_PI / 100};
  int periodic[] = {1};
  DimmedGrid<1> g(min, max, bin_spacing, periodic, 1, 1);
  for (int i = 0; i < g.grid_number_[0]; i++) {
    g.grid_[i] = sin(g.min_[0] + i * g.dx_[0]);
    g.grid_deriv_[i] = cos(g.min_[0] + i * g.dx_[0]);
  }
  double array[] = {M_PI / 4}, der[1];
  double fhat = g.get_value_deriv(array, der);
  REQUIRE(pow(fhat - sin(array[0]), 2) < 0.1);
  REQUIRE(pow(der[0] - cos(array[0]), 2) < 0.1);
  array[0] = 5 * M_PI / 4;
  fhat = g.get_value_deriv(array, der);
  REQUIRE(pow(fhat - sin(array[0]), 2) < 0.1);
  REQUIRE(pow(der[0] - cos(array[0]), 2) < 0.1);
}

static void boundary_remap() {  // edm_test.cpp:252-387
  {
    double min[] = {0, 0}, max[] = {10, 5}, bin_spacing[] = {1, 1}, sigma[] = {0.1, 0.1};
    int periodic[] = {1, 0, 0};
    DimmedGaussGrid<2> g(min, max, bin_spacing, periodic, 1, sigma);
    max[1] = 10;
    periodic[1] = 1;
    g.set_boundary(min, max, periodic);
    double in[6][2] = {{0, 1}, {-1, 1}, {9, 6}, {9, 11}, {9, 9}, {9, -1}};
    double out[6][2] = {{0, 1}, {9, 1}, {9, 6}, {9, 1}, {9, -1}, {9, -1}};
    for (int k = 0; k < 6; k++) {
      double p[2] = {in[k][0], in[k][1]};
      g.remap(p);
      REQUIRE(pow(p[0] - out[k][0], 2) < 0.1 && pow(p[1] - out[k][1], 2) < 0.1);
    }
  }
  {
    double min[] = {-2}, max[] = {7}, bin_spacing[] = {0.1}, sigma[] = {0.1};
    int periodic[] = {0};
    DimmedGaussGrid<1> g(min, max, bin_spacing, periodic, 1, sigma);
    min[0] = 0;
    max[0] = 10;
    periodic[0] = 1;
    g.set_boundary(min, max, periodic);
    double in[4] = {0, -1, 9, 6}, out[4] = {0, -1, -1, 6};
    for (int k = 0; k < 4; k++) {
      double p[1] = {in[k]};
      g.remap(p);
      REQUIRE(pow(p[0] - out[k], 2) < 0.1);
    }
    double point[] = {0.01}, der[1];
    g.add_value(point, 1);
    point[0] = 0;
    g.get_value_deriv(point, der);
    REQUIRE(fabs(der[0]) > 0.1);
  }
}

static void interp_3d_mixed() {  // edm_test.cpp:392-430
  double min[] = {-M_PI, -M_PI, 0}, max[] = {M_PI, M_PI, 10}, bin_spacing[] = {M_PI / 100, M_PI / 100, 1};
  int periodic[] = {1, 1, 0};
  DimmedGrid<3> g(min, max, bin_spacing, periodic, 1, 0);
  size_t index = 0;
  for (int i = 0; i < g.grid_number_[2]; i++)
    for (int j = 0; j < g.grid_number_[1]; j++)
      for (int k = 0; k < g.grid_number_[0]; k++) {
        double x = g.min_[0] + k * g.dx_[0], y = g.min_[1] + j * g.dx_[1], z = g.min_[2] + i * g.dx_[2];
        g.grid_[index] = cos(x) * sin(y) * z;
        g.grid_deriv_[index * 3 + 0] = -sin(x) * sin(y) * z;
        g.grid_deriv_[index * 3 + 1] = cos(x) * cos(y) * z;
        g.grid_deriv_[index * 3 + 2] = cos(x) * sin(y);
        index++;
      }
  g.set_interpolation(1);
  double array[] = {-10.75 * M_PI / 2, 8.43 * M_PI / 2, 3.5}, der[3];
  double fhat = g.get_value_deriv(array, der);
  double f = cos(array[0]) * sin(array[1]) * array[2];
  double td[] = {-sin(array[0]) * sin(array[1]) * array[2], cos(array[0]) * cos(array[1]) * array[2],
                 cos(array[0]) * sin(array[1])};
  REQUIRE(pow(f - fhat, 2) < 0.1);
  REQUIRE(pow(der[0] - td[0], 2) < 0.1 && pow(der[1] - td[1], 2) < 0.1 && pow(der[2] - td[2], 2) < 0.1);
}

static void gauss_grid_add_check() {  // edm_test.cpp:432-457
  double min[] = {-10}, max[] = {10}, sigma[] = {1}, bin_spacing[] = {1};
  int periodic[] = {1};
  DimmedGaussGrid<1> g(min, max, bin_spacing, periodic, 0, sigma);
  double x[] = {0}, der[1];
  g.add_value(x, 1);
  REQUIRE(pow(g.get_value(x) - 1 / sqrt(2 * M_PI), 2) < EPSILON);
  for (int i = -6; i < 7; i++) {
    x[0] = i;
    double value = g.get_value_deriv(x, der);
    REQUIRE(pow(value - exp(-x[0] * x[0] / 2.) / sqrt(2 * M_PI), 2) < 0.01);
    REQUIRE(pow(der[0] - (-x[0] * exp(-x[0] * x[0] / 2.)) / sqrt(2 * M_PI), 2) < 0.01);
  }
}

static void gauss_pbc_checks() {  // edm_test.cpp:460-534
  {
    double min[] = {2}, max[] = {10}, sigma[] = {1}, bin_spacing[] = {1};
    int periodic[] = {1};
    DimmedGaussGrid<1> g(min, max, bin_spacing, periodic, 0, sigma);
    double x[] = {2}, der[1];
    g.add_value(x, 1);
    for (int i = -6; i < 7; i++) {
      x[0] = i;
      double dx = x[0] - 2;
      dx -= round(dx / (min[0] - max[0])) * (min[0] - max[0]);
      double value = g.get_value_deriv(x, der);
      REQUIRE(pow(value - exp(-dx * dx / 2.) / sqrt(2 * M_PI), 2) < 0.01);
      REQUIRE(pow(der[0] - (-dx * exp(-dx * dx / 2.)) / sqrt(2 * M_PI), 2) < 0.01);
    }
  }
  {
    double min[] = {2}, max[] = {4}, sigma[] = {1}, bin_spacing[] = {1}, gauss_loc[] = {11}, x[1], der[1];
    int periodic[] = {0};
    DimmedGaussGrid<1> g(min, max, bin_spacing, periodic, 0, sigma);
    periodic[0] = 1;
    max[0] = 10;
    g.set_boundary(min, max, periodic);
    g.add_value(gauss_loc, 1);
    for (int i = 2; i < 4; i++) {
      x[0] = i;
      double dx = x[0] - gauss_loc[0];
      dx -= round(dx / (min[0] - max[0])) * (min[0] - max[0]);
      double value = g.get_value_deriv(x, der);
      REQUIRE(pow(value - exp(-dx * dx / 2.) / sqrt(2 * M_PI), 2) < 0.01);
      REQUIRE(pow(der[0] - (-dx * exp(-dx * dx / 2.)) / sqrt(2 * M_PI), 2) < 0.01);
    }
  }
}

static void gauss_grid_integral_tests() {  // edm_test.cpp:537-628
  for (int mcgdp = 0; mcgdp < 2; mcgdp++) {
    double min[] = {-100}, max[] = {100}, sigma[] = {mcgdp ? 10.0 : 1.2}, bin_spacing[] = {1};
    int periodic[] = {mcgdp ? 0 : 1};
    DimmedGaussGrid<1> g(min, max, bin_spacing, periodic, 1, sigma);
    int N = 20;
    double x[1], g_integral = 0;
    if (mcgdp) {
      x[0] = -100.0;
      g_integral += g.add_value(x, 1.5);
      x[0] = 100.0;
      g_integral += g.add_value(x, 1.5);
    }
    for (int i = 0; i < N; i++) {
      x[0] = rand() % 200 - 100 + i * (1. / N);
      g_integral += g.add_value(x, 1.5);
    }
    double dx = 0.1;
    int bins = (int)(200 / dx);
    std::vector<double> pts(bins), val(bins);
    for (int i = 0; i < bins; i++) pts[i] = -100 + i * dx;
    edm_check(edm_grid_get_value(g.device_grid(), bins, pts.data(), 1, val.data()), "test");
    double area = 0;
    for (int i = 0; i < bins; i++) area += val[i] * dx;
    REQUIRE(pow(area - (mcgdp ? N + 2 : N) * 1.5, 2) < (mcgdp ? 12 : 1));
    REQUIRE(pow(area - g_integral, 2) < 0.1);
  }
}

static void gauss_grid_derivative_tests() {  // edm_test.cpp:631-721
  for (int mcgdp = 0; mcgdp < 2; mcgdp++) {
    double min[] = {-100}, max[] = {100}, sigma[] = {1.2}, bin_spacing[] = {1};
    int periodic[] = {mcgdp ? 0 : 1};
    DimmedGaussGrid<1> g(min, max, bin_spacing, periodic, 1, sigma);
    int N = 20;
    double x[1];
    for (int i = 0; i < N; i++) {
      x[0] = rand() % 200 - 100 + i * (1. / N);
      g.add_value(x, 1.5);
    }
    double dx = 0.1;
    int bins = (int)(200 / dx);
    std::vector<double> pts(bins), val(bins), der(bins);
    for (int i = 0; i < bins; i++) pts[i] = -100 + i * dx;
    g.get_value_deriv_batch(bins, pts.data(), 1, val.data(), der.data());
    bool ok = true;
    for (int i = 2; i < bins; i++) {
      double approx = (val[i] - val[i - 2]) / (2 * dx);
      ok = ok && pow(approx - der[i - 1], 2) < (mcgdp ? 0.001 : 0.01);
    }
    REQUIRE(ok);
    if (mcgdp) {
      REQUIRE(pow(der[0], 2) < 0.001 && pow(der[1], 2) < 0.001);
      REQUIRE(pow(der[bins - 1], 2) < 0.01);
    }
  }
}

static void gauss_grid_interp_test_mcgdp() {  // edm_test.cpp:723-818
  {
    double min[] = {-100}, max[] = {100}, sigma[] = {10.0}, bin_spacing[] = {1};
    int periodic[] = {1};
    DimmedGaussGrid<1> g(min, max, bin_spacing, periodic, 1, sigma);
    periodic[0] = 0;
    min[0] = -50;
    max[0] = 50;
    g.set_boundary(min, max, periodic);
    double x[1], der[1];
    for (int i = 0; i < 20; i++) {
      x[0] = rand() % 200 - 100;
      g.add_value(x, 1.0);
    }
    REQUIRE(pow(g.grid_.grid_[50] - g.grid_.grid_[49], 2) < EPSILON);
    REQUIRE(pow(g.grid_.grid_[150] - g.grid_.grid_[151], 2) < EPSILON);
    x[0] = 50.0;
    g.get_value_deriv(x, der);
    REQUIRE(der[0] * der[0] < EPSILON);
    x[0] = -50.0;
    g.get_value_deriv(x, der);
    REQUIRE(der[0] * der[0] < EPSILON);
  }
  {
    double min[] = {-10, -10, -10}, max[] = {10, 10, 10}, sigma[] = {3.0, 3.0, 3.0}, bin_spacing[] = {0.9, 1.1, 1.4};
    int periodic[] = {1, 1, 1};
    DimmedGaussGrid<3> g(min, max, bin_spacing, periodic, 1, sigma);
    periodic[0] = periodic[1] = periodic[2] = 0;
    min[0] = min[1] = min[2] = -5;
    max[0] = max[1] = max[2] = 5;
    g.set_boundary(min, max, periodic);
    double x[3], der[3];
    for (int i = 0; i < 20; i++) {
      x[0] = rand() % 20 - 10;
      x[1] = rand() % 20 - 10;
      x[2] = rand() % 20 - 10;
      g.add_value(x, 5.0);
    }
    x[0] = x[2] = -5.0;
    x[1] = 5.0;
    g.get_value_deriv(x, der);
    REQUIRE(der[0] * der[0] < EPSILON);
  }
}

static void gauss_grid_integral_regression_1() {  // edm_test.cpp:823-843
  double min[] = {0}, max[] = {10}, bin_spacing[] = {0.009765625}, sigma[] = {0.1};
  int periodic[] = {1};
  GaussGrid* g = make_gauss_grid(1, min, max, bin_spacing, periodic, 1, sigma);
  g->set_boundary(min, max, periodic);
  double x[] = {-3.91944};
  REQUIRE(pow(g->add_value(x, 1.0) - 1.0, 2) < 0.1);
  delete g;
}

static void write_file(const std::string& fn, const std::string& text) {
  std::ofstream o(fn.c_str());
  o << text;
}

static void edm_bias_reader() {  // edm_test.cpp:846-852 (read_test.edm depends on 2.grid.test written above)
  if (!have_files()) return;
  EDMBias bias(g_src + "/read_test.edm");
  REQUIRE_EQUAL(bias.dim_, 2u);
  REQUIRE_EQUAL(bias.b_tempering_, 0);
  REQUIRE(pow(bias.bias_sigma_[0] - 2, 2) < EPSILON);
  REQUIRE(pow(bias.bias_dx_[1] - 1.0, 2) < EPSILON);
}

static void edm_sanity() {  // edm_test.cpp:856-905 with tests/sanity.edm restated inline
  write_file("sanity_host.edm",
             "tempering\t\t0\nhill_prefactor \t\t0.25\ndimension \t\t1\nbox_low\t\t\t0\nbox_high\t\t10\n"
             "bias_spacing\t\t0.009765625\nbias_sigma\t\t0.1\n");
  EDMBias bias("sanity_host.edm");
  bias.setup(1, 1);
  double low[] = {0, 0, 0}, high[] = {10, 0, 0}, skin[] = {0, 0, 0};
  int p[] = {1, 0, 0};
  bias.subdivide(low, high, low, high, p, skin);
  double** positions = (double**)malloc(sizeof(double*));
  positions[0] = (double*)malloc(sizeof(double));
  double runiform[] = {1};
  positions[0][0] = 5.0;
  bias.add_hills(1, positions, runiform);
  bias.write_bias("BIAS");
  REQUIRE(pow(bias.bias_->get_value(positions[0]) - bias.hill_prefactor_ / sqrt(2 * M_PI) / bias.bias_sigma_[0], 2) <
          EPSILON);
  REQUIRE(pow(bias.cum_bias_ - bias.hill_prefactor_, 2) < 0.001);
  double der[1];
  positions[0][0] = 4.99;
  bias.bias_->get_value_deriv(positions[0], der);
  REQUIRE(-der[0] < 0);
  positions[0][0] = 5.01;
  bias.bias_->get_value_deriv(positions[0], der);
  REQUIRE(-der[0] > 0);
  // update_forces through row pointers, and the HILLS line of the one hill
  double** forces = (double**)malloc(sizeof(double*));
  forces[0] = (double*)calloc(1, sizeof(double));
  positions[0][0] = 5.01;
  double e = bias.update_forces(1, positions, forces);
  REQUIRE(e > 0 && forces[0][0] > 0);
  bias.hill_output_.flush();
  std::ifstream hills("HILLS_0");
  std::string line;
  std::getline(hills, line);
  REQUIRE(line.substr(0, 16) == "0 h 1 5.00000000");
  free(positions[0]);
  free(positions);
  free(forces[0]);
  free(forces);
}

static void notebook_vector() {  // python-example/EDM.ipynb:103
  write_file("nb_host.edm",
             "tempering 0\nhill_prefactor 1.0\ndimension 1\nbox_low 0.0\nbox_high 1.0\nbias_spacing 0.01\n"
             "bias_sigma 0.5\nhills_filename HILLS_NB\n");
  EDMBias bias("nb_host.edm");
  bias.setup(1, 1);
  double low[] = {0, 0, 0}, high[] = {10, 0, 0}, skin[] = {0, 0, 0};
  int p[] = {0, 0, 0};
  bias.subdivide(low, high, low, high, p, skin);
  double x[] = {0.25};
  bias.pre_add_hill(1);
  bias.add_hill(x, 0.0);
  bias.post_add_hill();
  x[0] = 0.24;
  double force[] = {0};
  double e = bias.update_force(x, force);
  REQUIRE(fabs(e - 1.1002417338159258) < 1e-10 * 1.1002417338159258);
  REQUIRE(fabs(-force[0] - (-0.6144025830861709)) < 1e-10 * 0.6144025830861709);
}

// update_forces_add_hills (fix edm's post_force as one call) against update_forces + add_hills on a
// second bias fed the same LAMMPS-style rows: identical hills, forces and energies.
static void fused_step_matches_two_calls() {
  write_file("fused_host.edm",
             "tempering 1\nglobal_tempering -1\nbias_factor 5\nhill_prefactor 0.02\nbias_per_step 1000\nhill_density 60\n"
             "dimension 2\nbox_low 0 0\nbox_high 8 8\nbias_spacing 0.0625 0.0625\nbias_sigma 0.25 0.25\n"
             "hills_filename HILLS_FUSED\n");
  double low[] = {0, 0, 0}, high[] = {8, 8, 0}, skin[] = {0, 0, 0};
  int p[] = {1, 1, 0};
  EDMBias a("fused_host.edm"), b("fused_host.edm");
  a.setup(300, 0.0019872);
  b.setup(300, 0.0019872);
  a.subdivide(low, high, low, high, p, skin);
  b.subdivide(low, high, low, high, p, skin);
  const int n = 5000;
  std::vector<double> xs(3 * n), fa(3 * n), fb(3 * n), u(n);
  std::vector<double*> xr(n), far(n), fbr(n);
  std::vector<int> mask(n);
  for (int i = 0; i < n; i++) {
    xr[i] = &xs[3 * i];
    far[i] = &fa[3 * i];
    fbr[i] = &fb[3 * i];
  }
  for (int step = 0; step < 3; step++) {
    for (int i = 0; i < n; i++) {
      for (int d = 0; d < 3; d++) xs[3 * i + d] = 8.0 * rand() / RAND_MAX;
      fa[3 * i] = fa[3 * i + 1] = fb[3 * i] = fb[3 * i + 1] = 0.0;
      u[i] = (double)rand() / RAND_MAX;
      mask[i] = rand() % 4;
    }
    a.set_mask(&mask[0]);
    b.set_mask(&mask[0]);
    double ea = a.update_forces_add_hills(n, &xr[0], &far[0], &u[0], 2, 1);
    double eb = b.update_forces(n, &xr[0], &fbr[0], 2);
    b.add_hills(n, &xr[0], &u[0], 2);
    REQUIRE(fabs(ea - eb) <= 1e-12 * fabs(eb));
    bool same = true;
    for (int i = 0; i < 3 * n; i++) same = same && fa[i] == fb[i];
    REQUIRE(same);
    REQUIRE(a.cum_bias_ == b.cum_bias_);
  }
  REQUIRE(a.cum_bias_ > 0);
  double x[] = {4.0, 4.0};
  REQUIRE(a.bias_->get_value(x) == b.bias_->get_value(x));
}

int main(int argc, char** argv) {
  g_src = argc > 1 ? argv[1] : "-";
  g_filter = argc > 2 ? argv[2] : NULL;
  srand(12345);
  if (wanted("grid_1d_sanity")) grid_1d_sanity();
  if (wanted("grid_3d_sanity")) grid_3d_sanity();
  if (wanted("grid_reads")) grid_reads();
  if (wanted("grid_read_write_consistency")) grid_read_write_consistency();
  if (wanted("interpolation_1d")) interpolation_1d();
  if (wanted("interp_1d_periodic")) interp_1d_periodic();
  if (wanted("boundary_remap")) boundary_remap();
  if (wanted("interp_3d_mixed")) interp_3d_mixed();
  if (wanted("gauss_grid_add_check")) gauss_grid_add_check();
  if (wanted("gauss_pbc_checks")) gauss_pbc_checks();
  if (wanted("gauss_grid_integral_tests")) gauss_grid_integral_tests();
  if (wanted("gauss_grid_derivative_tests")) gauss_grid_derivative_tests();
  if (wanted("gauss_grid_interp_test_mcgdp")) gauss_grid_interp_test_mcgdp();
  if (wanted("gauss_grid_integral_regression_1")) gauss_grid_integral_regression_1();
  if (wanted("edm_bias_reader")) edm_bias_reader();
  if (wanted("edm_sanity")) edm_sanity();
  if (wanted("notebook_vector")) notebook_vector();
  if (wanted("fused_step_matches_two_calls")) fused_step_matches_two_calls();
  printf("%d checks, %d failed\n", g_checks, g_failed);
  return g_failed ? 1 : 0;
}
