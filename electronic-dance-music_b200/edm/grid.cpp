// Host side of EDM::DimmedGrid<DIM>: geometry bookkeeping, PLUMED-1 text I/O and the coherent host
// mirror.  Every numerical operation is a call into the CUDA library (include/edm_b200.h).
#include "grid.h"

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>

namespace EDM {

static int g_default_device = -1;
int default_device() {
  if (g_default_device < 0) {
    const char* e = getenv("EDM_B200_DEVICE");
    g_default_device = e ? atoi(e) : 0;
  }
  return g_default_device;
}
void set_default_device(int device) { g_default_device = device; }

// ------------------------------------------------------------------ coherent host mirror

void GridStore::attach(edm_grid_t* g, int dim_, size_t size_, HostArray* v, HostArray* d) {
  dev = g;
  dim = dim_;
  size = size_;
  values.assign(size, 0.0);
  derivs.assign(size * (size_t)dim, 0.0);
  host_newer = false;
  device_newer = false;
  v->store_ = this;
  v->which_ = 0;
  d->store_ = this;
  d->which_ = 1;
}

void GridStore::to_device() const {
  if (!host_newer || !dev) return;
  edm_check(edm_grid_upload(dev, values.data(), derivs.data()), "grid.cpp:GridStore::to_device");
  host_newer = false;
}

void GridStore::to_host() const {
  if (!device_newer || !dev) return;
  edm_check(edm_grid_download(dev, values.data(), derivs.data()), "grid.cpp:GridStore::to_host");
  device_newer = false;
}

double& HostArray::operator[](size_t i) {
  store_->to_host();
  store_->host_newer = true;  // a non-const element may be written through the reference
  return which_ == 0 ? store_->values[i] : store_->derivs[i];
}
const double& HostArray::operator[](size_t i) const {
  store_->to_host();
  return which_ == 0 ? store_->values[i] : store_->derivs[i];
}
HostArray::operator double*() {
  store_->to_host();
  store_->host_newer = true;
  return which_ == 0 ? store_->values.data() : store_->derivs.data();
}

// ------------------------------------------------------------------ construction

template <unsigned int DIM> void DimmedGrid<DIM>::adopt(edm_grid_t* g) {
  dev_ = g;
  int dim = 0;
  size_t size = 0;
  edm_check(edm_grid_geometry(g, &dim, grid_number_, dx_, min_, max_, b_periodic_, nullptr, &size),
            "grid.cpp:DimmedGrid::adopt");
  if ((unsigned)dim != DIM) edm_error("Dimension of this grid does not match the device grid", "grid.cpp:adopt");
  grid_size_ = size;
  edm_check(edm_grid_flags(g, &b_derivatives_, &b_interpolate_, nullptr), "grid.cpp:DimmedGrid::adopt");
  store_.attach(g, (int)DIM, size, &grid_, &grid_deriv_);
}

template <unsigned int DIM> void DimmedGrid<DIM>::release() {
  if (dev_ && owns_) edm_grid_destroy(dev_);
  dev_ = nullptr;
}

template <unsigned int DIM>
DimmedGrid<DIM>::DimmedGrid(const double* min, const double* max, const double* bin_spacing, const int* b_periodic,
                            int b_derivatives, int b_interpolate)
    : dev_(nullptr), owns_(true) {
  edm_grid_t* g = nullptr;
  edm_check(edm_grid_create(&g, default_device(), (int)DIM, min, max, bin_spacing, b_periodic, b_derivatives,
                            b_interpolate),
            "grid.h:DimmedGrid");
  adopt(g);
}

template <unsigned int DIM>
DimmedGrid<DIM>::DimmedGrid(const std::string& input_grid, int b_interpolate) : dev_(nullptr), owns_(true) {
  parse_file(input_grid, b_interpolate);
}

template <unsigned int DIM>
DimmedGrid<DIM>::DimmedGrid(const std::string& input_grid) : dev_(nullptr), owns_(true) {
  parse_file(input_grid, 1);
}

template <unsigned int DIM> DimmedGrid<DIM>::DimmedGrid(const DimmedGrid<DIM>& other) : dev_(nullptr), owns_(true) {
  // clone: same geometry, copy of the contents (lib/grid.h:233-251)
  int bins[DIM];
  double mx[DIM];
  for (unsigned i = 0; i < DIM; i++) {
    bins[i] = other.b_periodic_[i] ? other.grid_number_[i] : other.grid_number_[i] - 1;
    mx[i] = other.b_periodic_[i] ? other.max_[i] : other.max_[i] - other.dx_[i];
  }
  edm_grid_t* g = nullptr;
  edm_check(edm_grid_create_from_header(&g, default_device(), (int)DIM, bins, other.min_, mx, other.b_periodic_,
                                        other.b_derivatives_, other.b_interpolate_),
            "grid.h:DimmedGrid(clone)");
  adopt(g);
  other.store_.to_host();
  other.store_.to_device();
  store_.values = other.store_.values;
  store_.derivs = other.store_.derivs;
  store_.host_newer = true;
}

template <unsigned int DIM> DimmedGrid<DIM>::~DimmedGrid() { release(); }

// ------------------------------------------------------------------ index helpers (host, geometry only)

static int host_int_floor(double number) {  // lib/grid.h:17-20
  return (int)((int)number < 0.0 ? -ceil(fabs(number)) : floor(number));
}

template <unsigned int DIM> void DimmedGrid<DIM>::get_index(const double* x, size_t result[DIM]) const {
  for (unsigned i = 0; i < DIM; i++) {  // lib/grid.h:264-273
    double xi = x[i];
    if (b_periodic_[i]) xi -= (max_[i] - min_[i]) * host_int_floor((xi - min_[i]) / (max_[i] - min_[i]));
    result[i] = (size_t)floor((xi - min_[i]) / dx_[i]);
  }
}

template <unsigned int DIM> size_t DimmedGrid<DIM>::multi2one(const size_t index[DIM]) const {
  size_t r = index[DIM - 1];  // dim 0 fastest, lib/grid.h:315-325
  for (unsigned i = DIM - 1; i > 0; i--) r = r * grid_number_[i - 1] + index[i - 1];
  return r;
}

template <unsigned int DIM> void DimmedGrid<DIM>::one2multi(size_t index, size_t result[DIM]) const {
  unsigned i;
  for (i = 0; i < DIM - 1; i++) {  // lib/grid.h:330-338
    result[i] = index % grid_number_[i];
    index = (index - result[i]) / grid_number_[i];
  }
  result[i] = index;
}

template <unsigned int DIM> int DimmedGrid<DIM>::in_grid(const double x[DIM]) const {
  for (unsigned i = 0; i < DIM; i++)  // lib/grid.h:865-874
    if (!b_periodic_[i] && (x[i] < min_[i] || x[i] >= max_[i] - dx_[i])) return 0;
  return 1;
}

// ------------------------------------------------------------------ numerical entry points -> device

template <unsigned int DIM> double DimmedGrid<DIM>::get_value(const double* x) const {
  store_.to_device();
  double v = 0;
  edm_check(edm_grid_get_value(dev_, 1, x, (long)DIM, &v), "grid.h:get_value");
  return v;
}

template <unsigned int DIM> double DimmedGrid<DIM>::get_value_deriv(const double* x, double* der) const {
  store_.to_device();
  double v = 0;
  edm_check(edm_grid_eval(dev_, 1, x, (long)DIM, &v, der), "grid.h:get_value_deriv");
  return v;
}

template <unsigned int DIM>
void DimmedGrid<DIM>::get_value_deriv_batch(long n, const double* x, long xstride, double* value, double* der) const {
  store_.to_device();
  edm_check(edm_grid_eval(dev_, n, x, xstride, value, der), "grid.h:get_value_deriv_batch");
}

template <unsigned int DIM> double DimmedGrid<DIM>::add_value(const double* x0, double value) {
  if (b_interpolate_) edm_error("Cannot add_value when using derivatives", "grid.h:add_value");
  if (!in_grid(x0)) return 0;
  store_.to_device();
  edm_check(edm_grid_hist_add(dev_, 1, x0, (long)DIM, &value), "grid.h:add_value");
  store_.device_changed();
  return value;
}

template <unsigned int DIM> void DimmedGrid<DIM>::add(const Grid* other, double scale, double offset) {
  store_.to_device();
  edm_check(edm_grid_add(dev_, other->device_grid(), scale, offset), "grid.h:add");
  store_.device_changed();
}

template <unsigned int DIM> double DimmedGrid<DIM>::max_value() const {
  store_.to_device();
  double mn, mx;
  edm_check(edm_grid_minmax(dev_, &mn, &mx), "grid.h:max_value");
  return mx;
}

template <unsigned int DIM> double DimmedGrid<DIM>::min_value() const {
  store_.to_device();
  double mn, mx;
  edm_check(edm_grid_minmax(dev_, &mn, &mx), "grid.h:min_value");
  return mn;
}

template <unsigned int DIM> void DimmedGrid<DIM>::clear() {
  edm_check(edm_grid_clear(dev_), "grid.h:clear");
  store_.host_newer = false;
  store_.device_changed();
}

template <unsigned int DIM> void DimmedGrid<DIM>::set_interpolation(int b_interpolate) {
  b_interpolate_ = b_interpolate;
  edm_check(edm_grid_set_interpolation(dev_, b_interpolate), "grid.h:set_interpolation");
}

template <unsigned int DIM> double* DimmedGrid<DIM>::get_grid() { return (double*)grid_; }

// Setup-time scalar (lib/grid.h:692-710): target grids are read once, so this walks the host copy
// with libm; the result (expected_target_) then enters every hill height on the device.
template <unsigned int DIM> double DimmedGrid<DIM>::expected_bias() const {
  store_.to_host();
  const std::vector<double>& g = store_.values;
  double Z = 0, offset = 0, avg = 0;
  for (size_t i = 0; i < grid_size_; i++) offset = fmax(offset, g[i]);
  for (size_t i = 0; i < grid_size_; i++) Z += exp(-g[i] - offset);
  for (size_t i = 0; i < grid_size_; i++) avg += g[i] * exp(-g[i] - offset);
  return avg / Z;
}

// ------------------------------------------------------------------ PLUMED-1 text I/O

template <unsigned int DIM> void DimmedGrid<DIM>::write(const std::string& filename) const {
  using namespace std;
  store_.to_host();
  ofstream out(filename.c_str());
  out << "#! FORCE " << b_derivatives_ << endl;
  out << "#! NVAR " << DIM << endl;
  out << "#! TYPE ";
  for (unsigned i = 0; i < DIM; i++) out << GRID_TYPE << " ";
  out << endl << "#! BIN ";
  for (unsigned i = 0; i < DIM; i++) out << (b_periodic_[i] ? grid_number_[i] : grid_number_[i] - 1) << " ";
  out << endl << "#! MIN ";
  for (unsigned i = 0; i < DIM; i++) out << min_[i] << " ";
  out << endl << "#! MAX ";
  for (unsigned i = 0; i < DIM; i++) out << (b_periodic_[i] ? max_[i] : max_[i] - dx_[i]) << " ";
  out << endl << "#! PBC ";
  for (unsigned i = 0; i < DIM; i++) out << b_periodic_[i] << " ";
  out << endl;
  size_t idx[DIM];
  out << setprecision(8) << std::fixed;
  for (size_t p = 0; p < grid_size_; p++) {
    one2multi(p, idx);
    for (unsigned j = 0; j < DIM; j++) out << (min_[j] + dx_[j] * idx[j]) << " ";
    out << store_.values[p] << " ";
    if (b_derivatives_)
      for (unsigned j = 0; j < DIM; j++) out << -store_.derivs[p * DIM + j] << " ";  // files hold forces
    out << endl;
    if (idx[0] == (size_t)(grid_number_[0] - 1)) out << endl;
  }
}

template <unsigned int DIM> void DimmedGrid<DIM>::parse_file(const std::string& filename, int interpolate_flag) {
  using namespace std;
  ifstream in(filename.c_str());
  if (!in.is_open()) {
    cerr << "Cannot open input file \"" << filename << "\"" << endl;
    edm_error("", "grid.h:read");
  }
  string hash, word;
  int force = 0, nvar = 0, bins[DIM], pbc[DIM];
  double mn[DIM], mx[DIM];
  auto expect = [&](const char* key) {
    in >> hash >> word;
    if (word.compare(key) != 0) {
      cerr << "Mangled grid file: " << filename << " No " << key << " found" << endl;
      edm_error("", "grid.h:read");
    }
  };
  expect("FORCE");
  in >> force;
  expect("NVAR");
  in >> nvar;
  if ((unsigned)nvar != DIM) {
    cerr << "Dimension of this grid does not match the one found in the file" << endl;
    edm_error("", "grid.h:read");
  }
  expect("TYPE");
  for (unsigned i = 0; i < DIM; i++) {
    int t;
    in >> t;
    if (t != GRID_TYPE) cerr << "WARNING: Read grid type is the incorrect type" << endl;
  }
  expect("BIN");
  for (unsigned i = 0; i < DIM; i++) in >> bins[i];
  expect("MIN");
  for (unsigned i = 0; i < DIM; i++) in >> mn[i];
  expect("MAX");
  for (unsigned i = 0; i < DIM; i++) in >> mx[i];
  expect("PBC");
  for (unsigned i = 0; i < DIM; i++) in >> pbc[i];

  release();
  edm_grid_t* g = nullptr;
  edm_check(edm_grid_create_from_header(&g, default_device(), (int)DIM, bins, mn, mx, pbc, force, interpolate_flag),
            "grid.h:read");
  owns_ = true;
  adopt(g);
  for (size_t p = 0; p < grid_size_; p++) {
    for (unsigned j = 0; j < DIM; j++) in >> word;  // coordinates are implied by the header
    in >> store_.values[p];
    if (force)
      for (unsigned j = 0; j < DIM; j++) {
        in >> store_.derivs[p * DIM + j];
        store_.derivs[p * DIM + j] *= -1;  // stored negated, lib/grid.h:828
      }
  }
  store_.host_newer = true;
  store_.to_device();
}

template <unsigned int DIM> void DimmedGrid<DIM>::read(const std::string& filename) {
  parse_file(filename, b_interpolate_);
}

// Serial rendering of the reference's MPI gather-write (lib/grid.h:509-674): with a replicated grid
// one process holds everything, so the per-point collectives collapse into one batched evaluation.
template <unsigned int DIM>
void DimmedGrid<DIM>::multi_write(const std::string& filename, const double* box_min, const double* box_max,
                                  const int* b_periodic, int b_lammps_format) const {
  using namespace std;
  if (b_lammps_format == 1 && DIM > 1) edm_error("Lammps format only valid for 1D grids", "grid.h:multi_write");
  unsigned int counts[DIM];
  size_t total = 1;
  for (unsigned i = 0; i < DIM; i++) {
    counts[i] = (int)ceil((box_max[i] - box_min[i]) / dx_[i]);
    counts[i] = b_periodic[i] ? counts[i] : counts[i] + 1;
    total *= counts[i];
  }
  unsigned int extra_n = b_lammps_format ? (unsigned int)(box_min[0] / dx_[0]) : 0;
  vector<double> pts(total * DIM), val(total), der(total * DIM);
  vector<size_t> first(total);
  for (size_t p = 0; p < total; p++) {
    size_t t = p;
    for (unsigned j = 0; j < DIM; j++) {
      size_t k = (j < DIM - 1) ? t % counts[j] : t;
      t = (t - k) / counts[j];
      if (j == 0) first[p] = k;
      pts[p * DIM + j] = k * dx_[j] + box_min[j];
    }
  }
  // the reference evaluates DimmedGrid::get_value_deriv here (lib/grid.h:650-653): the plain grid, even when this
  // grid sits inside a GaussGrid whose boundary test would return 0 at the upper wall
  store_.to_device();
  edm_check(edm_grid_eval_plain(dev_, (long)total, pts.data(), (long)DIM, val.data(), der.data()), "grid.h:multi_write");
  ofstream out(filename.c_str());
  if (!b_lammps_format) {
    out << "#! FORCE " << b_derivatives_ << endl << "#! NVAR " << DIM << endl << "#! TYPE ";
    for (unsigned i = 0; i < DIM; i++) out << GRID_TYPE << " ";
    out << endl << "#! BIN ";
    for (unsigned i = 0; i < DIM; i++) out << (b_periodic[i] ? counts[i] : counts[i] - 1) << " ";
    out << endl << "#! MIN ";
    for (unsigned i = 0; i < DIM; i++) out << box_min[i] << " ";
    out << endl << "#! MAX ";
    for (unsigned i = 0; i < DIM; i++) out << box_max[i] << " ";
    out << endl << "#! PBC ";
    for (unsigned i = 0; i < DIM; i++) out << b_periodic[i] << " ";
    out << endl;
  } else {
    out << "#Auto generated by electronic-dance-music" << endl << endl << "EDM" << endl;
    out << "N " << extra_n + counts[0] << " R " << dx_[0] << " " << box_max[0] << endl << endl;
    for (unsigned int i = 1; i < extra_n; i++) out << i << " " << i * dx_[0] << " 0.0" << " 0.0" << endl;
  }
  out << setprecision(8) << std::fixed;
  for (size_t p = 0; p < total; p++) {
    if (!in_grid(&pts[p * DIM])) continue;
    if (b_lammps_format) out << p + extra_n << " ";
    for (unsigned j = 0; j < DIM; j++) out << pts[p * DIM + j] << " ";
    out << val[p] << " ";
    if (b_derivatives_)
      for (unsigned j = 0; j < DIM; j++) out << -der[p * DIM + j] << " ";
    out << endl;
    if (first[p] == counts[0] - 1) out << endl;
  }
}

template class DimmedGrid<1>;
template class DimmedGrid<2>;
template class DimmedGrid<3>;

Grid* make_grid(unsigned int dim, const double* min, const double* max, const double* bin_spacing,
                const int* b_periodic, int b_derivatives, int b_interpolate) {
  if (dim == 1) return new DimmedGrid<1>(min, max, bin_spacing, b_periodic, b_derivatives, b_interpolate);
  if (dim == 2) return new DimmedGrid<2>(min, max, bin_spacing, b_periodic, b_derivatives, b_interpolate);
  if (dim == 3) return new DimmedGrid<3>(min, max, bin_spacing, b_periodic, b_derivatives, b_interpolate);
  return NULL;
}

Grid* read_grid(unsigned int dim, const std::string& filename, int b_interpolate) {
  if (dim == 1) return new DimmedGrid<1>(filename, b_interpolate);
  if (dim == 2) return new DimmedGrid<2>(filename, b_interpolate);
  if (dim == 3) return new DimmedGrid<3>(filename, b_interpolate);
  return NULL;
}

Grid* read_grid(unsigned int dim, const std::string& filename) {
  if (dim == 1) return new DimmedGrid<1>(filename);
  if (dim == 2) return new DimmedGrid<2>(filename);
  if (dim == 3) return new DimmedGrid<3>(filename);
  return NULL;
}

}  // namespace EDM
