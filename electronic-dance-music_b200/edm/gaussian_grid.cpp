#include "gaussian_grid.h"

#include <cmath>
#include <vector>

namespace EDM {

template <int DIM> void DimmedGaussGrid<DIM>::bind(edm_grid_t* g) {
  grid_.adopt(g);
  int mini[DIM];
  edm_check(edm_grid_geometry(g, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, mini, nullptr),
            "gaussian_grid.h:bind");
  minisize_total_ = 1;
  for (int i = 0; i < DIM; i++) {
    minisize_[i] = (size_t)mini[i];
    minisize_total_ *= (2 * minisize_[i] + 1);
  }
  refresh_boundary();
}

template <int DIM> void DimmedGaussGrid<DIM>::refresh_boundary() {
  edm_check(edm_grid_boundary(grid_.dev_, boundary_min_, boundary_max_, b_periodic_boundary_, sigma_),
            "gaussian_grid.h:set_boundary");
}

template <int DIM>
DimmedGaussGrid<DIM>::DimmedGaussGrid(const double* min, const double* max, const double* bin_spacing,
                                      const int* b_periodic, int b_interpolate, const double* sigma) {
  edm_grid_t* g = nullptr;
  edm_check(edm_gauss_create(&g, default_device(), DIM, min, max, bin_spacing, b_periodic, b_interpolate, sigma),
            "gaussian_grid.h:DimmedGaussGrid");
  bind(g);
}

// Rebuild from a grid file; files do not store sigma (lib/gaussian_grid.h:85-93).  The spacing of
// the file's lattice is recovered from its header, the contents are uploaded afterwards.
template <int DIM> DimmedGaussGrid<DIM>::DimmedGaussGrid(const std::string& filename, const double* sigma) {
  DimmedGrid<DIM> file(filename);
  double mx[DIM], spacing[DIM];
  for (int i = 0; i < DIM; i++) {
    mx[i] = file.b_periodic_[i] ? file.max_[i] : file.max_[i] - file.dx_[i];
    spacing[i] = file.dx_[i] * (1.0 + 1e-12);  // ceil(L / spacing) then lands on the file's bin count
  }
  edm_grid_t* g = nullptr;
  edm_check(edm_gauss_create(&g, default_device(), DIM, file.min_, mx, spacing, file.b_periodic_, 1, sigma),
            "gaussian_grid.h:DimmedGaussGrid(file)");
  bind(g);
  if (grid_.grid_size_ != file.grid_size_) edm_error("grid file lattice does not reproduce", "gaussian_grid.h:read");
  std::vector<double> v(file.grid_size_), d(file.grid_size_ * DIM);
  edm_check(edm_grid_download(file.device_grid(), v.data(), d.data()), "gaussian_grid.h:read");
  edm_check(edm_grid_upload(grid_.dev_, v.data(), d.data()), "gaussian_grid.h:read");
  grid_.device_changed();
}

template <int DIM> DimmedGaussGrid<DIM>::~DimmedGaussGrid() {}

template <int DIM> void DimmedGaussGrid<DIM>::read(const std::string& filename) {
  DimmedGrid<DIM> file(filename);
  if (file.grid_size_ != grid_.grid_size_) edm_error("grid file does not match this grid", "gaussian_grid.h:read");
  std::vector<double> v(file.grid_size_), d(file.grid_size_ * DIM);
  edm_check(edm_grid_download(file.device_grid(), v.data(), d.data()), "gaussian_grid.h:read");
  edm_check(edm_grid_upload(grid_.device_grid(), v.data(), d.data()), "gaussian_grid.h:read");
  grid_.device_changed();
}

template <int DIM> double DimmedGaussGrid<DIM>::get_value(const double* x) const { return grid_.get_value(x); }

template <int DIM> double DimmedGaussGrid<DIM>::get_value_deriv(const double* x, double* der) const {
  return grid_.get_value_deriv(x, der);
}

template <int DIM>
void DimmedGaussGrid<DIM>::get_value_deriv_batch(long n, const double* x, long xstride, double* value,
                                                 double* der) const {
  grid_.get_value_deriv_batch(n, x, xstride, value, der);
}

template <int DIM> double DimmedGaussGrid<DIM>::add_value(const double* x0, double height) {
  double ba = 0;
  add_values(1, x0, &height, &ba);
  return ba;
}

template <int DIM>
void DimmedGaussGrid<DIM>::add_values(long n, const double* x, const double* heights, double* bias_added) {
  edm_check(edm_gauss_deposit(grid_.device_grid(), n, x, heights, bias_added), "gaussian_grid.h:add_value");
  grid_.device_changed();
}

template <int DIM>
void DimmedGaussGrid<DIM>::set_boundary(const double* min, const double* max, const int* b_periodic) {
  edm_check(edm_grid_set_boundary(grid_.device_grid(), min, max, b_periodic), "gaussian_grid.h:set_boundary");
  refresh_boundary();
}

template <int DIM> double DimmedGaussGrid<DIM>::get_volume() const {
  double vol = 1;
  for (int i = 0; i < DIM; i++) vol *= boundary_max_[i] - boundary_min_[i];
  return vol;
}

template <int DIM> int DimmedGaussGrid<DIM>::in_bounds(const double x[DIM]) const {
  for (int i = 0; i < DIM; i++)
    if (x[i] < boundary_min_[i] || x[i] > boundary_max_[i]) return 0;
  return 1;
}

template <int DIM> void DimmedGaussGrid<DIM>::remap(double x[DIM]) const {
  edm_check(edm_grid_remap(grid_.device_grid(), x), "gaussian_grid.h:remap");
}

template <int DIM> void DimmedGaussGrid<DIM>::multi_write(const std::string& filename) const {
  grid_.multi_write(filename, boundary_min_, boundary_max_, b_periodic_boundary_, 0);
}
template <int DIM> void DimmedGaussGrid<DIM>::lammps_multi_write(const std::string& filename) const {
  grid_.multi_write(filename, boundary_min_, boundary_max_, b_periodic_boundary_, 1);
}
template <int DIM>
void DimmedGaussGrid<DIM>::multi_write(const std::string& filename, const double* box_low, const double* box_high,
                                       const int* b_periodic, int) const {
  grid_.multi_write(filename, box_low, box_high, b_periodic, 0);  // lib/gaussian_grid.h:160-166
}

template class DimmedGaussGrid<1>;
template class DimmedGaussGrid<2>;
template class DimmedGaussGrid<3>;

GaussGrid* make_gauss_grid(unsigned int dim, const double* min, const double* max, const double* bin_spacing,
                           const int* b_periodic, int b_interpolate, const double* sigma) {
  if (dim == 1) return new DimmedGaussGrid<1>(min, max, bin_spacing, b_periodic, b_interpolate, sigma);
  if (dim == 2) return new DimmedGaussGrid<2>(min, max, bin_spacing, b_periodic, b_interpolate, sigma);
  if (dim == 3) return new DimmedGaussGrid<3>(min, max, bin_spacing, b_periodic, b_interpolate, sigma);
  return NULL;
}

GaussGrid* read_gauss_grid(unsigned int dim, const std::string& filename, const double* sigma) {
  if (dim == 1) return new DimmedGaussGrid<1>(filename, sigma);
  if (dim == 2) return new DimmedGaussGrid<2>(filename, sigma);
  if (dim == 3) return new DimmedGaussGrid<3>(filename, sigma);
  return NULL;
}

}  // namespace EDM
