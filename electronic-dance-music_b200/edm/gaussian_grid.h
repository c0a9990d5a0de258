// EDM::GaussGrid / EDM::DimmedGaussGrid<DIM> — the reference's hill grid (lib/gaussian_grid.h:41-56,
// 58-631) over the device-resident GaussGrid of include/edm_b200.h.  Hill deposition (add_value,
// incl. McGovern-De Pablo and zero-force hills), boundary remap and interpolation all run on the
// GPU; the class keeps the public members the reference exposes and adds a batched add_values.
#ifndef EDM_B200_GAUSS_GRID_H_
#define EDM_B200_GAUSS_GRID_H_

#include <string>

#include "edm.h"
#include "grid.h"

#define GAUSS_SUPPORT 8.0
#define BC_TABLE_SIZE 65536
#define BC_MAR 2.0
#define BC_CORRECTION

namespace EDM {

class GaussGrid : public Grid {  // lib/gaussian_grid.h:41-56
 public:
  virtual ~GaussGrid() {}
  virtual double add_value(const double* x, double height) = 0;
  virtual void set_boundary(const double* min, const double* max, const int* b_periodic) = 0;
  virtual double get_volume() const = 0;
  virtual int in_bounds(const double* x) const = 0;
  virtual void multi_write(const std::string& filename) const = 0;
  virtual void lammps_multi_write(const std::string& filename) const = 0;
  using Grid::multi_write;
  // B200 addition: n hills in list order with one launch; bias_added may be NULL
  virtual void add_values(long n, const double* x, const double* heights, double* bias_added) = 0;
};

template <int DIM>
class DimmedGaussGrid : public GaussGrid {
 public:
  DimmedGaussGrid(const double* min, const double* max, const double* bin_spacing, const int* b_periodic,
                  int b_interpolate, const double* sigma);
  DimmedGaussGrid(const std::string& filename, const double* sigma);
  ~DimmedGaussGrid();

  double get_value(const double* x) const;
  double get_value_deriv(const double* x, double* der) const;
  void get_value_deriv_batch(long n, const double* x, long xstride, double* value, double* der) const;
  double add_value(const double* x0, double height);
  void add_values(long n, const double* x, const double* heights, double* bias_added);
  void set_boundary(const double* min, const double* max, const int* b_periodic);
  double get_volume() const;
  int in_bounds(const double x[DIM]) const;
  void remap(double x[DIM]) const;

  void read(const std::string& filename);
  void write(const std::string& filename) const { grid_.write(filename); }
  void multi_write(const std::string& filename) const;
  void lammps_multi_write(const std::string& filename) const;
  void multi_write(const std::string& filename, const double* box_low, const double* box_high, const int* b_periodic,
                   int b_lammps_format) const;
  void set_interpolation(int b_interpolate) { grid_.set_interpolation(b_interpolate); }
  void one2multi(size_t index, size_t result[DIM]) const { grid_.one2multi(index, result); }
  double* get_grid() { return grid_.get_grid(); }
  const double* get_dx() const { return grid_.get_dx(); }
  const double* get_min() const { return grid_.get_min(); }
  const double* get_max() const { return grid_.get_max(); }
  double max_value() const { return grid_.max_value(); }
  double min_value() const { return grid_.min_value(); }
  void add(const Grid* other, double scale, double offset) { grid_.add(other, scale, offset); }
  double expected_bias() const { return grid_.expected_bias(); }
  void clear() { grid_.clear(); }
  size_t get_grid_size() const { return grid_.get_grid_size(); }
  edm_grid_t* device_grid() const { return grid_.device_grid(); }
  void device_changed() const { grid_.device_changed(); }

  // public state, lib/gaussian_grid.h:544-550 (the McGDP tables stay inside the device object)
  size_t minisize_[DIM];
  size_t minisize_total_;
  double sigma_[DIM];  // sigma * sqrt(2)
  double boundary_min_[DIM];
  double boundary_max_[DIM];
  int b_periodic_boundary_[DIM];
  DimmedGrid<DIM> grid_;

 private:
  void bind(edm_grid_t* g);
  void refresh_boundary();
};

GaussGrid* make_gauss_grid(unsigned int dim, const double* min, const double* max, const double* bin_spacing,
                           const int* b_periodic, int b_interpolate, const double* sigma);
GaussGrid* read_gauss_grid(unsigned int dim, const std::string& filename, const double* sigma);

}  // namespace EDM
#endif  // EDM_B200_GAUSS_GRID_H_
