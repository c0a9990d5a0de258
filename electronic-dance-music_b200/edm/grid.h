// EDM::Grid / EDM::DimmedGrid<DIM> — the reference's grid interface (lib/grid.h:142-182, 184-905)
// over a grid that lives in B200 HBM.  All arithmetic (interpolation, indexing, histogram bumps,
// Grid::add) runs in the CUDA library behind include/edm_b200.h; this class keeps the geometry,
// the PLUMED-1 text I/O (lib/grid.h:448-503, 712-835) and a host mirror for code that pokes
// grid_ / grid_deriv_ directly, as the reference's tests and Python binding do.
#ifndef EDM_B200_GRID_H_
#define EDM_B200_GRID_H_

#include <cstddef>
#include <string>
#include <vector>

#include "../../include/edm_b200.h"
#include "edm.h"

#define GRID_TYPE 32

namespace EDM {

class Grid {  // lib/grid.h:142-182
 public:
  virtual ~Grid() {}
  virtual double get_value(const double* x) const = 0;
  virtual double add_value(const double* x0, double value) = 0;
  virtual double get_value_deriv(const double* x, double* der) const = 0;
  virtual void write(const std::string& filename) const = 0;
  virtual void multi_write(const std::string& filename, const double* box_low, const double* box_high,
                           const int* b_periodic, int b_lammps_format) const = 0;
  virtual void read(const std::string& filename) = 0;
  virtual void set_interpolation(int b_interpolate) = 0;
  virtual double* get_grid() = 0;
  virtual const double* get_dx() const = 0;
  virtual const double* get_max() const = 0;
  virtual const double* get_min() const = 0;
  virtual double max_value() const = 0;
  virtual double min_value() const = 0;
  virtual void add(const Grid* other, double scale, double offset) = 0;
  virtual size_t get_grid_size() const = 0;
  virtual void one2multi(size_t index, size_t* result) const = 0;
  virtual double expected_bias() const = 0;
  virtual void clear() = 0;
  // B200 additions: the device object behind this grid, and batched evaluation on host buffers
  virtual edm_grid_t* device_grid() const = 0;
  virtual void get_value_deriv_batch(long n, const double* x, long xstride, double* value, double* der) const = 0;
  // tells the host mirror that a kernel launched elsewhere (EDMBias) rewrote the device copy
  virtual void device_changed() const = 0;
};

// Host view of a device array that keeps itself coherent: reading or writing an element pulls the
// device copy if a kernel changed it and marks the host copy as the newer one; device operations
// push it back first.  This is what makes `g.grid_[i] = v; g.get_value(x)` behave as in the
// reference although the numbers live in HBM.
class GridStore;
class HostArray {
 public:
  HostArray() : store_(nullptr), which_(0) {}
  double& operator[](size_t i);
  const double& operator[](size_t i) const;
  operator double*();
  bool operator==(const void* p) const { return p == nullptr && store_ == nullptr; }
  bool operator!=(const void* p) const { return !(*this == p); }

 private:
  friend class GridStore;
  GridStore* store_;
  int which_;  // 0 values, 1 derivatives
};

class GridStore {
 public:
  GridStore() : dev(nullptr), dim(0), size(0), host_newer(false), device_newer(false) {}
  void attach(edm_grid_t* g, int dim_, size_t size_, HostArray* values, HostArray* derivs);
  void to_device() const;   // push the host mirror if it is the newer copy
  void to_host() const;     // pull the device copy if it is the newer one
  void device_changed() const { device_newer = true; }
  edm_grid_t* dev;
  int dim;
  size_t size;
  mutable std::vector<double> values, derivs;
  mutable bool host_newer, device_newer;
};

template <int D> class DimmedGaussGrid;

template <unsigned int DIM>
class DimmedGrid : public Grid {
 public:
  DimmedGrid(const double* min, const double* max, const double* bin_spacing, const int* b_periodic,
             int b_derivatives, int b_interpolate);
  DimmedGrid(const std::string& input_grid, int b_interpolate);
  explicit DimmedGrid(const std::string& input_grid);
  DimmedGrid(const DimmedGrid<DIM>& other);
  ~DimmedGrid();

  void get_index(const double* x, size_t result[DIM]) const;
  size_t multi2one(const size_t index[DIM]) const;
  void one2multi(size_t index, size_t result[DIM]) const;
  int in_grid(const double x[DIM]) const;

  double get_value(const double* x) const;
  double add_value(const double* x0, double value);
  double get_value_deriv(const double* x, double* der) const;
  void get_value_deriv_batch(long n, const double* x, long xstride, double* value, double* der) const;
  void add(const Grid* other, double scale, double offset);
  double max_value() const;
  double min_value() const;
  void write(const std::string& filename) const;
  void multi_write(const std::string& filename, const double* box_low, const double* box_high,
                   const int* b_periodic, int b_lammps_format) const;
  void read(const std::string& filename);
  void clear();
  double expected_bias() const;
  void set_interpolation(int b_interpolate);
  double* get_grid();
  const double* get_dx() const { return dx_; }
  const double* get_min() const { return min_; }
  const double* get_max() const { return max_; }
  size_t get_grid_size() const { return grid_size_; }
  edm_grid_t* device_grid() const {
    store_.to_device();
    return dev_;
  }
  // marks the device copy as changed by a kernel the caller launched through device_grid()
  void device_changed() const { store_.device_changed(); }

  // public state, lib/grid.h:876-885
  size_t grid_size_;
  int b_derivatives_;
  int b_interpolate_;
  HostArray grid_;
  HostArray grid_deriv_;
  double dx_[DIM];
  double min_[DIM];
  double max_[DIM];
  int grid_number_[DIM];
  int b_periodic_[DIM];

 protected:
  template <int D> friend class DimmedGaussGrid;
  // used by DimmedGaussGrid, whose device object is created by edm_gauss_create
  DimmedGrid() : grid_size_(0), b_derivatives_(0), b_interpolate_(0), dev_(nullptr), owns_(true) {}
  void adopt(edm_grid_t* g);
  void release();
  edm_grid_t* dev_;
  bool owns_;
  mutable GridStore store_;

 private:
  void parse_file(const std::string& filename, int interpolate_flag);
};

Grid* make_grid(unsigned int dim, const double* min, const double* max, const double* bin_spacing,
                const int* b_periodic, int b_derivatives, int b_interpolate);
Grid* read_grid(unsigned int dim, const std::string& filename, int b_interpolate);
Grid* read_grid(unsigned int dim, const std::string& filename);

// device ordinal used for grids created through the factories (default 0; EDM_B200_DEVICE overrides)
int default_device();
void set_default_device(int device);

}  // namespace EDM
#endif  // EDM_B200_GRID_H_
