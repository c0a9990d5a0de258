// C entry points behind the Python class EDMBias_Py (electronic-dance-music_b200/python/edm_b200/compat.py).
//
// The reference exposes EDMBias to Python through boost::python (python/edm/edm_python.cxx:6-19,
// python/edm/edm_bias_py.cpp).  Boost.Python is a Python-2-era dependency; here the same nine methods
// are plain extern "C" functions over EDM::EDMBias and the Python side binds them with ctypes.
// One deliberate difference: the reference's set_box stores every periodic flag in b_periodic[3]
// (python/edm/edm_bias_py.cpp:43, one past the array) and passes the array on uninitialised; this
// shim passes the flags it was given.
#include <cstdlib>
#include <string>

#include "edm_bias.h"

using EDM::EDMBias;

extern "C" {

// EDMBias_Py::EDMBias_Py, python/edm/edm_bias_py.cpp:20-28
void* edm_py_new(const char* input_filename, double temperature, double boltzmann_constant) {
  EDMBias* b = new EDMBias(input_filename);
  b->setup(temperature, boltzmann_constant);
  return b;
}

void edm_py_delete(void* h) { delete static_cast<EDMBias*>(h); }

int edm_py_dim(void* h) { return (int)static_cast<EDMBias*>(h)->dim_; }

// subdivide_py, python/edm/edm_bias_py.cpp:32-52: sub-box = box, no skin
void edm_py_set_box(void* h, int n, const double* boxlo, const double* boxhi, const int* periodic) {
  double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, skin[3] = {0, 0, 0};
  int per[3] = {0, 0, 0};
  for (int i = 0; i < n && i < 3; i++) {
    lo[i] = boxlo[i];
    hi[i] = boxhi[i];
    per[i] = periodic[i];
  }
  static_cast<EDMBias*>(h)->subdivide(lo, hi, lo, hi, per, skin);
}

void edm_py_pre_add_hill(void* h, int est_hill_count) { static_cast<EDMBias*>(h)->pre_add_hill(est_hill_count); }
void edm_py_post_add_hill(void* h) { static_cast<EDMBias*>(h)->post_add_hill(); }
// add_hill_py, python/edm/edm_bias_py.cpp:55-64
void edm_py_add_hill(void* h, const double* position, double runiform) {
  static_cast<EDMBias*>(h)->add_hill(position, runiform);
}
void edm_py_write_bias(void* h, const char* output) { static_cast<EDMBias*>(h)->write_bias(output); }
void edm_py_write_lammps_table(void* h, const char* output) { static_cast<EDMBias*>(h)->write_lammps_table(output); }
void edm_py_write_histogram(void* h) { static_cast<EDMBias*>(h)->write_histogram(); }
void edm_py_clear_histogram(void* h) { static_cast<EDMBias*>(h)->clear_histogram(); }

// get_force_py, python/edm/edm_bias_py.cpp:67-84: (bias energy, +dV/dx) at one point
double edm_py_get_force(void* h, const double* position, double* force) {
  EDMBias* b = static_cast<EDMBias*>(h);
  for (unsigned i = 0; i < b->dim_; i++) force[i] = 0;
  return b->bias_->get_value_deriv(position, force);
}

}  // extern "C"
