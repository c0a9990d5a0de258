#include "edm_bias.h"

#include <cmath>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <sstream>

namespace EDM {

namespace {

// The edm input format (lib/edm_bias.cpp:19-31, 997-1004): one `key value...`
// per line, the first occurrence of a key wins.
typedef std::map<std::string, std::string> KeyValues;

KeyValues parse_key_values(std::istream& in) {
  KeyValues kv;
  std::string key, rest;
  while (in >> key) {
    std::getline(in, rest);
    if (kv.find(key) == kv.end()) kv[key] = rest;
  }
  return kv;
}

bool get_int(const KeyValues& kv, const std::string& key, bool required, int* out) {
  KeyValues::const_iterator it = kv.find(key);
  if (it == kv.end()) {
    if (required) std::cerr << "Could not find key " << key << std::endl;
    return false;
  }
  *out = atoi(it->second.c_str());
  return true;
}

// a parsed 0.0 counts as invalid, as in the reference (lib/edm_bias.cpp:937-940, T16)
bool get_double(const KeyValues& kv, const std::string& key, bool required, double* out) {
  KeyValues::const_iterator it = kv.find(key);
  if (it == kv.end()) {
    if (required) std::cerr << "Could not find key " << key << std::endl;
    return false;
  }
  *out = atof(it->second.c_str());
  if (*out == 0.0) {
    std::cerr << "Invalid value found for " << key << std::endl;
    return false;
  }
  return true;
}

bool get_doubles(const KeyValues& kv, const std::string& key, bool required, double* out, int n) {
  KeyValues::const_iterator it = kv.find(key);
  if (it == kv.end()) {
    if (required) std::cerr << "Could not find key " << key << std::endl;
    return false;
  }
  std::istringstream is(it->second);
  for (int i = 0; i < n; i++) is >> out[i];
  return true;
}

}  // namespace

EDMBias::EDMBias(const std::string& input_filename)
    : b_tempering_(0), b_targeting_(0), mpi_rank_(0), mpi_size_(0), dim_(0), global_tempering_(0), bias_factor_(0),
      boltzmann_factor_(0), temperature_(-1.0), hill_prefactor_(0), bias_per_step_(0), hill_density_(-1),
      cum_bias_(0), total_volume_(0), expected_target_(0), b_outofbounds_(0), bias_dx_(NULL), bias_sigma_(NULL),
      min_(NULL), max_(NULL), b_periodic_boundary_(NULL), target_(NULL), initial_bias_(NULL), bias_(NULL), mask_(NULL),
      dev_(NULL), comm_(NULL), comm_cap_(0), cv_hist_(NULL), est_hill_count_(0), in_round_(0), steps_(0) {
  read_input(input_filename);  // a failed parse leaves a half-initialised object, as in the reference
}

EDMBias::~EDMBias() {
  if (dev_) edm_bias_destroy(dev_);
  delete target_;
  delete bias_;
  delete cv_hist_;
  delete initial_bias_;
  free(bias_dx_);
  free(bias_sigma_);
  free(min_);
  free(max_);
  free(b_periodic_boundary_);
}

std::string EDMBias::clean_string(const std::string& input, int append_rank) {
  std::string result(input);
  size_t found = result.find_first_not_of(" \t");
  if (found != std::string::npos) result = result.substr(found);
  if (append_rank) {
    std::ostringstream oss;
    oss << result << "_" << mpi_rank_;
    return oss.str();
  }
  return result;
}

int EDMBias::read_input(const std::string& input_filename) {
  std::ifstream input(input_filename.c_str());
  if (!input.is_open()) {
    std::cerr << "Cannot open input file " << input_filename << std::endl;
    return 0;
  }
  KeyValues kv = parse_key_values(input);
  if (!get_int(kv, "tempering", true, &b_tempering_)) {
    std::cerr << "Must specify if tempering is enabled, ex: tempering 1 or tempering 0" << std::endl;
    return 0;
  }
  if (b_tempering_) {
    if (!get_double(kv, "bias_factor", true, &bias_factor_)) return 0;
    get_double(kv, "global_tempering", false, &global_tempering_);
  }
  if (!get_double(kv, "hill_prefactor", true, &hill_prefactor_)) return 0;
  if (!get_double(kv, "bias_per_step", false, &bias_per_step_)) bias_per_step_ = hill_prefactor_;
  get_double(kv, "hill_density", false, &hill_density_);
  int d = 0;
  if (!get_int(kv, "dimension", true, &d)) return 0;
  dim_ = d;
  if (dim_ == 0 || dim_ > 3) {
    std::cerr << "Invalid dimesion " << dim_ << std::endl;
    return 0;
  }
  bias_dx_ = (double*)malloc(sizeof(double) * dim_);
  bias_sigma_ = (double*)malloc(sizeof(double) * dim_);
  min_ = (double*)malloc(sizeof(double) * dim_);
  max_ = (double*)malloc(sizeof(double) * dim_);
  b_periodic_boundary_ = (int*)malloc(sizeof(int) * dim_);
  if (!get_doubles(kv, "bias_spacing", true, bias_dx_, dim_)) return 0;
  if (!get_doubles(kv, "bias_sigma", true, bias_sigma_, dim_)) return 0;
  if (!get_doubles(kv, "box_low", true, min_, dim_)) return 0;
  if (!get_doubles(kv, "box_high", true, max_, dim_)) return 0;

  if (kv.find("target_filename") == kv.end()) {
    b_targeting_ = 0;
    expected_target_ = 0;
  } else {
    b_targeting_ = 1;
    target_ = read_grid(dim_, clean_string(kv["target_filename"], 0), 0);  // no interpolation
    expected_target_ = target_->expected_bias();
    std::cout << "Expected Target is " << expected_target_ << std::endl;
  }
  if (kv.find("initial_bias_filename") != kv.end())
    initial_bias_ = read_grid(dim_, clean_string(kv["initial_bias_filename"], 0), 1);

  std::string hills = kv.find("hills_filename") != kv.end() ? kv["hills_filename"] : std::string("HILLS");
  hill_output_.open(clean_string(hills, 1).c_str());
  std::string hist = kv.find("histogram_filename") != kv.end() ? kv["histogram_filename"] : std::string("HIST");
  hist_output_ = clean_string(hist, 0);
  return 1;
}

void EDMBias::setup(double temperature, double boltzmann_constant) {
  temperature_ = temperature;
  boltzmann_factor_ = boltzmann_constant * temperature;
}

// lib/edm_bias.cpp:98-222 for a replicated grid (serial build): decides grid vs boundary
// periodicity, creates the device grids, re-sets the boundary to the global box (T9).
void EDMBias::subdivide(const double sublo[3], const double subhi[3], const double boxlo[3], const double boxhi[3],
                        const int b_periodic[3], const double skin[3]) {
  if (bias_ != NULL) return;
  if (temperature_ < 0) edm_error("Must call setup before subdivide", "edm_bias.cpp:subdivide");
  int grid_period[3] = {0, 0, 0};
  double lo[3], hi[3];
  int bounds_flag = 1;
  for (unsigned i = 0; i < dim_; i++) {
    b_periodic_boundary_[i] = 0;
    if (fabs(boxlo[i] - min_[i]) < 0.000001 && fabs(boxhi[i] - max_[i]) < 0.000001)
      b_periodic_boundary_[i] = b_periodic[i];
  }
  for (unsigned i = 0; i < dim_; i++) {
    lo[i] = sublo[i];
    hi[i] = subhi[i];
    if (fabs(sublo[i] - min_[i]) < 0.000001 && fabs(subhi[i] - max_[i]) < 0.000001) {
      grid_period[i] = b_periodic[i];
      bounds_flag = 0;
    } else {
      lo[i] -= skin[i];
      hi[i] += skin[i];
    }
    bounds_flag &= (lo[i] >= max_[i] || hi[i] <= min_[i]);
  }
  bias_ = make_gauss_grid(dim_, lo, hi, bias_dx_, grid_period, INTERPOLATE, bias_sigma_);
  cv_hist_ = make_grid(dim_, lo, hi, bias_sigma_, grid_period, 0, 0);
  bias_->set_boundary(min_, max_, b_periodic_boundary_);
  if (initial_bias_ != NULL) bias_->add(initial_bias_, 1.0, 0.0);
  if (bounds_flag) {
    std::cout << "I am out of bounds!" << std::endl;
    b_outofbounds_ = 1;
    return;
  }
  total_volume_ = 0;
  total_volume_ += bias_->get_volume();
  create_device_state();
}

void EDMBias::create_device_state() {
  edm_bias_params_t p;
  p.dim = (int)dim_;
  p.b_tempering = b_tempering_;
  p.b_targeting = b_targeting_;
  p.global_tempering = global_tempering_;
  p.bias_factor = bias_factor_;
  p.boltzmann_factor = boltzmann_factor_;
  p.hill_prefactor = hill_prefactor_;
  p.bias_per_step = bias_per_step_;
  p.hill_density = hill_density_;
  p.expected_target = expected_target_;
  p.total_volume = total_volume_;
  edm_check(edm_bias_create(&dev_, bias_->device_grid(), cv_hist_->device_grid(),
                            target_ ? target_->device_grid() : NULL, &p),
            "edm_bias.cpp:subdivide");
  if (comm_) edm_check(edm_bias_set_comm(dev_, comm_, comm_cap_), "edm_bias.cpp:set_comm");
}

void EDMBias::set_comm(edm_comm_t* comm, long block_capacity) {
  comm_ = comm;
  comm_cap_ = block_capacity;
  mpi_rank_ = 0;
  mpi_size_ = 0;  // the serial build's values (lib/edm_bias.cpp:36-37)
  if (comm) {
    int n = 1, r = 0;
    edm_check(edm_comm_info(comm, &n, &r, NULL), "edm_bias.cpp:set_comm");
    mpi_rank_ = r;
    mpi_size_ = n;
  }
  if (dev_) edm_check(edm_bias_set_comm(dev_, comm_, comm_cap_), "edm_bias.cpp:set_comm");
}

void EDMBias::set_mask(const int* mask) { mask_ = mask; }

// LAMMPS hands rows of one contiguous n x 3 block; tests hand individually malloc'ed rows.  A
// contiguous block is passed through untouched, anything else is gathered into a scratch block.
const double* EDMBias::pack_rows(int n, const double* const* rows, int width, std::vector<double>& scratch,
                                 long* stride) const {
  if (n <= 0) {
    *stride = width;
    return NULL;
  }
  long s = n > 1 ? (long)(rows[1] - rows[0]) : width;
  bool contiguous = s >= width;
  for (int i = 1; contiguous && i < n; i++) contiguous = (rows[i] == rows[0] + (long)i * s);
  if (contiguous) {
    *stride = s;
    return rows[0];
  }
  scratch.resize((size_t)n * width);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < width; j++) scratch[(size_t)i * width + j] = rows[i][j];
  *stride = width;
  return scratch.data();
}

double EDMBias::update_forces(int nlocal, const double* const* positions, double** forces) const {
  return update_forces(nlocal, positions, forces, -1);
}

double EDMBias::update_forces(int nlocal, const double* const* positions, double** forces, int apply_mask) const {
  if (b_outofbounds_ || nlocal <= 0) return 0.0;
  long xs, fs;
  const double* x = pack_rows(nlocal, positions, (int)dim_, scratch_x_, &xs);
  double* f = const_cast<double*>(pack_rows(nlocal, forces, (int)dim_, scratch_f_, &fs));
  const bool gathered = (f == scratch_f_.data());
  double energy = 0;
  edm_check(edm_bias_update_forces(dev_, nlocal, x, xs, f, fs, mask_, apply_mask, &energy),
            "edm_bias.cpp:update_forces");
  if (gathered)
    for (int i = 0; i < nlocal; i++)
      for (unsigned j = 0; j < dim_; j++) forces[i][j] = f[(size_t)i * dim_ + j];
  return energy;
}

double EDMBias::update_forces_add_hills(int nlocal, const double* const* positions, double** forces,
                                        const double* runiform, int apply_mask, int do_hills) {
  if (b_outofbounds_ || nlocal <= 0) return 0.0;
  long xs, fs;
  const double* x = pack_rows(nlocal, positions, (int)dim_, scratch_x_, &xs);
  double* f = const_cast<double*>(pack_rows(nlocal, forces, (int)dim_, scratch_f_, &fs));
  const bool gathered = (f == scratch_f_.data());
  double energy = 0;
  edm_check(edm_bias_step_coords(dev_, nlocal, x, xs, f, fs, mask_, apply_mask, do_hills, runiform, 0,
                                 (unsigned long long)steps_, &energy),
            "edm_bias.cpp:update_forces_add_hills");
  if (gathered)
    for (int i = 0; i < nlocal; i++)
      for (unsigned j = 0; j < dim_; j++) forces[i][j] = f[(size_t)i * dim_ + j];
  if (do_hills) {
    drain_hill_log();
    refresh_state();
  }
  return energy;
}

double EDMBias::update_force(const double* positions, double* forces) const {
  if (b_outofbounds_) return 0.0;
  double energy = 0;
  edm_check(edm_bias_update_forces(dev_, 1, positions, (long)dim_, forces, (long)dim_, NULL, -1, &energy),
            "edm_bias.cpp:update_force");
  return energy;
}

void EDMBias::add_hills(int nlocal, const double* const* positions, const double* runiform) {
  add_hills(nlocal, positions, runiform, -1);
}

void EDMBias::add_hills(int nlocal, const double* const* positions, const double* runiform, int apply_mask) {
  if (b_outofbounds_) return;
  long xs;
  const double* x = pack_rows(nlocal, positions, (int)dim_, scratch_x_, &xs);
  edm_check(edm_bias_add_hills(dev_, nlocal, x, xs, runiform, mask_, apply_mask, 0, (unsigned long long)steps_),
            "edm_bias.cpp:add_hills");
  drain_hill_log();
  refresh_state();
}

void EDMBias::pre_add_hill(int est_hill_count) {
  if (b_outofbounds_) return;
  est_hill_count_ = est_hill_count;
  edm_check(edm_bias_pre_add_hill(dev_, est_hill_count), "edm_bias.cpp:pre_add_hill");
  in_round_ = 1;
  pending_x_.clear();
  pending_u_.clear();
}

// Candidates are queued on the host and handed to the device in batches: selection happens there.
void EDMBias::add_hill(const double* position, double runiform) {
  if (!in_round_) edm_error("Must call pre_add_hill before add_hill", "edm_bias.cpp:add_hill");
  for (unsigned i = 0; i < dim_; i++) pending_x_.push_back(position[i]);
  pending_u_.push_back(runiform);
  if (pending_u_.size() >= (1u << 20)) {
    edm_check(edm_bias_add_hill_batch(dev_, (long)pending_u_.size(), pending_x_.data(), pending_u_.data()),
              "edm_bias.cpp:add_hill");
    pending_x_.clear();
    pending_u_.clear();
  }
}

void EDMBias::post_add_hill() {
  if (!in_round_) return;
  if (!pending_u_.empty())
    edm_check(edm_bias_add_hill_batch(dev_, (long)pending_u_.size(), pending_x_.data(), pending_u_.data()),
              "edm_bias.cpp:post_add_hill");
  pending_x_.clear();
  pending_u_.clear();
  edm_check(edm_bias_post_add_hill(dev_), "edm_bias.cpp:post_add_hill");
  in_round_ = 0;
  drain_hill_log();
  refresh_state();
}

double EDMBias::pair_step(long natoms, const double* x, double* f, const int* type, int itype, int jtype,
                          const double box[3], double cutoff, int do_hills, long long est_hill_count,
                          unsigned long long seed, unsigned long long step, long long* ncalls) {
  edm_pair_result_t r;
  edm_check(edm_pair_step_cells(dev_, natoms, x, f, type, itype, jtype, box, cutoff, do_hills, est_hill_count, seed,
                                step, &r),
            "edm_bias.cpp:pair_step");
  if (ncalls) *ncalls = r.n_calls;
  if (do_hills) {
    drain_hill_log();
    refresh_state();
  }
  return r.energy;
}

void EDMBias::refresh_state() {
  bias_->device_changed();
  cv_hist_->device_changed();
  edm_bias_state_t s;
  edm_check(edm_bias_state(dev_, &s), "edm_bias.cpp:refresh_state");
  cum_bias_ = s.cum_bias;
  steps_ = s.steps;
}

void EDMBias::drain_hill_log() {
  std::vector<edm_hill_event_t> ev(4096);
  for (;;) {
    long n = 0;
    edm_check(edm_bias_log_read(dev_, NULL, 0, &n), "edm_bias.cpp:output_hill");
    if (n <= 0) break;
    if ((size_t)n > ev.size()) ev.resize(n);
    edm_check(edm_bias_log_read(dev_, ev.data(), (long)ev.size(), &n), "edm_bias.cpp:output_hill");
    hill_output_ << std::setprecision(8) << std::fixed;
    for (long k = 0; k < n; k++) {  // lib/edm_bias.cpp:590-599
      const edm_hill_event_t& e = ev[k];
      hill_output_ << e.steps << " " << (char)e.type << " " << e.hills_added << " ";
      for (unsigned i = 0; i < dim_; i++) hill_output_ << e.pos[i] << " ";
      hill_output_ << e.height << " " << e.bias_added << " " << e.cum_over_vol << std::endl;
    }
    break;
  }
}

void EDMBias::write_bias(const std::string& output) const { bias_->write(output); }
void EDMBias::write_histogram() const { cv_hist_->write(hist_output_); }
void EDMBias::clear_histogram() { cv_hist_->clear(); }
void EDMBias::write_lammps_table(const std::string& output) const { bias_->write(output); }  // serial build

}  // namespace EDM
