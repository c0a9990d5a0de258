/* edm_b200.h — C ABI of the B200-native EDM per-timestep bias engine.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.
 * The reference has no plugin ABI — consumers include <edm/edm_bias.h> and link libedm.so — so
 * each entry point below names the reference member function it stands in for (file:line into
 * whitead/electronic-dance-music).  The C++ classes EDM::Grid / EDM::GaussGrid / EDM::EDMBias in
 * electronic-dance-music_b200/edm/ are thin host code over exactly these calls, and
 * INTEGRATION.md shows how fix_edm.cpp / fix_edm_pair.cpp bind to them.
 *
 * Conventions
 *   - every function returns EDM_OK (0) or a negative edm_status_t; edm_last_error() gives text.
 *     (The reference aborts through edm_error(), lib/edm.cpp:4-7; the C++ mirror keeps that.)
 *   - "_dev" entry points take DEVICE pointers and a cudaStream_t passed as void* (NULL = the
 *     legacy default stream) and never synchronise; the others take HOST pointers, copy
 *     host->device->host inside the call and return when the result is in host memory.
 *   - all arithmetic is fp64; arrays are row-major; positions/forces are rows of `stride` doubles
 *     of which the first `dim` are used (LAMMPS passes stride 3, the reference tests stride 1).
 *   - there is no CPU fallback: without a CUDA device every call fails with EDM_ERR_NO_DEVICE.
 */
#ifndef EDM_B200_H
#define EDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum edm_status {
  EDM_OK = 0,
  EDM_ERR_ARG = -1,         /* bad argument */
  EDM_ERR_NO_DEVICE = -2,   /* no CUDA device / driver */
  EDM_ERR_CUDA = -3,        /* a CUDA call failed */
  EDM_ERR_STATE = -4,       /* call order violated (add_hill before pre_add_hill, ...) */
  EDM_ERR_CAPACITY = -5,    /* accepted-hill or hill-log buffer exhausted */
  EDM_ERR_BACKLOG_FULL = -6,/* overflow deque full: the reference aborts, lib/edm_bias.cpp:503-507 */
  EDM_ERR_COMM = -7         /* NCCL missing or a collective / rendezvous failed */
} edm_status_t;

typedef struct edm_grid edm_grid_t; /* Grid / GaussGrid resident in HBM */
typedef struct edm_bias edm_bias_t; /* EDMBias step state resident in HBM */
typedef struct edm_comm edm_comm_t; /* one rank's end of the hill exchange (an NCCL communicator) */

#define EDM_BUFFER_SLOTS 2048 /* BIAS_BUFFER_SIZE, lib/edm_bias.h:15 */
#define EDM_BUFFER_DBLS 8192  /* BIAS_BUFFER_DBLS, lib/edm_bias.h:16 */

const char* edm_last_error(void);
int edm_device_count(int* count);
/* Device-side counter-based uniform in [0,1) used when the caller passes no runiform array
 * (stand-in for LAMMPS RanMars, lammps/fix_edm.cpp:149-151); exposed so hosts can reproduce it. */
double edm_uniform(uint64_t seed, uint64_t step, uint64_t counter);
/* The two proposals of pair `pairkey` (which = 0, 1) share one hash: 32-bit resolution each. */
double edm_uniform_pair(uint64_t seed, uint64_t step, uint64_t pairkey, int which);

/* Page-locks / releases a host array the caller owns (cudaHostRegister), so that the host-buffer entry points
 * copy it at full PCIe rate and truly asynchronously; pageable arrays work everywhere, a few times slower.
 * For hosts without CUDA headers (the LAMMPS fixes pin atom->x and atom->f when EDM_B200_PIN is set).
 * Unpin before the array is freed or reallocated. */
int edm_host_pin(void* ptr, size_t bytes);
int edm_host_unpin(void* ptr);

/* ------------------------------------------------------------------ Grid / GaussGrid */

/* make_grid, lib/grid.h:911 -> DimmedGrid ctor lib/grid.h:190-213 */
int edm_grid_create(edm_grid_t** out, int device, int dim, const double* min, const double* max,
                    const double* spacing, const int* periodic, int b_derivatives, int b_interpolate);
/* geometry as DimmedGrid::read computes it from a PLUMED-1 header, lib/grid.h:759-806:
 * bins[] is the "#! BIN" line, max[] the "#! MAX" line (un-extended). */
int edm_grid_create_from_header(edm_grid_t** out, int device, int dim, const int* bins, const double* min,
                                const double* max, const int* periodic, int b_derivatives, int b_interpolate);
/* make_gauss_grid, lib/gaussian_grid.h:636 -> DimmedGaussGrid ctor lib/gaussian_grid.h:65-80 */
int edm_gauss_create(edm_grid_t** out, int device, int dim, const double* min, const double* max,
                     const double* spacing, const int* periodic, int b_interpolate, const double* sigma);
int edm_grid_destroy(edm_grid_t* g);
/* GaussGrid::set_boundary, lib/gaussian_grid.h:378-435 (rebuilds the McGDP tables) */
int edm_grid_set_boundary(edm_grid_t* g, const double* min, const double* max, const int* periodic);
/* public members of DimmedGrid / DimmedGaussGrid, lib/grid.h:876-885, lib/gaussian_grid.h:544-549.
 * Any output pointer may be NULL.  minisize is 0 for a plain grid. */
int edm_grid_geometry(const edm_grid_t* g, int* dim, int* n, double* dx, double* min, double* max,
                      int* periodic, int* minisize, size_t* size);
int edm_grid_flags(const edm_grid_t* g, int* b_derivatives, int* b_interpolate, int* is_gauss);
int edm_grid_boundary(const edm_grid_t* g, double* bmin, double* bmax, int* bperiodic, double* sigma_sqrt2);
/* Grid::set_interpolation, lib/grid.h:837 */
int edm_grid_set_interpolation(edm_grid_t* g, int b_interpolate);
/* values[size], derivs[size*dim] in the reference's layout (grid_[p], grid_deriv_[p*dim+d],
 * dim 0 fastest, +dV/dx); derivs may be NULL.  Stand-ins for direct grid_ access and for
 * Grid::read / Grid::write (lib/grid.h:448-503, 712-835), whose text I/O stays on the host. */
int edm_grid_upload(edm_grid_t* g, const double* values, const double* derivs);
int edm_grid_download(const edm_grid_t* g, double* values, double* derivs);
/* Grid::clear, lib/grid.h:679 */
int edm_grid_clear(edm_grid_t* g);
/* Grid::get_value_deriv batched, lib/grid.h:390-446 / lib/gaussian_grid.h:118-138 */
int edm_grid_eval(const edm_grid_t* g, long n, const double* x, long xstride, double* value, double* der);
int edm_grid_eval_dev(const edm_grid_t* g, long n, const double* x, long xstride, double* value, double* der,
                      void* stream);
/* DimmedGrid::get_value_deriv of the grid INSIDE a GaussGrid (lib/grid.h:390-446 alone, without the boundary test
 * and remap of lib/gaussian_grid.h:118-138): what grid_.multi_write evaluates, lib/grid.h:650-653.  On a plain
 * grid the same as edm_grid_eval. */
int edm_grid_eval_plain(const edm_grid_t* g, long n, const double* x, long xstride, double* value, double* der);
/* Grid::get_value batched, lib/grid.h:343-365 / lib/gaussian_grid.h:99-116 */
int edm_grid_get_value(const edm_grid_t* g, long n, const double* x, long xstride, double* value);
/* DimmedGrid::add_value (histogram bump), lib/grid.h:370-385; fails on interpolated grids */
int edm_grid_hist_add(edm_grid_t* g, long n, const double* x, long xstride, const double* v);
/* Grid::add, lib/grid.h:275-290: this += scale * other(x_p) + offset at every grid point */
int edm_grid_add(edm_grid_t* g, const edm_grid_t* other, double scale, double offset);
/* Grid::max_value / min_value, lib/grid.h:292-309 */
int edm_grid_minmax(const edm_grid_t* g, double* min_value, double* max_value);
/* GaussGrid::remap, lib/gaussian_grid.h:504-541, on one point (in place) */
int edm_grid_remap(const edm_grid_t* g, double* x);
/* GaussGrid::add_value batched, lib/gaussian_grid.h:176-372: deposits n hills in list order and
 * returns each hill's integrated bias.  Equivalent to n sequential add_value calls. */
int edm_gauss_deposit(edm_grid_t* g, long n, const double* centres, const double* heights, double* bias_added);
int edm_gauss_deposit_dev(edm_grid_t* g, long n, const double* centres, const double* heights,
                          double* bias_added, void* stream);

/* ------------------------------------------------------------------ EDMBias */

typedef struct edm_bias_params { /* the numbers read_input parses, lib/edm_bias.cpp:1009-1064 */
  int dim;
  int b_tempering;
  int b_targeting;
  double global_tempering;
  double bias_factor;
  double boltzmann_factor; /* kB*T, lib/edm_bias.cpp:267 */
  double hill_prefactor;
  double bias_per_step;
  double hill_density;    /* < 0: every candidate deposits, lib/edm_bias.cpp:543 */
  double expected_target; /* lib/edm_bias.cpp:1062 */
  double total_volume;    /* lib/edm_bias.cpp:211-220 */
} edm_bias_params_t;

typedef struct edm_bias_state { /* lib/edm_bias.h:130,161-178 */
  double cum_bias;
  double temp_hill_cum; /* last completed round's added bias */
  long long steps;
  int hills_added;
  int skipped;          /* b_skip_hill_add_ of the last round */
  long backlog_left, backlog_right;
  long n_accepted;      /* candidates that passed selection in the last round */
  long log_dropped;     /* hill-log records lost to a full log buffer */
} edm_bias_state_t;

typedef struct edm_hill_event { /* one HILLS line, lib/edm_bias.cpp:586-599 */
  long long steps;
  int type; /* 'h','u','b','v' (lib/edm_bias.h:20-25) */
  int hills_added;
  double pos[3];
  double height;
  double bias_added;
  double cum_over_vol;
} edm_hill_event_t;

/* bias and cv_hist are borrowed (the C++ EDMBias owns them); target may be NULL. */
int edm_bias_create(edm_bias_t** out, edm_grid_t* bias, edm_grid_t* cv_hist, edm_grid_t* target,
                    const edm_bias_params_t* params);
int edm_bias_destroy(edm_bias_t* b);
int edm_bias_state(edm_bias_t* b, edm_bias_state_t* out);
int edm_bias_set_cum_bias(edm_bias_t* b, double cum_bias);
/* overflow deque, lib/edm_bias.h:175-177: buffer holds EDM_BUFFER_DBLS doubles */
int edm_bias_backlog_get(edm_bias_t* b, long* left, long* right, double* buffer);
int edm_bias_backlog_set(edm_bias_t* b, long left, long right, const double* buffer);
/* drains the device hill log (HILLS records in deposit order) */
int edm_bias_log_read(edm_bias_t* b, edm_hill_event_t* out, long cap, long* n);

/* EDMBias::update_forces, lib/edm_bias.cpp:276-295: f[i][0..dim) -= dV/dx(x[i]), returns sum V.
 * mask may be NULL when apply_mask < 0. */
int edm_bias_update_forces(edm_bias_t* b, long n, const double* x, long xstride, double* f, long fstride,
                           const int* mask, int apply_mask, double* energy);
/* energy is a DEVICE pointer to one double */
int edm_bias_update_forces_dev(edm_bias_t* b, long n, const double* x, long xstride, double* f, long fstride,
                               const int* mask, int apply_mask, double* energy, void* stream);

/* The energy of the last edm_bias_update_forces_dev call made with energy = NULL (a one-CTA sum of the per-CTA
 * partials, in CTA order).  For callers that overlap a hill round with the force update: issuing it after the
 * round keeps it off the path force update -> deposit. */
int edm_bias_energy_dev(edm_bias_t* b, double* energy, void* stream);

/* One-shot: the next hill round launched for `b` also writes that energy to `energy` (device pointer): an idle CTA
 * of its deposit kernel adds the partials up, so the sum costs no launch of its own on the critical path. */
int edm_bias_energy_with_round(edm_bias_t* b, double* energy);

/* FixEDM::post_force, lammps/fix_edm.cpp:134-162, as ONE call on host buffers: update_forces over every
 * atom and, if do_hills (ntimestep % stride == 0, fix_edm.cpp:142), add_hills over the same atoms.  The
 * coordinates are uploaded once and the atoms stream through in chunks, so both PCIe directions and
 * the kernels overlap.  Same results as edm_bias_update_forces followed by edm_bias_add_hills.
 * runiform may be NULL (uniforms = edm_uniform(seed, step, i)); energy may be NULL. */
int edm_bias_step_coords(edm_bias_t* b, long n, const double* x, long xstride, double* f, long fstride,
                         const int* mask, int apply_mask, int do_hills, const double* runiform, uint64_t seed,
                         uint64_t step, double* energy);

/* The same step on DEVICE buffers (energy: device pointer to one double, may be NULL).  The round's selection,
 * plan, integrals and decision only read the grid and run on an internal stream beside the force update; the
 * deposit waits for the force update, `stream` waits for the round.  Never synchronises the host. */
int edm_bias_step_coords_dev(edm_bias_t* b, long n, const double* x, long xstride, double* f, long fstride,
                             const int* mask, int apply_mask, int do_hills, const double* runiform, uint64_t seed,
                             uint64_t step, double* energy, void* stream);

/* EDMBias::add_hills, lib/edm_bias.cpp:401-411 (= pre_add_hill(n); add_hill per masked atom;
 * post_add_hill()).  runiform may be NULL: uniforms then come from edm_uniform(seed, step, i). */
int edm_bias_add_hills(edm_bias_t* b, long n, const double* x, long xstride, const double* runiform,
                       const int* mask, int apply_mask, uint64_t seed, uint64_t step);
int edm_bias_add_hills_dev(edm_bias_t* b, long n, const double* x, long xstride, const double* runiform,
                           const int* mask, int apply_mask, uint64_t seed, uint64_t step, void* stream);

/* The streaming triple, lib/edm_bias.cpp:413-442, 528-563, 565-583.  add_hill_batch may be called
 * any number of times between pre and post; candidates keep their call order. */
int edm_bias_pre_add_hill(edm_bias_t* b, int est_hill_count);
int edm_bias_add_hill_batch(edm_bias_t* b, long n, const double* x, const double* runiform);
int edm_bias_post_add_hill(edm_bias_t* b);

/* ------------------------------------------------------------------ pair-distance CV (fix edm_pair) */

typedef struct edm_pair_result {
  double energy;        /* sum of V(r) over evaluated pairs, fix_edm_pair.cpp:217 */
  long long n_pairs;    /* pairs evaluated (update_force-equivalents) */
  long long n_calls;    /* hill proposals made: the caller's next est_hill_count, fix_edm_pair.cpp:245 */
} edm_pair_result_t;

/* FixEDMPair::post_force, lammps/fix_edm_pair.cpp:139-256, for a periodic orthorhombic box
 * [0,box)^3 with the neighbour search done on the device: every pair i<j with minimum-image
 * distance < cutoff whose types match (itype,jtype) is evaluated once (half list, both atoms
 * local), forces are accumulated into f, and if do_hills each pair proposes two hills
 * (fix_edm_pair.cpp:230-236) with uniforms edm_uniform_pair(seed, step, i*natoms+j, {0,1}),
 * accepted hills ordered by (i, j).  Lib-level order: all evaluations see the start-of-step bias
 * (SURVEY 3.2).  type may be NULL (all atoms match). */
int edm_pair_step_cells(edm_bias_t* b, long natoms, const double* x, double* f, const int* type, int itype,
                        int jtype, const double* box, double cutoff, int do_hills, long long est_hill_count,
                        uint64_t seed, uint64_t step, edm_pair_result_t* result);
/* result is a HOST pointer filled after an internal stream synchronise only if non-NULL */
int edm_pair_step_cells_dev(edm_bias_t* b, long natoms, const double* x, double* f, const int* type, int itype,
                            int jtype, const double* box, double cutoff, int do_hills,
                            long long est_hill_count, uint64_t seed, uint64_t step, edm_pair_result_t* result,
                            void* stream);
/* The same step for ONE RANK'S SHARE of a larger system, in LAMMPS' own picture (lammps/fix_edm_pair.cpp:177-236 with
 * newton off): the nall atoms handed over are the rank's nlocal local atoms followed by its ghosts (images and
 * neighbours' atoms within the cutoff of the sub-box, coordinates already shifted), all inside [lo, hi).  Dimensions
 * flagged periodic are wrapped by this rank itself (it holds the whole period); the others end at lo / hi.  As in
 * the reference: the loop runs over local atoms, so a pair of two ghosts is skipped; a pair with one ghost is
 * evaluated (its energy counted) but puts no force on the ghost (:223) and proposes ONE hill (:233) — the rank that
 * owns the ghost evaluates the pair too.  f holds nlocal rows (host form) / nall rows (device form, ghost rows
 * untouched).  With every dimension periodic and nlocal = nall this is edm_pair_step_cells. */
typedef struct edm_pair_domain {
  double lo[3], hi[3];
  int periodic[3];
  long nlocal;
} edm_pair_domain_t;
int edm_pair_step_cells_domain(edm_bias_t* b, long nall, const double* x, double* f, const int* type, int itype,
                               int jtype, const edm_pair_domain_t* dom, double cutoff, int do_hills,
                               long long est_hill_count, uint64_t seed, uint64_t step, edm_pair_result_t* result);
int edm_pair_step_cells_domain_dev(edm_bias_t* b, long nall, const double* x, double* f, const int* type, int itype,
                                   int jtype, const edm_pair_domain_t* dom, double cutoff, int do_hills,
                                   long long est_hill_count, uint64_t seed, uint64_t step, edm_pair_result_t* result,
                                   void* stream);
/* selection only (evaluation + proposals, no round): the first half of the multi-GPU step, see below */
int edm_pair_select_cells_domain_dev(edm_bias_t* b, long nall, const double* x, double* f, const int* type, int itype,
                                     int jtype, const edm_pair_domain_t* dom, double cutoff, long long est_hill_count,
                                     uint64_t seed, uint64_t step, double* energy_dev, void* stream);

/* Same loop over a caller-supplied half neighbour list in LAMMPS NeighList form flattened to CSR
 * (ilist[inum], first[inum+1] offsets into jlist[]; j already masked with NEIGHMASK; j >= nlocal
 * are ghosts whose force is not updated, fix_edm_pair.cpp:223).  runiform (2 per listed pair, in
 * list order) may be NULL: uniforms then come from edm_uniform(seed, step, 2*k+{0,1}). */
int edm_pair_step_list(edm_bias_t* b, long nall, long nlocal, const double* x, double* f, const int* type,
                       int itype, int jtype, long inum, const int* ilist, const long* first, const int* jlist,
                       int do_hills, long long est_hill_count, const double* runiform, uint64_t seed,
                       uint64_t step, edm_pair_result_t* result);

/* The same with the list kept on the device between calls: LAMMPS rebuilds its neighbour list only every few steps
 * (neighbor->ago == 0), so the fix uploads it with edm_pair_list_set on those steps and calls edm_pair_step_listed on
 * every step; only positions, forces and types cross PCIe then.  edm_pair_step_list = both in one call. */
int edm_pair_list_set(edm_bias_t* b, long inum, const int* ilist, const long* first, const int* jlist);
int edm_pair_step_listed(edm_bias_t* b, long nall, long nlocal, const double* x, double* f, const int* type, int itype,
                         int jtype, int do_hills, long long est_hill_count, const double* runiform, uint64_t seed,
                         uint64_t step, edm_pair_result_t* result);

/* How the last edm_pair_*_cells step searched for pairs: the brick of home cells a CTA owned
 * (0,0,0 = direct search), the density inflation used to size bricks, and how many steps so far
 * had to fall back to the direct search (which then shrinks the bricks).  Outputs may be NULL. */
int edm_pair_search_info(edm_bias_t* b, int* brick_dims, double* density_scale, long long* fallbacks);

/* ------------------------------------------------------------------ multi-GPU hill exchange */

/* Replaces flush_buffers / check_for_flush, lib/edm_bias.cpp:614-706 (broadcast mode): each rank
 * selects candidates locally (edm_*_select), packs its accepted hills into a fixed-capacity device
 * block, the caller all-gathers the blocks (NCCL), and every rank commits the rank-major
 * concatenation so all replicas deposit identical hills in one canonical order.
 * Block layout: double[0] = count, then count records of (dim) doubles, capacity cap records. */
size_t edm_hill_block_doubles(int dim, long cap);
int edm_bias_select_dev(edm_bias_t* b, long n, const double* x, long xstride, const double* runiform,
                        const int* mask, int apply_mask, long long est_hill_count, uint64_t seed, uint64_t step,
                        uint64_t first_counter, void* stream);
int edm_pair_select_cells_dev(edm_bias_t* b, long natoms, const double* x, double* f, const int* type,
                              int itype, int jtype, const double* box, double cutoff, long long est_hill_count,
                              uint64_t seed, uint64_t step, double* energy_dev, void* stream);
int edm_bias_hills_pack_dev(edm_bias_t* b, double* block, long cap, void* stream);
int edm_bias_hills_commit_dev(edm_bias_t* b, const double* blocks, int nblocks, long cap,
                              long long est_hill_count, void* stream);
/* One-shot, preferred over edm_bias_round_after: the next hill round launched for `b` puts its grid-WRITING
 * kernels (deposit, in-order tail) on `stream` — behind whatever is already enqueued there, i.e. this step's
 * force update — and only its read-only part (unpack, plan, integrals, decision) on the stream it was launched
 * on; the two are linked by an internal event.  The round is complete when `stream` is: no join needed, and
 * nothing wakes up across streams on the critical path force update -> deposit -> next force update. */
int edm_bias_round_commit_on(edm_bias_t* b, void* stream);
/* One-shot: the next hill round launched for `b` (add_hills_dev, hills_commit_dev, ...) waits for `event`
 * (a cudaEvent_t) right before its first write to the bias grid.  Lets a caller run selection, exchange
 * and the round's read-only kernels on a second stream while this step's force update, recorded by
 * `event`, still reads the start-of-step bias. */
int edm_bias_round_after(edm_bias_t* b, void* event);

/* The exchange itself, inside the library (lib/edm_bias.cpp:565-583 post_add_hill -> flush_buffers
 * :614-706 -> update_height :922-931, where the reference talks to MPI_COMM_WORLD directly).  NCCL is
 * resolved at run time; without it these calls fail with EDM_ERR_COMM.
 *
 * Bootstrap, one of:
 *   - one process per GPU: rank 0 calls edm_comm_unique_id and ships the 128 bytes to the others by any
 *     means (MPI_Bcast in a LAMMPS build, a torch.distributed broadcast in bench.py), every rank then
 *     calls edm_comm_init_rank; or edm_comm_init_file, which passes the id through a file all ranks see;
 *   - one process driving several devices: edm_comm_init_all (ncclCommInitAll);
 *   - an ncclComm_t the application already owns: edm_comm_from_nccl (borrowed, never destroyed here).
 * Creating a communicator of more than one rank is COLLECTIVE: every rank must make the call (the NVLink peer
 * windows are set up and agreed on over the communicator itself, see edm_comm_peer_windows). */
#define EDM_COMM_ID_BYTES 128
int edm_comm_nccl_version(int* version);
int edm_comm_unique_id(unsigned char* id /* [EDM_COMM_ID_BYTES] */);
int edm_comm_init_rank(edm_comm_t** out, const unsigned char* id, int nranks, int rank, int device);
int edm_comm_init_file(edm_comm_t** out, const char* path, int nranks, int rank, int device, double timeout_s);
int edm_comm_init_all(edm_comm_t** out /* [ndev] */, int ndev, const int* devices);
int edm_comm_from_nccl(edm_comm_t** out, void* nccl_comm, int nranks, int rank, int device);
int edm_comm_destroy(edm_comm_t* comm);
int edm_comm_info(const edm_comm_t* comm, int* nranks, int* rank, int* device);
/* How the exchange of this communicator travels.  When every rank of the job sits on one node and can map the
 * others' memory (CUDA IPC between processes, peer access inside one process), communicator creation also sets up
 * an NVLink peer window per rank, and an exchange is ONE kernel: it stores this rank's accepted hills into every
 * peer's window, publishes them with a system-scope release flag and waits for the peers' flags -- only the hills
 * that exist travel, and no collective is launched on the step's critical path (*enabled = 1).  Otherwise, or with
 * EDM_B200_NO_P2P=1 in the environment, it is pack -> one ncclAllGather of fixed-capacity blocks (*enabled = 0).
 * The choice is job-wide (agreed by an all-reduce at creation), the result bit-identical either way.  A peer that
 * never delivers raises EDM_ERR_COMM at the next edm_bias_check-style call after EDM_B200_PEER_TIMEOUT seconds
 * (default 10) instead of hanging the GPU. */
int edm_comm_peer_windows(const edm_comm_t* comm, int* enabled);
int edm_comm_group_start(void);
int edm_comm_group_end(void);
/* in-place sum over ranks of n doubles on the device (the bias energy, when a job-wide scalar is wanted;
 * replaces the MPI_Allreduce of lib/edm_bias.cpp:925 for callers that need one) */
int edm_comm_allreduce_sum_dev(edm_comm_t* comm, double* buf, long n, void* stream);
/* pack -> all-gather (peer windows or ncclAllGather, see edm_comm_peer_windows) -> commit in one call, on `stream`, never synchronising the host: what
 * post_add_hill does between the local add_hill calls and update_height.  est_total = the job-wide
 * est_hill_count (the same on every rank); cap = records per rank in the block (the same on every rank;
 * a rank that accepted more raises EDM_ERR_CAPACITY at the next edm_bias_state-style check instead of
 * dropping hills). */
int edm_bias_exchange_dev(edm_bias_t* b, edm_comm_t* comm, long cap, long long est_total, void* stream);
/* the same for n replicas driven by one thread (communicators from edm_comm_init_all) */
int edm_bias_exchange_all_dev(int n, edm_bias_t** b, edm_comm_t** comm, long cap, long long est_total,
                              void** streams);
/* Attaches a communicator: from now on every whole-round entry point of `b` (edm_bias_add_hills, the
 * pre/add/post triple, edm_bias_step_coords, edm_pair_step_cells, edm_pair_step_listed, and their _dev
 * forms) exchanges the accepted hills before the limiter, and its est_hill_count argument is this
 * rank's share: the job-wide count is est * nranks (the reference divides hill_density and
 * hill_prefactor by mpi_size_ instead, lib/edm_bias.cpp:175-180 — same acceptance, same heights).
 * The select/pack/commit building blocks above are not affected.  comm = NULL detaches. */
int edm_bias_set_comm(edm_bias_t* b, edm_comm_t* comm, long cap);
/* Error state of the last round(s): EDM_ERR_BACKLOG_FULL, EDM_ERR_CAPACITY (accepted-hill buffer or
 * exchange block exhausted) or EDM_OK.  Synchronises on a small device-to-host copy. */
int edm_bias_check(edm_bias_t* b);

/* ------------------------------------------------------------------ measurement hooks (bench.py) */

/* Kernels launched by this library since load (all devices, this process). */
int edm_launch_count(long long* count);
/* When on, the pair kernels are bracketed by CUDA events on their launch stream. */
int edm_bias_set_profiling(edm_bias_t* b, int on);
/* Duration of the last profiled pair kernel in ms (synchronises on its end event). */
int edm_bias_profile_ms(edm_bias_t* b, double* pair_kernel_ms);
/* The same interval split at the boundary between the block search and the block evaluation
 * kernel (both 0 when the generic search ran). */
int edm_bias_profile_pair_ms(edm_bias_t* b, double* search_ms, double* eval_ms);

/* Device-side split of the last profiled edm_pair_step_cells call (host buffers): positions host->device,
 * kernels (binning .. hill round), forces device->host (overlaps the hill round), and the span from the
 * first copy to the later of the two ends.  Outputs may be NULL. */
int edm_bias_profile_e2e_ms(edm_bias_t* b, double* x_up_ms, double* kernels_ms, double* f_down_ms, double* span_ms);

/* Device-clock (%globaltimer) stamps of the last hill round, microseconds since the plan kernel began:
 * [0..6] plan phases, [7,8] decision begin/end, [9,10] first deposit taken / last deposit done,
 * [11,12] in-order kernel begin/end, [13] the deposit grid's last CTA left, [14] the last force update began,
 * [15] the in-order kernel became resident (before it waits for its predecessor).
 * out holds 16 doubles.  Synchronises on a small copy. */
int edm_bias_round_times_us(edm_bias_t* b, double* out16);
/* The same clock for the selection and the exchange that fed that round: [0] selection began, [1] its last CTA
 * left, [2] the exchange kernel got past its predecessor, [3] this rank's block delivered to every peer, [4] every
 * peer's block in (peer-window transport; the ncclAllGather path leaves [2..4] untouched).  out holds 5 doubles. */
int edm_bias_exchange_times_us(edm_bias_t* b, double* out5);

/* How the hill rounds so far ran.  `parallel`: planned, integrated and deposited all hills at once.
 * `split`: the hills before the one at which the running sum reaches bias_per_step went in at once,
 * the rest of the round hill by hill.  `in_order`: pre_add_hill / add_hill / post_add_hill walked
 * hill by hill (backlog not empty, or a 1-D grid with local tempering).  Same results either way.
 * Outputs may be NULL. */
int edm_bias_round_info(edm_bias_t* b, long long* parallel, long long* split, long long* in_order);

#ifdef __cplusplus
}
#endif
#endif /* EDM_B200_H */
