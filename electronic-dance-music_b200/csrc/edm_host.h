// Host-side object layouts behind the opaque C handles of include/edm_b200.h.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/edm_b200.h"
#include "edm_device.cuh"

namespace edm {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define EDM_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) return ::edm::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

#define EDM_REQUIRE(cond, msg)        \
  do {                                \
    if (!(cond)) {                    \
      ::edm::set_error(msg);          \
      return EDM_ERR_ARG;             \
    }                                 \
  } while (0)

#define EDM_TRY(call)            \
  do {                           \
    int rc__ = (call);           \
    if (rc__ != EDM_OK) return rc__; \
  } while (0)

int ensure_device(int device);
// multiprocessors of a device (cudaDevAttrMultiProcessorCount, cached): grids are sized in multiples of it
int sm_count(int device);
// 1-D owner-computes deposit split in two so a caller can decide between them (edm_grid.cu):
// stage = prepare + accumulate into scratch (+ per-hill integrals), commit = write the grid back
// only if *flag >= want.  n is read from *n_dev (clamped to n_max) when n_dev is not NULL.
bool deposit1d_eligible(const edm_grid* g);
int deposit1d_stage(edm_grid* g, const double* centres, const double* heights, double* ba, const int* n_dev,
                    long n_max, cudaStream_t st);
int deposit1d_commit_if(edm_grid* g, const int* flag, int want, cudaStream_t st);
// duplicate_boundary (lib/gaussian_grid.h:571-630) after a batch of hills, only if *gate == want
int edm_grid_dup_boundary_if(edm_grid* g, const int* gate, int want, cudaStream_t st);
void count_launches(int n);

// Programmatic dependent launch (sm_90+): the kernel may be scheduled while its predecessor in the stream is still
// running and waits at pdl_wait() — the launch latency (several microseconds between the hill round's small
// dependent kernels) overlaps with the predecessor instead of following it.  Kernels launched this way call
// pdl_trigger() first thing (their own successor may be staged at once) and pdl_wait() before they touch anything
// a predecessor wrote; both are no-ops under an ordinary <<<>>> launch.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// Device scratch that grows on demand and is reused across calls (no allocation on the steady
// per-step path once sizes settle).
struct Scratch {
  void* p = nullptr;
  size_t bytes = 0;
  int reserve(size_t need);
  void release();
  template <typename T> T* as() { return static_cast<T*>(p); }
};

}  // namespace edm

struct edm_grid {
  int device = 0;
  edm::GridDesc d;          // what kernels receive by value
  double sigma_user[3];     // sigma as given (bias_sigma), before the sqrt(2)
  std::vector<double> bc_denom[3], bc_deriv[3];  // McGDP tables, lib/gaussian_grid.h:551-552
  double* d_ptab[3] = {nullptr, nullptr, nullptr};
  long long* d_dup = nullptr;
  edm::Scratch io;          // staging for upload/download/eval of host buffers
  edm::Scratch work;        // deposit scratch (prepared hills, partials, slots)
  int* d_flags = nullptr;   // [0] dirty-bounds flag
  double* stage_partial = nullptr;  // accumulators left by the last deposit1d_stage call
};

struct HillAccepted {  // one selected candidate
  unsigned long long key;
  double x[3];
};

#define EDM_ROUND_MAX 2048  // hills one parallel round plans; larger rounds run in order

struct BiasDev {  // lives in HBM; mirrors the mutable members of EDMBias, lib/edm_bias.h:130,161-178
  double cum_bias;
  double temp_hill_cum;
  long long steps;
  int hills_added;
  int skip;
  long long left, right;
  int n_accepted;
  int accepted_overflow;
  int accepted_sorted;  // the accepted list is already in key order (packed or unpacked): rounds skip their sort
  int log_n;
  int log_dropped;
  int backlog_full;
  int round_mode;  // 0: in-order round, 1: parallel round planned, 2: committed in full, 3: committed hills [0, n_fast)
  int n_fast;      // entries planned, then committed, by the parallel round
  int n_plan_b;    // how many of the planned entries are backlog slots (they come first)
  int n_accepted_last;  // candidates the last finished round consumed
  int rounds_parallel, rounds_split, rounds_in_order;  // how the rounds so far ran (edm_bias_round_info)
  int ticket;      // next hill a CTA of the parallel deposit takes
  int int_done;    // CTAs of round_integrals_kernel that signed off (the last one takes the decision)
  // %globaltimer stamps of the last round (ns): plan phases 0-6, decide begin/end, deposit first/last, in-order begin/end
  unsigned long long stamp[16];
  int round_epoch; // value a finished hill leaves in hill_done[]: one more per parallel round
  int exchange_timeout;  // a peer's hill block did not arrive in time (edm_bias_check reports EDM_ERR_COMM)
  // more %globaltimer stamps (ns): [0] selection began, [1] selection's last CTA left, [2] exchange kernel past its
  // predecessor, [3] own block delivered to every peer, [4] every peer's block in (edm_bias_exchange_times_us)
  unsigned long long stamp2[8];
  unsigned long long n_pairs;
  unsigned long long n_pairs_ghost;  // of those, pairs with one ghost atom (one hill proposal instead of two)
  double overflow[EDM_BUFFER_DBLS + 8];  // T19: slack for the D=3 write one record past the array
  int hill_done[EDM_ROUND_MAX];          // parallel deposit: hill k finished in round hill_done[k]
};

// Host-mapped (zero-copy) block a host-buffer step writes its scalars to with its last kernel, so the host
// reads them right after the stream synchronise instead of issuing three blocking 8-byte copies.
struct HostReport {
  int fallback;  // a pair step fell back to the direct search (read without synchronising by the next launch)
  int backlog_full, accepted_overflow, pad;
  double energy;
  unsigned long long n_pairs, n_calls;
};

// Peer window of the hill exchange (edm_comm.cu): a small buffer on every rank that every other rank of the node can
// store into over NVLink (CUDA IPC between processes, peer access inside one).  Layout: 2 x 64 int flags
// (flag[parity][source rank] = epoch of the block that source last delivered), then 2 x nranks block slots.
#define EDM_PEER_MAX_RANKS 64
#define EDM_PEER_FLAG_BYTES (2 * EDM_PEER_MAX_RANKS * sizeof(int))
#define EDM_PEER_SLOT_DOUBLES (1 + 4096 * 3)  // one block of the default capacity in 3-D

struct edm_comm {  // one rank's end of the hill exchange (edm_comm.cu)
  void* nccl = nullptr;  // ncclComm_t
  int nranks = 1, rank = 0, device = 0;
  int owned = 0;         // the ncclComm_t is destroyed with the handle
  // NVLink peer window (p2p != 0): the exchange is one kernel that stores this rank's hills into every peer's window
  // and waits for theirs -- no collective launch on the step's critical path
  int p2p = 0;
  char* win = nullptr;            // this rank's window
  char** d_peer = nullptr;        // device array [nranks]: every rank's window as seen from this device
  void* opened[EDM_PEER_MAX_RANKS] = {};  // IPC mappings to close
  int epoch = 0;                  // exchanges issued so far (the same on every rank)
};

struct edm_bias {
  int device = 0;
  edm_bias_params_t prm;
  edm_grid* bias = nullptr;
  edm_grid* hist = nullptr;
  edm_grid* target = nullptr;
  BiasDev* d_state = nullptr;
  HillAccepted* d_accepted = nullptr;
  long accepted_cap = 0;
  edm_hill_event_t* d_log = nullptr;
  long log_cap = 0;
  double* d_energy_partial = nullptr;  // per-CTA energy partials
  int n_partial = 0;
  int last_partials = 0;               // CTAs (= partials) of the last force update
  double* energy_with_round = nullptr; // one-shot: the next round also sums those partials into this device double
  double* d_scalar = nullptr;          // [0] energy
  edm::Scratch io, io2, io3, io4;      // host<->device staging for the host-pointer entry points
  edm::Scratch cells;                  // cell-list scratch of the pair kernels
  edm::Scratch cand;                   // candidate items of the block pair search, one slice per brick
  volatile int* h_pair_flags = nullptr; // host-mapped HostReport: [0] = a pair step fell back to the direct search
  int* d_pair_flags = nullptr;
  cudaEvent_t ev_prev = nullptr;        // "everything enqueued on the default stream before this call"
  cudaEvent_t ev_e2e[5] = {};           // profiling: start, x is up, kernels done, f-down start, f-down end
  int e2e_valid = 0;
  int brick_dims[3] = {0, 0, 0};       // bricks of the last pair step (0: direct search)
  long long pair_fallbacks = 0;        // steps that fell back to the direct search
  double brick_scale = 1.0;            // density inflation used to size the bricks (grows after a fallback)
  edm::Scratch list;                   // neighbour list kept between calls: ilist | first | jlist | row_of
  long list_inum = 0, list_nlisted = 0;
  int list_valid = 0;
  int* list_ilist = nullptr;
  long* list_first = nullptr;
  int* list_jlist = nullptr;
  int* list_row = nullptr;
  edm::Scratch fast;                   // centres | heights | bias_added of the parallel hill round
  // multi-GPU: communicator attached by edm_bias_set_comm (borrowed) and the exchange buffers
  edm_comm_t* comm = nullptr;
  long comm_cap = 4096;
  edm::Scratch xchg;                   // this rank's hill block | the rank-major concatenation of all blocks
  // streaming triple
  int in_round = 0;
  long long round_est = 0;
  unsigned long long round_count = 0;
  int profiling = 0;
  int eval_attr_set = 0;               // block_eval_kernel's dynamic shared-memory limit raised on this device
  cudaStream_t st_main = nullptr, st_copy = nullptr;  // host-buffer pair step: kernels / overlapped copies
  cudaEvent_t ev_f_up = nullptr, ev_f_final = nullptr;
  // host-buffer coordinate step: chunks pipelined over upload / kernels / download
  static constexpr int kMaxChunks = 16;
  cudaStream_t st_up = nullptr;
  cudaEvent_t ev_chunk_up[kMaxChunks] = {}, ev_chunk_done[kMaxChunks] = {};
  double* d_chunk_energy = nullptr;
  // device-buffer coordinate step: the read-only part of the hill round runs beside the force update
  cudaStream_t st_side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_forces = nullptr, ev_join = nullptr;
  cudaEvent_t round_after = nullptr;  // borrowed, one-shot: the next round's first grid write waits for it
  // one-shot (edm_bias_round_commit_on): the next round's grid-writing kernels go to this stream
  cudaStream_t commit_stream = nullptr;
  int commit_stream_set = 0;
  cudaEvent_t ev_round_ready = nullptr;  // the round's read-only part is done
  cudaEvent_t forces_event = nullptr; // one-shot: recorded right after the next forces_kernel launch
  // one-shot: the next round takes its candidates from these exchange blocks (the plan unpacks them itself)
  const double* round_blocks = nullptr;
  int round_nblocks = 0;
  long round_block_cap = 0;
  cudaEvent_t ev_pair[3] = {nullptr, nullptr, nullptr};  // pair kernels: begin, end, between search and evaluation
};

// ---- internal entry points shared by the translation units.  est is always the JOB-WIDE est_hill_count here;
// the public whole-round entry points convert their per-rank argument with edm_job_est.
int edm_bias_reset_accepted(edm_bias* b, cudaStream_t st);
int edm_bias_size_accepted(edm_bias* b, double candidates, long long est);
// the hill round over whatever sits in the accepted buffer; with a communicator attached and
// exchange = true the accepted hills of all ranks are gathered first (edm_bias_exchange_round)
int edm_bias_launch_round(edm_bias* b, long long est, cudaStream_t st, bool exchange = true);
int edm_bias_exchange_round(edm_bias* b, long long est, cudaStream_t st);
int edm_bias_hills_push_dev(edm_bias* b, edm_comm* c, long cap, cudaStream_t st, const double** blocks_out);
int edm_bias_check_round(edm_bias* b);
int edm_host_report_ensure(edm_bias* b);
inline long long edm_job_est(const edm_bias* b, long long est) {
  return (b->comm && b->comm->nranks > 1) ? est * b->comm->nranks : est;
}
