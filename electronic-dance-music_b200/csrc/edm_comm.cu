// Multi-GPU hill exchange inside the library: NCCL communicator bootstrap and the all-gather of the
// per-rank hill blocks.  Replaces flush_buffers / check_for_flush / update_height's reductions
// (lib/edm_bias.cpp:614-706, 922-931): one ncclAllGather of fixed-capacity blocks instead of 2*size
// MPI_Bcast + 2 MPI_Allreduce per hill step; every rank then commits the rank-major concatenation, so
// the replicas stay bit-identical and cum_bias_ needs no reduction.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2, or the copy a host such as PyTorch already
// mapped), so single-GPU users of libedm_b200.so do not need it; a multi-GPU call without NCCL fails
// loudly with EDM_ERR_COMM.
#include <dlfcn.h>
#include <fcntl.h>
#include <nccl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "edm_host.h"

namespace edm {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;
static std::string g_nccl_error;

static void nccl_load() {
  const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  void* h = nullptr;
  for (int i = 0; names[i] && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already mapped?
  for (int i = 0; names[i] && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    g_nccl_error = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "");
    return;
  }
  NcclApi a;
  a.handle = h;
#define EDM_NCCL_SYM(field, sym)                                              \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, sym));               \
  if (!a.field) {                                                             \
    g_nccl_error = std::string("NCCL symbol missing: ") + sym;                \
    return;                                                                   \
  }
  EDM_NCCL_SYM(GetVersion, "ncclGetVersion")
  EDM_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  EDM_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  EDM_NCCL_SYM(CommInitAll, "ncclCommInitAll")
  EDM_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  EDM_NCCL_SYM(AllGather, "ncclAllGather")
  EDM_NCCL_SYM(AllReduce, "ncclAllReduce")
  EDM_NCCL_SYM(GroupStart, "ncclGroupStart")
  EDM_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  EDM_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef EDM_NCCL_SYM
  g_nccl = a;
}

static const NcclApi* nccl_api() {
  std::call_once(g_nccl_once, nccl_load);
  if (!g_nccl.handle) {
    set_error(g_nccl_error);
    return nullptr;
  }
  return &g_nccl;
}

static int nccl_fail(const NcclApi* a, ncclResult_t r, const char* what) {
  char buf[384];
  snprintf(buf, sizeof(buf), "NCCL error %d (%s) in %s", (int)r, a ? a->GetErrorString(r) : "?", what);
  set_error(buf);
  return EDM_ERR_COMM;
}

#define EDM_NCCL(api, call)                                          \
  do {                                                               \
    ncclResult_t r__ = (api)->call;                                  \
    if (r__ != ncclSuccess) return nccl_fail((api), r__, #call);     \
  } while (0)

// ---- NVLink peer windows -------------------------------------------------------------------------------------
// What every rank tells the others about its window.  Same pid: the ranks share a process (threads driving one
// device each) and use the raw pointer after cudaDeviceEnablePeerAccess; other pid, same host: CUDA IPC.
struct PeerCard {
  cudaIpcMemHandle_t handle;
  unsigned long long ptr;
  long long pid;
  unsigned long long host;  // hash of the host name: ranks on another node cannot be reached this way
  int device, ok;
};

static size_t peer_window_bytes(int nranks) {
  return EDM_PEER_FLAG_BYTES + 2 * (size_t)nranks * EDM_PEER_SLOT_DOUBLES * sizeof(double);
}

static unsigned long long host_hash() {
  char name[256] = {0};
  gethostname(name, sizeof(name) - 1);
  unsigned long long h = 1469598103934665603ULL;
  for (const char* p = name; *p; p++) h = (h ^ (unsigned char)*p) * 1099511628211ULL;
  return h;
}

static void peer_release(edm_comm* c) {
  cudaSetDevice(c->device);
  for (int r = 0; r < EDM_PEER_MAX_RANKS; r++)
    if (c->opened[r]) {
      cudaIpcCloseMemHandle(c->opened[r]);
      c->opened[r] = nullptr;
    }
  if (c->d_peer) cudaFree(c->d_peer);
  if (c->win) cudaFree(c->win);
  c->d_peer = nullptr;
  c->win = nullptr;
  c->p2p = 0;
}

static bool peer_disabled_by_env() {
  const char* e = getenv("EDM_B200_NO_P2P");
  return e && atoi(e) != 0;
}

// Maps the windows described by cards[0..nranks) into this rank's device; false if any cannot be reached.
static bool peer_map(edm_comm* c, const PeerCard* cards, char** peers) {
  const long long me = (long long)getpid();
  const unsigned long long host = cards[c->rank].host;
  for (int r = 0; r < c->nranks; r++) {
    if (!cards[r].ok || cards[r].host != host) return false;
    if (r == c->rank) {
      peers[r] = c->win;
    } else if (cards[r].pid == me) {
      if (cards[r].device != c->device) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, c->device, cards[r].device) != cudaSuccess || !can) return false;
        cudaError_t e = cudaDeviceEnablePeerAccess(cards[r].device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
          cudaGetLastError();
          return false;
        }
        cudaGetLastError();
      }
      peers[r] = reinterpret_cast<char*>(cards[r].ptr);
    } else {
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, cards[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return false;
      }
      c->opened[r] = p;
      peers[r] = static_cast<char*>(p);
    }
  }
  return true;
}

static bool peer_alloc(edm_comm* c, PeerCard* card) {
  memset(card, 0, sizeof(*card));
  card->pid = (long long)getpid();
  card->host = host_hash();
  card->device = c->device;
  if (c->nranks > EDM_PEER_MAX_RANKS || peer_disabled_by_env()) return false;
  if (cudaSetDevice(c->device) != cudaSuccess) return false;
  const size_t bytes = peer_window_bytes(c->nranks);
  if (cudaMalloc(&c->win, bytes) != cudaSuccess || cudaMemset(c->win, 0, bytes) != cudaSuccess ||
      cudaDeviceSynchronize() != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  card->ptr = reinterpret_cast<unsigned long long>(c->win);
  if (cudaIpcGetMemHandle(&card->handle, c->win) != cudaSuccess) {
    cudaGetLastError();  // no IPC on this platform: still usable between devices of one process
    memset(&card->handle, 0, sizeof(card->handle));
  }
  card->ok = 1;
  return true;
}

static bool peer_finish(edm_comm* c, char** peers) {
  if (cudaMalloc(&c->d_peer, c->nranks * sizeof(char*)) != cudaSuccess) return false;
  if (cudaMemcpy(c->d_peer, peers, c->nranks * sizeof(char*), cudaMemcpyHostToDevice) != cudaSuccess) return false;
  c->p2p = 1;
  return true;
}

// One rank per process (edm_comm_init_rank, edm_comm_from_nccl): the cards travel over the communicator itself, and
// so does the verdict -- the window is used only if EVERY rank mapped every peer.  Collective.
static void peer_setup_collective(edm_comm* c, const NcclApi* a) {
  if (c->nranks < 2 || c->nranks > EDM_PEER_MAX_RANKS) return;
  PeerCard mine;
  peer_alloc(c, &mine);
  std::vector<PeerCard> cards(c->nranks);
  PeerCard* d_cards = nullptr;
  int* d_ok = nullptr;
  int ok = 0;
  bool comm_fine = cudaMalloc(&d_cards, (c->nranks + 1) * sizeof(PeerCard)) == cudaSuccess &&
                   cudaMalloc(&d_ok, sizeof(int)) == cudaSuccess;
  if (comm_fine) {
    cudaMemcpy(d_cards + c->nranks, &mine, sizeof(PeerCard), cudaMemcpyHostToDevice);
    comm_fine = a->AllGather(d_cards + c->nranks, d_cards, sizeof(PeerCard), ncclChar, (ncclComm_t)c->nccl, 0) == ncclSuccess &&
                cudaStreamSynchronize(0) == cudaSuccess;
  }
  if (comm_fine) {
    cudaMemcpy(cards.data(), d_cards, c->nranks * sizeof(PeerCard), cudaMemcpyDeviceToHost);
    char* peers[EDM_PEER_MAX_RANKS];
    ok = mine.ok && peer_map(c, cards.data(), peers) && peer_finish(c, peers) ? 1 : 0;
    cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice);
    comm_fine = a->AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, (ncclComm_t)c->nccl, 0) == ncclSuccess &&
                cudaStreamSynchronize(0) == cudaSuccess;
    if (comm_fine) cudaMemcpy(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost);
  }
  if (d_cards) cudaFree(d_cards);
  if (d_ok) cudaFree(d_ok);
  cudaGetLastError();
  if (!comm_fine || !ok) peer_release(c);  // the all-gather path remains
}

}  // namespace edm

using namespace edm;

// pack -> all-gather; leaves the rank-major concatenation in b->xchg (gathered part) and returns it
static int exchange_gather(edm_bias* b, edm_comm* c, long cap, cudaStream_t st, const double** blocks_out) {
  const size_t bw = edm_hill_block_doubles(b->prm.dim, cap);
  if (c->p2p && c->nranks > 1 && bw <= (size_t)EDM_PEER_SLOT_DOUBLES)  // the blocks fit the peer window's slots
    return edm_bias_hills_push_dev(b, c, cap, st, blocks_out);
  EDM_TRY(b->xchg.reserve((size_t)(c->nranks + 1) * bw * sizeof(double)));
  double* block = b->xchg.as<double>();
  double* gathered = block + bw;
  EDM_TRY(edm_bias_hills_pack_dev(b, block, cap, st));
  if (c->nranks == 1) {
    *blocks_out = block;
    return EDM_OK;
  }
  const NcclApi* a = nccl_api();
  if (!a) return EDM_ERR_COMM;
  EDM_NCCL(a, AllGather(block, gathered, bw, ncclDouble, (ncclComm_t)c->nccl, st));
  *blocks_out = gathered;
  return EDM_OK;
}

int edm_bias_exchange_round(edm_bias* b, long long est_total, cudaStream_t st) {
  edm_comm* c = b->comm;
  const double* blocks = nullptr;
  EDM_TRY(exchange_gather(b, c, b->comm_cap, st, &blocks));
  return edm_bias_hills_commit_dev(b, blocks, c->nranks, b->comm_cap, est_total, st);
}

extern "C" {

int edm_comm_nccl_version(int* version) {
  EDM_REQUIRE(version != nullptr, "NULL argument");
  const NcclApi* a = nccl_api();
  if (!a) return EDM_ERR_COMM;
  EDM_NCCL(a, GetVersion(version));
  return EDM_OK;
}

int edm_comm_unique_id(unsigned char* id) {
  EDM_REQUIRE(id != nullptr, "NULL argument");
  static_assert(EDM_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "unique id size");
  const NcclApi* a = nccl_api();
  if (!a) return EDM_ERR_COMM;
  ncclUniqueId u;
  EDM_NCCL(a, GetUniqueId(&u));
  memcpy(id, u.internal, EDM_COMM_ID_BYTES);
  return EDM_OK;
}

int edm_comm_init_rank(edm_comm_t** out, const unsigned char* id, int nranks, int rank, int device) {
  EDM_REQUIRE(out && id && nranks > 0 && rank >= 0 && rank < nranks, "bad argument");
  EDM_TRY(ensure_device(device));
  edm_comm* c = new edm_comm();
  c->nranks = nranks;
  c->rank = rank;
  c->device = device;
  if (nranks > 1) {
    const NcclApi* a = nccl_api();
    if (!a) {
      delete c;
      return EDM_ERR_COMM;
    }
    ncclUniqueId u;
    memcpy(u.internal, id, EDM_COMM_ID_BYTES);
    ncclResult_t r = a->CommInitRank(reinterpret_cast<ncclComm_t*>(&c->nccl), nranks, u, rank);
    if (r != ncclSuccess) {
      delete c;
      return nccl_fail(a, r, "ncclCommInitRank");
    }
    c->owned = 1;
    peer_setup_collective(c, a);
  }
  *out = c;
  return EDM_OK;
}

// Rendezvous through a file every rank can see (no MPI in the picture): rank 0 creates `path` exclusively
// with the unique id, the others poll for it; ncclCommInitRank is itself a barrier, after which rank 0
// removes the file.  `path` must not exist before the job starts.
int edm_comm_init_file(edm_comm_t** out, const char* path, int nranks, int rank, int device, double timeout_s) {
  EDM_REQUIRE(out && path && nranks > 0 && rank >= 0 && rank < nranks, "bad argument");
  unsigned char id[EDM_COMM_ID_BYTES];
  if (nranks == 1) {
    memset(id, 0, sizeof(id));
    return edm_comm_init_rank(out, id, 1, 0, device);
  }
  if (rank == 0) {
    EDM_TRY(edm_comm_unique_id(id));
    const std::string tmp = std::string(path) + ".tmp";
    int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0600);
    if (fd < 0 || write(fd, id, sizeof(id)) != (ssize_t)sizeof(id)) {
      if (fd >= 0) close(fd);
      set_error(std::string("cannot write the rendezvous file ") + tmp);
      return EDM_ERR_COMM;
    }
    close(fd);
    if (link(tmp.c_str(), path) != 0) {  // fails if `path` exists: a stale file of an earlier job
      unlink(tmp.c_str());
      set_error(std::string("rendezvous file already exists: ") + path);
      return EDM_ERR_COMM;
    }
    unlink(tmp.c_str());
  } else {
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
      int fd = open(path, O_RDONLY);
      if (fd >= 0) {
        const ssize_t got = read(fd, id, sizeof(id));
        close(fd);
        if (got == (ssize_t)sizeof(id)) break;
      }
      const double waited = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (waited > timeout_s) {
        set_error(std::string("timed out waiting for the rendezvous file ") + path);
        return EDM_ERR_COMM;
      }
      std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
  }
  const int rc = edm_comm_init_rank(out, id, nranks, rank, device);
  if (rank == 0) unlink(path);
  return rc;
}

// One process driving several devices (ncclCommInitAll): out[i] is the communicator of devices[i]
// (devices == NULL: 0..ndev-1).  Collectives issued from one thread for several of these must sit
// between edm_comm_group_start/end, which is what edm_bias_exchange_all_dev does.
int edm_comm_init_all(edm_comm_t** out, int ndev, const int* devices) {
  EDM_REQUIRE(out && ndev > 0 && ndev <= 64, "bad argument");
  int devs[64];
  for (int i = 0; i < ndev; i++) devs[i] = devices ? devices[i] : i;
  for (int i = 0; i < ndev; i++) EDM_TRY(ensure_device(devs[i]));
  ncclComm_t comms[64];
  if (ndev > 1) {
    const NcclApi* a = nccl_api();
    if (!a) return EDM_ERR_COMM;
    EDM_NCCL(a, CommInitAll(comms, ndev, devs));
  }
  for (int i = 0; i < ndev; i++) {
    edm_comm* c = new edm_comm();
    c->nranks = ndev;
    c->rank = i;
    c->device = devs[i];
    c->nccl = ndev > 1 ? comms[i] : nullptr;
    c->owned = ndev > 1;
    out[i] = c;
  }
  if (ndev > 1 && ndev <= EDM_PEER_MAX_RANKS) {  // one process: the windows are mapped through peer access
    std::vector<PeerCard> cards(ndev);
    bool ok = true;
    for (int i = 0; i < ndev; i++) ok = peer_alloc(out[i], &cards[i]) && ok;
    for (int i = 0; i < ndev && ok; i++) {
      char* peers[EDM_PEER_MAX_RANKS];
      cudaSetDevice(out[i]->device);
      ok = peer_map(out[i], cards.data(), peers) && peer_finish(out[i], peers);
    }
    if (!ok)
      for (int i = 0; i < ndev; i++) peer_release(out[i]);
    cudaGetLastError();
  }
  return EDM_OK;
}

// Wraps a communicator the host application already owns (an ncclComm_t passed as void*); not destroyed here.
int edm_comm_from_nccl(edm_comm_t** out, void* nccl_comm, int nranks, int rank, int device) {
  EDM_REQUIRE(out && (nccl_comm || nranks == 1) && nranks > 0 && rank >= 0 && rank < nranks, "bad argument");
  edm_comm* c = new edm_comm();
  c->nccl = nccl_comm;
  c->nranks = nranks;
  c->rank = rank;
  c->device = device;
  c->owned = 0;
  if (nranks > 1) {
    EDM_TRY(ensure_device(device));
    const NcclApi* a = nccl_api();
    if (a) peer_setup_collective(c, a);
  }
  *out = c;
  return EDM_OK;
}

int edm_comm_destroy(edm_comm_t* c) {
  if (!c) return EDM_OK;
  peer_release(c);
  if (c->owned && c->nccl) {
    const NcclApi* a = nccl_api();
    if (a) {
      cudaSetDevice(c->device);
      a->CommDestroy((ncclComm_t)c->nccl);
    }
  }
  delete c;
  return EDM_OK;
}

int edm_comm_info(const edm_comm_t* c, int* nranks, int* rank, int* device) {
  EDM_REQUIRE(c != nullptr, "NULL argument");
  if (nranks) *nranks = c->nranks;
  if (rank) *rank = c->rank;
  if (device) *device = c->device;
  return EDM_OK;
}

// 1 when the hill exchange of this communicator runs over NVLink peer windows (every rank reached every other
// rank's window), 0 when it is one ncclAllGather per exchange.  EDM_B200_NO_P2P=1 forces the latter.
int edm_comm_peer_windows(const edm_comm_t* c, int* enabled) {
  EDM_REQUIRE(c && enabled, "NULL argument");
  *enabled = c->p2p;
  return EDM_OK;
}

int edm_comm_group_start(void) {
  const NcclApi* a = nccl_api();
  if (!a) return EDM_ERR_COMM;
  EDM_NCCL(a, GroupStart());
  return EDM_OK;
}

int edm_comm_group_end(void) {
  const NcclApi* a = nccl_api();
  if (!a) return EDM_ERR_COMM;
  EDM_NCCL(a, GroupEnd());
  return EDM_OK;
}

// sum over ranks of n doubles in place (the bias energy, if the caller wants the job-wide scalar:
// LAMMPS reduces its own thermo energy, lammps/fix_edm.cpp:184)
int edm_comm_allreduce_sum_dev(edm_comm_t* c, double* buf, long n, void* stream) {
  EDM_REQUIRE(c && buf && n > 0, "bad argument");
  if (c->nranks == 1) return EDM_OK;
  EDM_TRY(ensure_device(c->device));
  const NcclApi* a = nccl_api();
  if (!a) return EDM_ERR_COMM;
  EDM_NCCL(a, AllReduce(buf, buf, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)c->nccl, (cudaStream_t)stream));
  return EDM_OK;
}

// From now on every whole-round entry point of `b` (add_hills, the pre/add/post triple, step_coords,
// pair_step_cells, pair_step_listed and their _dev forms) exchanges its accepted hills over `comm`
// before the limiter and takes its est_hill_count as this rank's share (the job-wide count is
// est * nranks: the reference divides hill_density by mpi_size_ instead, lib/edm_bias.cpp:175-180).
// cap = records per rank in the exchange block (0: 4096); the same on every rank.
int edm_bias_set_comm(edm_bias_t* b, edm_comm_t* comm, long cap) {
  EDM_REQUIRE(b != nullptr && cap >= 0, "bad argument");
  EDM_REQUIRE(!comm || comm->device == b->device, "communicator and bias live on different devices");
  b->comm = comm;
  b->comm_cap = cap > 0 ? cap : 4096;
  return EDM_OK;
}

int edm_bias_exchange_dev(edm_bias_t* b, edm_comm_t* comm, long cap, long long est_total, void* stream) {
  EDM_REQUIRE(b && comm && cap > 0, "bad argument");
  EDM_REQUIRE(comm->device == b->device, "communicator and bias live on different devices");
  EDM_TRY(ensure_device(b->device));
  const double* blocks = nullptr;
  EDM_TRY(exchange_gather(b, comm, cap, (cudaStream_t)stream, &blocks));
  return edm_bias_hills_commit_dev(b, blocks, comm->nranks, cap, est_total, stream);
}

// The same for n replicas driven by ONE thread (communicators from edm_comm_init_all): the all-gathers
// are grouped so that no call waits for a peer this thread has not reached yet; the commits follow.
int edm_bias_exchange_all_dev(int n, edm_bias_t** b, edm_comm_t** comm, long cap, long long est_total,
                              void** streams) {
  EDM_REQUIRE(n > 0 && b && comm && cap > 0, "bad argument");
  const double* blocks[64];
  EDM_REQUIRE(n <= 64, "too many replicas");
  if (n > 1) EDM_TRY(edm_comm_group_start());
  for (int i = 0; i < n; i++) {
    EDM_REQUIRE(b[i] && comm[i] && comm[i]->device == b[i]->device, "bad replica");
    EDM_TRY(ensure_device(b[i]->device));
    EDM_TRY(exchange_gather(b[i], comm[i], cap, streams ? (cudaStream_t)streams[i] : nullptr, &blocks[i]));
  }
  if (n > 1) EDM_TRY(edm_comm_group_end());
  for (int i = 0; i < n; i++) {
    EDM_TRY(ensure_device(b[i]->device));
    EDM_TRY(edm_bias_hills_commit_dev(b[i], blocks[i], comm[i]->nranks, cap, est_total,
                                      streams ? streams[i] : nullptr));
  }
  return EDM_OK;
}

}  // extern "C"
