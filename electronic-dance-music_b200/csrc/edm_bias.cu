// EDMBias step logic in HBM: batched force update (update_forces), candidate selection,
// height scaling, bias_per_step limiter, overflow backlog and hill log (add_hills).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "edm_host.h"

namespace edm {

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ------------------------------------------------------------------ K1: update_forces

// EDMBias::update_forces, lib/edm_bias.cpp:276-295: thread per atom, coalesced row loads,
// f -= dV/dx, energy reduced per CTA in a fixed order (partials summed in CTA order afterwards).
// Random coordinates make every evaluation a chain of dependent DRAM round trips (x -> corner
// records -> f): the old force is loaded together with the coordinate, before the grid walk, and
// the register budget is set for the occupancy that measured fastest on B200 (C3: 3 CTAs/SM
// 0.70 ms vs 0.81 / 0.86 ms at 2 / 4; C4: 2 CTAs/SM).  d_interp_cell issues all 2^DIM corner loads
// before it blends (spatially ordered atoms: C3 0.36 -> 0.31 ms, C4 1.15 -> 0.86 ms; unordered atoms
// are bound by the DRAM random-access rate either way: ~4.9 TB/s of mostly half-used 64 B bursts).
// Measured and rejected (tools/experiments/k1_variants.sh): two atoms per thread, a software pipeline
// with L2 prefetch of the next atom's corners (unordered 0.67 -> 1.14 ms), 64-register builds, loading the
// next trip's coordinates and old force one trip ahead (ordered C3 -3 %, C4 +5 %: spills).
#ifndef EDM_FORCES_UNROLL
#define EDM_FORCES_UNROLL 1
#endif
template <int DIM>
__global__ void __launch_bounds__(256, (DIM == 3) ? 2 : 3) forces_kernel(GridDesc g, long n, const double* __restrict__ x, long xs,
                                                     double* __restrict__ f, long fs, const int* __restrict__ mask,
                                                     int apply_mask, double* __restrict__ partial, BiasDev* st) {
  __shared__ double red[33];
  if (blockIdx.x == 0 && threadIdx.x == 0) st->stamp[14] = global_ns();  // measurement: when the force update began
  double e = 0.0;
  const long stride = (long)gridDim.x * blockDim.x;
  constexpr int U = EDM_FORCES_UNROLL;
  for (long i0 = (long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += U * stride) {
    double xi[U][DIM], fo[U][DIM], der[U][DIM];
    bool on[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const long i = i0 + u * stride;
      on[u] = i < n && !(apply_mask >= 0 && !(mask[i] & apply_mask));
      if (on[u]) {
#pragma unroll
        for (int d = 0; d < DIM; d++) {
          xi[u][d] = x[i * xs + d];
          fo[u][d] = f[i * fs + d];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      if (on[u]) e += d_eval_point<DIM>(g, xi[u], der[u], g.b_interp != 0);
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (on[u]) {
        const long i = i0 + u * stride;
#pragma unroll
        for (int d = 0; d < DIM; d++) f[i * fs + d] = fo[u][d] - der[u][d];
      }
    }
  }
  double tot = block_sum(e, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void sum_partials_kernel(int n, const double* __restrict__ partial, double* out, BiasDev* st) {
  __shared__ double red[33];
  if (st && threadIdx.x == 0) st->stamp[13] = global_ns();  // measurement: the force update has just finished
  double e = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) e += partial[i];
  double tot = block_sum(e, red);
  if (threadIdx.x == 0) out[0] = tot;
}

// ------------------------------------------------------------------ K4: selection

// add_hill's acceptance test, lib/edm_bias.cpp:543 (T17): runiform < hill_density / est.
// Accepted candidates are appended through one atomic counter and ordered by key afterwards, so
// the limiter sees them in candidate order whatever the launch geometry.
template <int DIM>
__global__ void __launch_bounds__(256) select_kernel(long n, const double* __restrict__ x, long xs,
                                                     const double* __restrict__ runiform, const int* __restrict__ mask,
                                                     int apply_mask, double thresh, int accept_all, uint64_t key,
                                                     uint64_t first_counter, BiasDev* st, HillAccepted* acc,
                                                     long cap) {
  pdl_trigger();
  pdl_wait();
  if (blockIdx.x == 0 && threadIdx.x == 0) st->stamp2[0] = global_ns();
  long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (apply_mask >= 0 && !(apply_mask & mask[i])) continue;
    bool take = accept_all;
    if (!take) {
      double u = runiform ? runiform[i] : uniform_from_key(key, first_counter + (uint64_t)i);
      take = u < thresh;
    }
    if (take) {
      int slot = atomicAdd(&st->n_accepted, 1);
      if (slot < cap) {
        acc[slot].key = first_counter + (uint64_t)i;
#pragma unroll
        for (int d = 0; d < DIM; d++) acc[slot].x[d] = x[i * xs + d];
      } else {
        st->accepted_overflow = 1;
      }
    }
  }
  if (threadIdx.x == 0) atomicMax(&st->stamp2[1], global_ns());  // measurement: when the last CTA left
}

// ------------------------------------------------------------------ K4: the hill round

struct RoundParams {
  int b_tempering, b_targeting;
  double global_tempering, bias_factor, boltzmann_factor;
  double hill_prefactor, bias_per_step, hill_density, expected_target, total_volume;
  long long est_hill_count;
  long accepted_cap, log_cap;
};

// GaussGrid::add_value for one hill by the whole CTA, lib/gaussian_grid.h:176-372
template <int DIM>
__device__ double cta_deposit(const GridDesc& g, const double* x0, double h, double* red, int* sflag, AxisEntry* axis) {
  constexpr int W = RecW<DIM>::value;
  if (threadIdx.x == 0) *sflag = 0;
  bool dirty;
  double tot = cta_window_pass<DIM, kPassStore>(g, x0, h, red, dirty, axis);
  if (dirty) *sflag = 1;
  __syncthreads();
  if (*sflag) {  // duplicate_boundary, lib/gaussian_grid.h:365-368
    for (int k = threadIdx.x; k < g.n_dup; k += blockDim.x)
      g.rec[g.dup_pairs[2 * k] * W] = g.rec[g.dup_pairs[2 * k + 1] * W];
  }
  __syncthreads();
  return tot;
}

// cv_hist_->add_value(position, v), lib/grid.h:370-385 (T21)
template <int DIM, bool ATOMIC>
__device__ __forceinline__ void d_hist_bump(const GridDesc& hist, const double* pos, double v) {
  if (!hist.rec) return;
  long long lin = 0, pstride = 1;
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    double xd = pos[d];
    if (!hist.periodic[d] && (xd < hist.min[d] || xd >= hist.upper[d])) return;
    if (hist.periodic[d]) xd = d_wrap(xd, hist.min[d], hist.len[d]);
    long long idx = (long long)floor(__ddiv_rn(__dsub_rn(xd, hist.min[d]), hist.dx[d]));
    idx = idx < 0 ? 0 : (idx > hist.n[d] - 1 ? hist.n[d] - 1 : idx);
    lin += idx * pstride;
    pstride *= hist.n[d];
  }
  if (ATOMIC)
    atomicAdd(&hist.rec[lin * hist.rec_w], v);  // +-1.0 adds are exact: the order does not matter
  else
    hist.rec[lin * hist.rec_w] += v;
}

// The scalars of BiasDev a round reads and writes, held in shared memory while the in-order kernel
// runs: thread 0 walks the round's state machine on this copy (a dependent global load per field
// per hill was most of the kernel's time) and writes it back once at the end.
struct RoundState {
  double cum_bias, temp_hill_cum;
  long long steps, left, right;
  int hills_added, log_n, log_dropped, backlog_full, skip;
};

// output_hill, lib/edm_bias.cpp:586-612 (thread 0 only)
template <int DIM>
__device__ void log_event(RoundState& rs, const GridDesc& hist, edm_hill_event_t* log, long log_cap, const double* pos,
                          double height, double bias_added, int type, double total_volume) {
  if (rs.log_n < log_cap) {
    edm_hill_event_t& e = log[rs.log_n++];
    e.steps = rs.steps;
    e.type = type;
    e.hills_added = rs.hills_added;
    for (int d = 0; d < 3; d++) e.pos[d] = d < DIM ? pos[d] : 0.0;
    e.height = height;
    e.bias_added = bias_added;
    e.cum_over_vol = rs.cum_bias / total_volume;
  } else {
    rs.log_dropped++;
  }
  double v = (type == 'b' || type == 'h' || type == 'n') ? 1.0 : ((type == 'u' || type == 'v') ? -1.0 : 0.0);
  if (v != 0.0) d_hist_bump<DIM, true>(hist, pos, v);  // a RED: nothing waits for the old count
}

// Orders the accepted candidates by key (= candidate order).  Keys are unique.  Rank sort for
// the usual few hundred entries, bitonic network in global memory beyond that.
__device__ void cta_sort_accepted(HillAccepted* a, HillAccepted* tmp, int n) {
  if (n <= 1) return;
  if (n <= 4096) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      unsigned long long k = a[i].key;
      int rank = 0;
      for (int j = 0; j < n; j++) rank += (a[j].key < k);
      tmp[rank] = a[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) a[i] = tmp[i];
    __syncthreads();
    return;
  }
  // bitonic network over np2 real slots: the buffer capacity is a power of two (ensure_accepted),
  // the tail is padded with maximal keys that sort to the end
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  for (int i = n + threadIdx.x; i < np2; i += blockDim.x) a[i].key = ~0ULL;
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        int l = i ^ j;
        if (l > i) {
          bool up = ((i & k) == 0);
          unsigned long long ki = a[i].key, kl = a[l].key;
          if ((ki > kl) == up && ki != kl) {
            HillAccepted t = a[i];
            a[i] = a[l];
            a[l] = t;
          }
        }
      }
      __syncthreads();
    }
  }
}

// pre_add_hill -> add_hill... -> post_add_hill for one round, in the reference's order
// (lib/edm_bias.cpp:413-442, 528-563, 444-526, 313-380, 565-583).  A single CTA walks the
// sequence; the work inside each step (a hill's window, the sort) is spread over its threads.
// Sequential on purpose: with local well-tempering hill k's height reads the bias left by hills
// < k, and the limiter is a running sum with an undo ("Hard parts").
// round_mode on entry: 0 = the whole round is this kernel's; 2 = the parallel round committed it;
// 3 = the parallel round committed entries [0, n_fast) of its list (n_plan_b backlog slots, then the
// candidates) — every one a plain full deposit — and this kernel resumes at entry n_fast, the first
// at which the running sum reaches bias_per_step.
template <int DIM>
__global__ void __launch_bounds__(512) hill_round_kernel(GridDesc bias, GridDesc hist, GridDesc target,
                                                          RoundParams prm, BiasDev* st, HillAccepted* acc,
                                                          HillAccepted* acc_tmp, edm_hill_event_t* log) {
  __shared__ double red[33];
  __shared__ int sflag;
  __shared__ double s_pos[3];
  __shared__ double s_h;
  __shared__ int s_go, s_tail;
  __shared__ double s_prefactor;
  __shared__ RoundState rs;
  __shared__ AxisEntry s_axis[DIM > 1 ? DIM * kAxisMax : 1];
  pdl_trigger();
  if (threadIdx.x == 0) st->stamp[15] = global_ns();  // measurement: resident, about to wait for the predecessor
  pdl_wait();
  const int mode = st->round_mode;
  if (threadIdx.x == 0) st->stamp[11] = global_ns();
  if (mode == 2) {  // the parallel round already committed this round
    if (threadIdx.x == 0) {
      st->stamp[12] = global_ns();
      st->n_accepted_last = st->n_accepted;
      st->n_accepted = 0;  // the next selection starts a fresh candidate list
      st->accepted_sorted = 0;
    }
    return;
  }
  const int W1 = DIM + 1;
  const bool t0 = threadIdx.x == 0;
  const double bps = prm.bias_per_step;

  // ---- pre_add_hill
  if (t0) {
    rs.cum_bias = st->cum_bias;
    rs.steps = st->steps;
    rs.left = st->left;
    rs.right = st->right;
    rs.log_n = st->log_n;
    rs.log_dropped = st->log_dropped;
    rs.backlog_full = st->backlog_full;
    rs.temp_hill_cum = 0.0;
    rs.hills_added = mode == 3 ? st->hills_added : 0;
    rs.skip = 0;
    double pf = prm.hill_prefactor;
    if (prm.global_tempering > 0) {  // T15: threshold tempering
      double avg = rs.cum_bias / prm.total_volume;
      if (avg >= prm.global_tempering)
        pf *= exp(-(avg - prm.global_tempering) /
                  (prm.global_tempering * (prm.bias_factor - 1) * prm.boltzmann_factor));
    }
    s_prefactor = pf;
  }
  __syncthreads();

  // ---- flush_bias_buffer(bias_per_step); in mode 3 the first n_fast slots are already drained
  double drained = mode == 3 ? st->temp_hill_cum : 0.0;  // meaningful on thread 0
  while (true) {
    if (t0) {
      s_go = (rs.left < rs.right) ? 1 : 0;
      if (s_go) {
        const double* slot = &st->overflow[rs.left * W1];
        for (int d = 0; d < DIM; d++) s_pos[d] = slot[d];
        s_h = slot[DIM];
      }
    }
    __syncthreads();
    if (!s_go) break;
    double pos[DIM];
    for (int d = 0; d < DIM; d++) pos[d] = s_pos[d];
    double h = s_h;
    double temp = cta_deposit<DIM>(bias, pos, h, red, &sflag, s_axis);
    if (t0) {
      rs.hills_added++;
      drained += temp;
      log_event<DIM>(rs, hist, log, prm.log_cap, pos, h, temp, 'b', prm.total_volume);
      s_go = 0;
      if (drained > bps) {
        double hh = fmax(bps - drained, -h);  // T20
        st->overflow[rs.left * W1 + DIM] = -hh;
        s_h = hh;
        s_go = 1;
      }
    }
    __syncthreads();
    if (s_go) {
      double hh = s_h;
      double t2 = cta_deposit<DIM>(bias, pos, hh, red, &sflag, s_axis);
      if (t0) {
        log_event<DIM>(rs, hist, log, prm.log_cap, pos, hh, t2, 'v', prm.total_volume);
        rs.hills_added++;
        drained += t2;
      }
      break;  // uniform: s_go is shared
    }
    if (t0) rs.left++;
    __syncthreads();
  }
  __syncthreads();
  if (t0) {
    if (rs.left == rs.right) rs.left = rs.right = 0;
    rs.temp_hill_cum += drained;
    rs.skip = (rs.left == 0 && rs.right == 0) ? 0 : 1;  // T18
  }
  __syncthreads();

  // ---- add_hill for every accepted candidate, in candidate order
  int nacc = st->n_accepted;
  if (nacc > prm.accepted_cap) nacc = (int)prm.accepted_cap;
  int k_first = mode == 3 ? st->n_fast - st->n_plan_b : 0;  // planned entries: backlog slots, then candidates
  if (k_first < 0) k_first = 0;
  if (!rs.skip && nacc > k_first) {
    if (mode == 0 && !st->accepted_sorted) cta_sort_accepted(acc, acc_tmp, nacc);  // mode 3: the plan ordered them
    double nxt[DIM];
    for (int d = 0; d < DIM; d++) nxt[d] = acc[k_first].x[d];
    for (int k = k_first; k < nacc; k++) {
      double pos[DIM];
      for (int d = 0; d < DIM; d++) pos[d] = nxt[d];
      if (k + 1 < nacc)  // the next centre travels while this hill is worked on
        for (int d = 0; d < DIM; d++) nxt[d] = acc[k + 1].x[d];
      if (t0) {
        double this_h = s_prefactor;
        if (prm.b_targeting) this_h *= exp(d_get_value<DIM>(target, pos) - prm.expected_target);
        if (prm.b_tempering && prm.global_tempering < 0)  // T15: local well-tempering
          this_h *= exp(-d_get_value<DIM>(bias, pos) / ((prm.bias_factor - 1) * prm.boltzmann_factor));
        if (prm.hill_density < 0)
          this_h /= (double)(int)prm.est_hill_count;
        else
          this_h /= prm.hill_density;
        this_h = fmin(this_h, 1.0 * bps);  // BIAS_CLAMP, lib/edm_bias.h:14
        s_h = this_h;
        s_go = (rs.temp_hill_cum < bps) ? 1 : 0;
        // Once the sum has reached bias_per_step it cannot change any more: every remaining hill is
        // logged with height 0 and pushed whole (lib/edm_bias.cpp:492-523).  If they all fit to the
        // right of the backlog and into the log, the CTA does them at once.
        s_tail = (!s_go && rs.right + (nacc - k) <= EDM_BUFFER_SLOTS && rs.log_n + (nacc - k) <= prm.log_cap) ? 1 : 0;
      }
      __syncthreads();
      if (s_tail) {
        const double cov = rs.cum_bias / prm.total_volume;
        for (int i = k + threadIdx.x; i < nacc; i += blockDim.x) {
          double p[DIM];
          for (int d = 0; d < DIM; d++) p[d] = acc[i].x[d];
          double hh = s_prefactor;
          if (prm.b_targeting) hh *= exp(d_get_value<DIM>(target, p) - prm.expected_target);
          if (prm.b_tempering && prm.global_tempering < 0)
            hh *= exp(-d_get_value<DIM>(bias, p) / ((prm.bias_factor - 1) * prm.boltzmann_factor));
          if (prm.hill_density < 0)
            hh /= (double)(int)prm.est_hill_count;
          else
            hh /= prm.hill_density;
          hh = fmin(hh, 1.0 * bps);
          edm_hill_event_t& e = log[rs.log_n + (i - k)];
          e.steps = rs.steps;
          e.type = 'h';
          e.hills_added = rs.hills_added;
          for (int d = 0; d < 3; d++) e.pos[d] = d < DIM ? p[d] : 0.0;
          e.height = 0.0;
          e.bias_added = 0.0;
          e.cum_over_vol = cov;
          d_hist_bump<DIM, true>(hist, p, 1.0);
          double* slot = &st->overflow[(rs.right + 1 + (i - k)) * W1];  // T19: the index moves before the write
          for (int d = 0; d < DIM; d++) slot[d] = p[d];
          slot[DIM] = hh;
        }
        __syncthreads();
        if (t0) {
          rs.log_n += nacc - k;
          rs.right += nacc - k;
        }
        break;
      }
      double this_h = s_h;
      int buffer_flag = 0;  // thread 0
      if (s_go) {
        double ba = cta_deposit<DIM>(bias, pos, this_h, red, &sflag, s_axis);
        if (t0) {
          rs.temp_hill_cum += ba;
          rs.hills_added++;
          log_event<DIM>(rs, hist, log, prm.log_cap, pos, this_h, ba, 'h', prm.total_volume);
          s_go = 0;
          if (rs.temp_hill_cum > bps) {
            s_h = fmax(bps - rs.temp_hill_cum, -this_h);  // T20
            s_go = 1;
          }
        }
        __syncthreads();
        if (s_go) {
          double temp_h = s_h;
          double ba2 = cta_deposit<DIM>(bias, pos, temp_h, red, &sflag, s_axis);
          if (t0) {
            rs.hills_added++;
            log_event<DIM>(rs, hist, log, prm.log_cap, pos, temp_h, ba2, 'u', prm.total_volume);
            rs.temp_hill_cum += ba2;
            buffer_flag = 1;
            this_h = -temp_h;
          }
        }
      } else if (t0) {
        log_event<DIM>(rs, hist, log, prm.log_cap, pos, 0.0, 0.0, 'h', prm.total_volume);
        buffer_flag = 1;
      }
      if (t0 && buffer_flag) {  // lib/edm_bias.cpp:498-523 incl. the off-by-one push (T19)
        if (rs.right == EDM_BUFFER_SLOTS) {
          if (rs.left == 0) {
            rs.backlog_full = 1;  // the reference aborts here
          } else {
            rs.left--;
            for (int d = 0; d < DIM; d++) st->overflow[rs.left * W1 + d] = pos[d];
            st->overflow[rs.left * W1 + DIM] = this_h;
          }
        } else {
          rs.right++;
          for (int d = 0; d < DIM; d++) st->overflow[rs.right * W1 + d] = pos[d];
          st->overflow[rs.right * W1 + DIM] = this_h;
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  // ---- post_add_hill (serial build: no exchange), update_height (T22)
  if (t0) {
    st->cum_bias = rs.cum_bias + rs.temp_hill_cum;
    st->temp_hill_cum = rs.temp_hill_cum;
    st->steps = rs.steps + 1;
    st->left = rs.left;
    st->right = rs.right;
    st->hills_added = rs.hills_added;
    st->skip = rs.skip;
    st->log_n = rs.log_n;
    st->log_dropped = rs.log_dropped;
    st->backlog_full = rs.backlog_full;
    if (mode == 3)
      st->rounds_split++;
    else
      st->rounds_in_order++;
    st->n_accepted_last = st->n_accepted;
    st->n_accepted = 0;  // the next selection starts a fresh candidate list
    st->accepted_sorted = 0;
    st->stamp[12] = global_ns();
  }
}

// ------------------------------------------------------------------ K4: the parallel hill round
//
// A round is a list of deposits in a fixed order: the backlog slots flush_bias_buffer drains first,
// then the accepted candidates.  As long as the running sum stays below bias_per_step every entry is
// a plain full deposit whose height is known up front, so the round splits into: plan (order the
// candidates, scale the heights) -> every entry's integral, all at once -> decide (the running sum,
// in list order: how long is the plain prefix?) -> deposit the prefix, all at once.  The first entry
// at which the sum reaches bias_per_step — the undo hill, the push to the backlog, the skipped rest —
// is left to the in-order kernel above, which resumes exactly there (round_mode 3), so the
// limiter/undo/backlog semantics are never re-expressed here.
//
// Local well-tempering (T15) makes hill k's height read the bias left by hills < k of the round.
// The plan resolves that exactly without depositing: the corner records of k's interpolation cell
// are patched with h_j * term_j(corner) for every earlier hill j whose window reaches a corner, in
// order j (the adds a sequential deposit would have made to those records), and k's height is
// interpolated from the patched corners.  Hills with no such predecessor are scaled in parallel;
// the others follow in order in one warp (a handful per round in 2-D/3-D, where windows are small
// against the grid).
//
// 1-D: the deposit is the owner-computes kernel staged into scratch (edm_grid.cu) and committed
// after the decision; 2-D/3-D: one CTA per hill, integrals first, fp64 REDs after the decision.

template <int DIM> struct PlanHill {
  HillGeom<DIM> hg;
  int ok;            // d_hill_prepare succeeded: the hill deposits something
  int valid;         // the centre lies inside the grid: get_value interpolates (else 0)
  int lo[DIM], up[DIM];
  double X0[DIM];
  double hb;         // height before the local tempering factor and the density division
  double rec[1 << DIM][RecW<DIM>::value];
};

// can hill j's window (centre cell xi, half-width minisize) write grid index i of dim d?
// mirrors d_window_index (lib/gaussian_grid.h:229-268)
__device__ __forceinline__ bool d_window_reaches(const GridDesc& g, int d, int xi, int i) {
  const int m = g.minisize[d];
  int off = i - xi;
  if (!g.periodic[d]) return off >= -m && off <= m;
  const int n = g.n[d];
  off %= n;
  if (off < 0) off += n;                 // off in [0, n): i = xi + off (mod n)
  if (off <= m) return true;             // reached from above the centre (xi + off may wrap once or more)
  return off - n >= -m && xi + off - n >= -n;  // from below: the reference adds n once only
}

// Cheap superset of "hill j's window can write a corner of hill k's interpolation cell", for the
// O(n^2) scan: cj = hill j's centre cell (periodic dims folded into [0, n)) and ok flag, lo = the low
// corner of k's cell.  A hill adds nothing beyond supp cells from its centre and the upper corner
// is one cell further; the exact per-corner test (d_window_reaches) follows for the survivors.
template <int DIM>
__device__ __forceinline__ bool d_hill_near(const GridDesc& g, const int4& cj, const int* lo) {
  if (!cj.w) return false;
  const int xi[3] = {cj.x, cj.y, cj.z};
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    int dist = xi[d] - lo[d];
    dist = dist < 0 ? -dist : dist;
    if (g.periodic[d]) dist = min(dist, g.n[d] - dist);
    if (dist > g.supp[d] + 1) return false;  // beyond supp cells the support test fails: nothing is added
  }
  return true;
}

__device__ __forceinline__ int d_fold(int i, int n, bool periodic) {
  if (!periodic) return i;
  i %= n;
  return i < 0 ? i + n : i;
}

template <int DIM>
__device__ __forceinline__ double d_local_height(const GridDesc& bias, const RoundParams& prm, const double* X0,
                                                 const double (*rec)[RecW<DIM>::value], bool valid, double hb) {
  double v = 0.0;
  if (valid) {
    if (bias.b_interp && bias.b_deriv) {
      double der[DIM];
#pragma unroll
      for (int d = 0; d < DIM; d++) der[d] = 0.0;
      v = d_interp_cell<DIM>(bias, X0, der, [&](int c, double* r) {
#pragma unroll
        for (int m = 0; m < RecW<DIM>::value; m++) r[m] = rec[c][m];
      });
    } else {
      v = rec[0][0];
    }
  }
  double h = hb * exp(-v / ((prm.bias_factor - 1) * prm.boltzmann_factor));
  if (prm.hill_density < 0)
    h /= (double)(int)prm.est_hill_count;
  else
    h /= prm.hill_density;
  return fmin(h, 1.0 * prm.bias_per_step);  // BIAS_CLAMP, lib/edm_bias.h:14
}

// The plan executes each of its phases once per round, on instruction caches the force update has just flushed:
// with everything inlined it was 120 KB (2-D) to 213 KB (3-D) of straight-line code and spent most of its time
// fetching it (sm__icc hit rate 76 %, 3.2 warps stalled on "no instruction" per issue).  The heavy device
// functions are therefore called, not inlined, here: one copy each, shared by all call sites and corners.
template <int DIM>
__device__ __noinline__ bool plan_hill_prepare(const GridDesc& g, const double* x0, HillGeom<DIM>& hg) {
  return d_hill_prepare<DIM>(g, x0, hg);
}
template <int DIM>
__device__ __noinline__ bool plan_hill_term(const GridDesc& g, const HillGeom<DIM>& hg, const int* idx, double& etot,
                                            double* force) {
  bool cnz;
  return d_hill_term<DIM>(g, hg, idx, etot, force, cnz);
}
template <int DIM>
__device__ __noinline__ double plan_local_height(const GridDesc& bias, const RoundParams& prm, const double* X0,
                                                 const double (*rec)[RecW<DIM>::value], bool valid, double hb) {
  return d_local_height<DIM>(bias, prm, X0, rec, valid, hb);
}
template <int DIM> __device__ __noinline__ bool plan_locate(const GridDesc& g, const double* x, CellLoc<DIM>& L) {
  return d_locate<DIM>(g, x, L);
}
template <int DIM> __device__ __noinline__ double plan_target_value(const GridDesc& target, const double* x) {
  return d_get_value<DIM>(target, x);
}

// The plan runs as ONE thread-block cluster of kPlanCtas CTAs (hardware cluster barriers between its phases, all
// exchange through global memory, which the barrier's release/acquire covers cluster-wide): as a single CTA it
// was a latency chain of 62-108 us at 8 % warp occupancy (profiles/r01_k_plan_c3_ncu_selected.txt) and the
// longest kernel of every 2-D/3-D round.  Phases:
//   0  every CTA derives the round's mode and the threshold-tempered prefactor from the same state words;
//      rank 0 publishes them.  Hills arrive either in the accepted buffer (selection on this device) or as the
//      rank-major blocks of the exchange (`blocks` != NULL): then the unpack is folded in here.
//   1  (accepted buffer only) rank sort by key, one candidate per thread across the cluster.
//   2  one entry per thread: centre, target scaling, hill geometry, interpolation cell and its corner records.
//   3  (local tempering) a warp per entry counts the earlier entries that can reach its corners; entries with
//      none take their height from the start-of-round records at once.
//   4  (local tempering) list offsets (every CTA scans the counts redundantly), then a warp per dependent entry
//      lists the per-unit-height terms of its predecessors.
//   5  (local tempering) one warp walks the dependent entries in list order.
constexpr int kPlanCtas = 8;

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_cta_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

template <int DIM>
__global__ void __cluster_dims__(kPlanCtas, 1, 1) __launch_bounds__(512)
    round_plan_kernel(const __grid_constant__ GridDesc bias, const __grid_constant__ GridDesc target,
                      const __grid_constant__ RoundParams prm, int n_max, BiasDev* st, HillAccepted* acc,
                      HillAccepted* acc_tmp, double* __restrict__ centres, double* heights, PlanHill<DIM>* plan,
                      int4* __restrict__ cells, int4* __restrict__ cells_folded, int* __restrict__ ndep_g,
                      double* terms, int* term_j, int term_cap, const double* __restrict__ blocks, int nblocks,
                      long block_cap) {
  constexpr int W = RecW<DIM>::value;
  constexpr int NC = 1 << DIM;
  __shared__ double s_prefactor;
  __shared__ int s_mode, s_nb, s_nacc, s_total;
  // where entry k's list of reaching predecessors starts (16 bits: lists hold at most term_cap <= 65535 entries,
  // and a round with more bails out before any offset is used)
  __shared__ unsigned short s_off[EDM_ROUND_MAX];
  __shared__ int4 s_cells[EDM_ROUND_MAX];          // folded centre cell + ok of every planned entry
  __shared__ unsigned short s_ndep[EDM_ROUND_MAX]; // earlier entries that can reach entry k's corners
  __shared__ int s_boff[65];                       // exchange blocks: first entry of each block
  __shared__ unsigned char s_ready[EDM_ROUND_MAX]; // phase 5: hill k's height is final
  __shared__ int s_ndeps;                          // phase 5: how many hills have predecessors
  __shared__ double s_rec5[16][NC * W];            // phase 5: one patched corner block per warp
  const bool local = prm.b_tempering && prm.global_tempering < 0;
  const int W1 = DIM + 1;
  const unsigned crank = cluster_cta_rank();
  const int ctid = (int)crank * blockDim.x + threadIdx.x, cthreads = kPlanCtas * blockDim.x;
  const int lane = threadIdx.x & 31;
  const int cwarp = ctid >> 5, cwarps = cthreads >> 5;
  const size_t bw = 1 + (size_t)block_cap * DIM;

  pdl_wait();  // the trigger is left to the exit: the successor's grid is large and would sit resident on every SM
  const bool stamper = crank == 0 && threadIdx.x == 0;
  if (stamper) st->stamp[0] = global_ns();
  // ---- phase 0
  if (threadIdx.x == 0) {
    const long long nb = st->right - st->left;
    long long nacc = st->n_accepted;
    int overflow = st->accepted_overflow;
    if (blocks) {  // the rank-major concatenation is the candidate list; its order is the canonical order
      nacc = 0;
      for (int b = 0; b < nblocks; b++) {
        s_boff[b] = (int)nacc;
        nacc += (long long)blocks[(size_t)b * bw];
      }
      s_boff[nblocks] = (int)nacc;
      if (nacc > prm.accepted_cap) overflow = 1;
    }
    int mode = (!overflow && nb + nacc <= n_max) ? 1 : 0;
    // a periodic window wider than the grid revisits its own points: concurrent hills (DIM > 1) and the plan's
    // one-term-per-predecessor corner patches (local tempering) both assume a single visit
    if (bias.dup_possible && (DIM > 1 || local)) mode = 0;
    if (local && bias.n_dup > 0) mode = 0;         // duplicate_boundary rewrites records between hills
    s_mode = mode;
    s_nb = (int)nb;
    s_nacc = (int)(nacc > prm.accepted_cap ? prm.accepted_cap : nacc);
    double pf = prm.hill_prefactor;
    if (prm.global_tempering > 0) {  // T15: threshold tempering, lib/edm_bias.cpp:419-426
      double avg = st->cum_bias / prm.total_volume;
      if (avg >= prm.global_tempering)
        pf *= exp(-(avg - prm.global_tempering) / (prm.global_tempering * (prm.bias_factor - 1) * prm.boltzmann_factor));
    }
    s_prefactor = pf;
  }
  __syncthreads();
  const int nb = s_nb, nacc = s_nacc, nall = nb + nacc;
  const int mode0 = s_mode;
  // every CTA has read the state words by now; only then may rank 0 change them
  cluster_sync_all();
  if (crank == 0 && threadIdx.x == 0) {
    st->n_plan_b = nb;
    st->round_mode = mode0;
    st->ticket = 0;
    st->int_done = 0;
    st->round_epoch++;
    st->n_fast = mode0 ? nall : 0;
    if (blocks) {
      st->n_accepted = s_boff[nblocks];
      st->accepted_sorted = 1;  // keys are the positions in the rank-major concatenation
      if (s_boff[nblocks] > prm.accepted_cap) st->accepted_overflow = 1;
    }
  }
  if (blocks) {  // the unpack: the in-order kernel (and the next plan phases) read the accepted buffer
    for (int i = ctid; i < nacc; i += cthreads) {
      int b = 0;
      while (b + 1 < nblocks && s_boff[b + 1] <= i) b++;
      const double* src = blocks + (size_t)b * bw + 1 + (size_t)(i - s_boff[b]) * DIM;
      HillAccepted a;
      a.key = (unsigned long long)i;
#pragma unroll
      for (int d = 0; d < 3; d++) a.x[d] = d < DIM ? src[d] : 0.0;
      acc[i] = a;
    }
  }
  if (mode0 == 0) return;  // uniform across the cluster: the in-order kernel takes the whole round
  if (stamper) st->stamp[1] = global_ns();

  // ---- phase 1: candidate order
  if (!blocks) {
    const bool sorted = st->accepted_sorted != 0;  // not modified by this kernel on this path
    if (!sorted) {
      for (int i = ctid; i < nacc; i += cthreads) {
        const unsigned long long k = acc[i].key;
        int rank = 0;
        for (int j = 0; j < nacc; j++) rank += (acc[j].key < k);
        acc_tmp[rank] = acc[i];
      }
      cluster_sync_all();
      for (int i = ctid; i < nacc; i += cthreads) acc[i] = acc_tmp[i];
    }
  }
  cluster_sync_all();
  if (stamper) st->stamp[2] = global_ns();

  // ---- phase 2: one entry per thread
  const long long left = st->left;
  for (int k = ctid; k < nall; k += cthreads) {
    double pos[DIM];
    const bool slot = k < nb;  // a backlog slot: its height is what the slot holds
    const double* src = slot ? &st->overflow[(left + k) * W1] : acc[k - nb].x;
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      pos[d] = src[d];
      centres[(long)k * DIM + d] = pos[d];
    }
    double h = s_prefactor;
    if (slot)
      heights[k] = src[DIM];
    else if (prm.b_targeting)
      h *= exp(plan_target_value<DIM>(target, pos) - prm.expected_target);
    HillGeom<DIM> hg;
    bool ok = false;
    if (DIM > 1 || local) {  // centre cells for the deposit's overlap test / the reach tests of local tempering
      ok = plan_hill_prepare<DIM>(bias, pos, hg);
      if (DIM > 1) cells[k] = make_int4(hg.xi[0], hg.xi[DIM > 1 ? 1 : 0], hg.xi[DIM > 2 ? 2 : 0], ok ? 1 : 0);
    }
    if (!local) {
      if (prm.hill_density < 0)
        h /= (double)(int)prm.est_hill_count;
      else
        h /= prm.hill_density;
      if (!slot) heights[k] = fmin(h, 1.0 * prm.bias_per_step);
    } else {
      PlanHill<DIM>& p = plan[k];
      p.hb = h;
      p.ok = ok ? 1 : 0;
      p.hg = hg;
      const int4 c = make_int4(hg.xi[0], hg.xi[DIM > 1 ? 1 : 0], hg.xi[DIM > 2 ? 2 : 0], p.ok);
      cells_folded[k] = make_int4(d_fold(c.x, bias.n[0], bias.periodic[0] != 0),
                                  d_fold(c.y, bias.n[DIM > 1 ? 1 : 0], bias.periodic[DIM > 1 ? 1 : 0] != 0),
                                  d_fold(c.z, bias.n[DIM > 2 ? 2 : 0], bias.periodic[DIM > 2 ? 2 : 0] != 0), p.ok);
      CellLoc<DIM> L;
      p.valid = plan_locate<DIM>(bias, pos, L) ? 1 : 0;
      if (p.valid) {
#pragma unroll
        for (int d = 0; d < DIM; d++) {
          p.lo[d] = L.idx[d];
          p.up[d] = (bias.periodic[d] && L.idx[d] == bias.n[d] - 1) ? 0 : L.idx[d] + 1;
          p.X0[d] = L.X0[d];
        }
#pragma unroll
        for (int c2 = 0; c2 < NC; c2++) RecLoad<W>::ld(bias.rec + (L.base + d_corner_shift<DIM>(L, c2)) * W, p.rec[c2]);
      }
    }
  }
  if (stamper) st->stamp[3] = st->stamp[4] = st->stamp[5] = st->stamp[6] = global_ns();
  if (!local) return;
  cluster_sync_all();
  if (stamper) st->stamp[3] = global_ns();

  // ---- phase 3: a warp per new hill counts the earlier entries that can reach its corners
  for (int k = threadIdx.x; k < nall; k += blockDim.x) s_cells[k] = cells_folded[k];
  __syncthreads();
  for (int k = nb + cwarp; k < nall; k += cwarps) {
    const PlanHill<DIM>& p = plan[k];
    int ndep = 0;
    if (p.valid) {
      int lo[DIM];
#pragma unroll
      for (int d = 0; d < DIM; d++) lo[d] = p.lo[d];
      for (int j0 = 0; j0 < k; j0 += 32) {
        const int j = j0 + lane;
        const bool near = j < k && d_hill_near<DIM>(bias, s_cells[j], lo);
        ndep += __popc(__ballot_sync(0xffffffffu, near));
      }
    }
    if (lane == 0) {
      ndep_g[k] = ndep;
      // nobody earlier can reach it: height from the start-of-round records
      if (ndep == 0) heights[k] = plan_local_height<DIM>(bias, prm, p.X0, p.rec, p.valid != 0, p.hb);
    }
  }
  cluster_sync_all();
  if (stamper) st->stamp[4] = global_ns();

  // ---- phase 4: list offsets (each CTA for itself), then the per-unit-height terms, a warp per dependent hill
  for (int k = threadIdx.x; k < nall; k += blockDim.x) s_ndep[k] = k < nb ? (unsigned short)0 : (unsigned short)ndep_g[k];
  __syncthreads();
  if (threadIdx.x < 32) {
    int carry = 0;
    for (int base = nb; base < nall; base += 32) {
      const int k = base + lane;
      const int v = k < nall ? (int)s_ndep[k] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      if (k < nall) s_off[k] = (unsigned short)(carry + inc - v);
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) s_total = carry;
  }
  __syncthreads();
  if (s_total > term_cap) {  // more interplay than the lists hold: the in-order kernel takes the round
    if (crank == 0 && threadIdx.x == 0) {
      st->round_mode = 0;
      st->n_fast = 0;
    }
    return;  // uniform across the cluster: every CTA computed the same total
  }
  for (int k = nb + cwarp; k < nall; k += cwarps) {
    if (s_ndep[k] == 0) continue;
    const PlanHill<DIM>& p = plan[k];
    int lo[DIM], up[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      lo[d] = p.lo[d];
      up[d] = p.up[d];
    }
    int pos = s_off[k];
    for (int j0 = 0; j0 < k; j0 += 32) {
      const int j = j0 + lane;
      const bool near = j < k && d_hill_near<DIM>(bias, s_cells[j], lo);
      const unsigned m = __ballot_sync(0xffffffffu, near);
      if (near) {
        const int slot = pos + __popc(m & ((1u << lane) - 1u));
        term_j[slot] = j;
        double* T = terms + (size_t)slot * (NC * W);
        const HillGeom<DIM> hgj = plan[j].hg;  // one bulk load instead of a dependent L2 round trip per field
#pragma unroll 1
        for (int c = 0; c < NC; c++) {
          int idx[DIM];
          bool reach = true;
#pragma unroll
          for (int d = 0; d < DIM; d++) {
            idx[d] = ((c >> d) & 1) ? up[d] : lo[d];
            reach = reach && d_window_reaches(bias, d, hgj.xi[d], idx[d]);
          }
          double etot = 0.0, force[DIM];
#pragma unroll
          for (int d = 0; d < DIM; d++) force[d] = 0.0;
          if (reach && !plan_hill_term<DIM>(bias, hgj, idx, etot, force)) {
            etot = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; d++) force[d] = 0.0;
          }
          T[c * W] = etot;
#pragma unroll
          for (int d = 0; d < DIM; d++) T[c * W + 1 + d] = force[d];
          if (W > DIM + 1) T[c * W + W - 1] = 0.0;
        }
      }
      pos += __popc(m);
    }
  }
  cluster_sync_all();
  if (stamper) st->stamp[5] = global_ns();
  if (crank != 0) return;

  // ---- phase 5 (rank 0): the dependent hills, each by one warp: corner records + sum over its list of
  // h_j * term in list order (the adds a hill-by-hill deposit would have made to those records, in the same
  // order), interpolate, scale.  A hill waits only for listed predecessors, which all have lower indices; a warp
  // takes its hills in increasing order, so the lowest unfinished hill is always being worked on by its warp and
  // never waits for an unfinished one: no deadlock, and sparse dependencies (2-D/3-D) resolve in parallel while a
  // 1-D chain degrades to the order it must have anyway.
  if (threadIdx.x == 0) s_ndeps = 0;
  for (int k = threadIdx.x; k < nall; k += blockDim.x) s_ready[k] = (k < nb || s_ndep[k] == 0) ? 1 : 0;
  __syncthreads();
  // the dependent hills as a compact ascending list (rewriting s_ndep's tail is not possible: counts are needed),
  // kept in the cells_folded scratch of this CTA's shared copy: s_cells is dead from here on
  int* s_deplist = reinterpret_cast<int*>(s_cells);
  if (threadIdx.x < 32) {
    int base = 0;
    for (int k0 = nb; k0 < nall; k0 += 32) {
      const int k = k0 + lane;
      const bool dep = k < nall && s_ndep[k] != 0;
      const unsigned m = __ballot_sync(0xffffffffu, dep);
      if (dep) s_deplist[base + __popc(m & ((1u << lane) - 1u))] = k;
      base += __popc(m);
    }
    if (lane == 0) s_ndeps = base;
  }
  __syncthreads();
  {
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int ndeps = s_ndeps;
    for (int i = warp; i < ndeps; i += nwarps) {  // the i-th dependent hill goes to warp i % nwarps, ascending
      const int k = s_deplist[i];
      const int cnt = s_ndep[k];
      const PlanHill<DIM>& p = plan[k];
      double accv = 0.0;
      if (lane < NC * W) accv = p.rec[lane / W][lane % W];
      const int first = s_off[k];
      for (int e0 = 0; e0 < cnt; e0 += 32) {
        const int e = e0 + lane;
        double hj = 0.0;
        if (e < cnt) {
          const int j = term_j[first + e];
          while (*(volatile unsigned char*)&s_ready[j] == 0) {
          }
          hj = __ldcg(&heights[j]);  // written by another warp of this CTA (or an earlier phase): read through L2
        }
        __syncwarp();
        const int m = min(32, cnt - e0);
        for (int q = 0; q < m; q++) {
          const double h = __shfl_sync(0xffffffffu, hj, q);
          const double t = lane < NC * W ? terms[(size_t)(first + e0 + q) * (NC * W) + lane] : 0.0;
          accv += h * t;
        }
      }
      double* myrec = &s_rec5[warp][0];
      if (lane < NC * W) myrec[lane] = accv;
      __syncwarp();
      if (lane == 0) {
        heights[k] = plan_local_height<DIM>(bias, prm, p.X0, reinterpret_cast<const double(*)[W]>(myrec), true, p.hb);
        __threadfence_block();  // consumers inside this phase are warps of this CTA; later kernels see it anyway
        *(volatile unsigned char*)&s_ready[k] = 1;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (stamper) st->stamp[6] = global_ns();
}

template <int DIM>
__device__ void round_decide(const GridDesc& hist, const RoundParams& prm, BiasDev* st, const double* __restrict__ centres,
                             const double* __restrict__ heights, const double* __restrict__ ba, edm_hill_event_t* log);

// add_value's integral of every planned hill, one CTA per hill; the CTA that finishes last also takes the
// decision (round_decide below) — one launch and one dependent round trip less than a separate kernel.
template <int DIM>
__global__ void __launch_bounds__(512, 2) round_integrals_kernel(GridDesc bias, GridDesc hist, RoundParams prm, BiasDev* st,
                                                              const double* __restrict__ centres,
                                                              const double* __restrict__ heights,
                                                              double* __restrict__ ba, edm_hill_event_t* log) {
  __shared__ double red[33];
  __shared__ int s_last;
  __shared__ AxisEntry s_axis[DIM > 1 ? DIM * kAxisMax : 1];
  pdl_trigger();
  pdl_wait();
  if (st->round_mode != 1) return;  // uniform over the grid: written by the plan, an earlier launch
  const int n = st->n_fast;
  for (int k = blockIdx.x; k < n; k += gridDim.x) {
    double pos[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) pos[d] = centres[(long)k * DIM + d];
    bool dirty;
    double tot = cta_window_pass<DIM, kPassIntegrate>(bias, pos, heights[k], red, dirty, s_axis);
    if (threadIdx.x == 0) ba[k] = tot;
  }
  // last CTA out decides: every CTA publishes its integrals (fence) before it signs off
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(&st->int_done, 1) == (int)gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) st->stamp[7] = global_ns();
  round_decide<DIM>(hist, prm, st, centres, heights, ba, log);
  if (threadIdx.x == 0) st->stamp[8] = global_ns();
}

// The deposit itself once the decision fell (round_mode 2 or 3: hills [0, n_fast)), 2-D/3-D.  CTAs take hills by ticket,
// in candidate order; before depositing, hill k waits for every earlier hill whose window can overlap
// its own.  A hill only ever waits for lower tickets, which running CTAs hold, so the wait cannot
// deadlock whatever the residency; every grid point receives its adds in candidate order — the
// result is bit-identical to the in-order kernel and the same on every replica.
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int DIM>
__device__ __forceinline__ bool d_windows_overlap(const GridDesc& g, const int4& a, const int4& b) {
  if (!a.w || !b.w) return false;
  const int ca[3] = {a.x, a.y, a.z}, cb[3] = {b.x, b.y, b.z};
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    int dist = ca[d] - cb[d];
    dist = dist < 0 ? -dist : dist;
    if (g.periodic[d]) {
      dist %= g.n[d];
      dist = min(dist, g.n[d] - dist);
    }
    if (dist > 2 * g.supp[d]) return false;  // a deposit pass touches supp cells either side of its centre cell, no more
  }
  return true;
}

template <int DIM>
__global__ void __launch_bounds__(512, 2) round_deposit_kernel(GridDesc bias, BiasDev* st,
                                                            const double* __restrict__ centres,
                                                            const double* __restrict__ heights,
                                                            const int4* __restrict__ cells, int* flags,
                                                            const double* __restrict__ energy_partial, int n_partials,
                                                            double* energy_out) {
  __shared__ double red[33];
  __shared__ int s_k;
  __shared__ AxisEntry s_axis[DIM > 1 ? DIM * kAxisMax : 1];
  pdl_trigger();
  pdl_wait();
  if (st->round_mode < 2) return;
  const int n = st->n_fast;
  const int epoch = st->round_epoch;
  while (true) {
    if (threadIdx.x == 0) s_k = atomicAdd(&st->ticket, 1);
    __syncthreads();
    const int k = s_k;
    if (k >= n) {
      if (threadIdx.x == 0 && k == n + (int)gridDim.x - 1) st->stamp[13] = global_ns();  // measurement: last CTA leaves
      // The first CTA left without a hill adds up the energy partials of this step's force update (CTA order, so
      // run-to-run deterministic): that sum rides along here instead of a one-CTA kernel of its own in the stream.
      if (k == n && energy_out) {
        double e = 0.0;
        for (int i = threadIdx.x; i < n_partials; i += blockDim.x) e += energy_partial[i];
        e = block_sum(e, red);
        if (threadIdx.x == 0) energy_out[0] = e;
      }
      break;
    }
    if (k == 0 && threadIdx.x == 0) st->stamp[9] = global_ns();
    const int4 ck = cells[k];
    for (int j = threadIdx.x; j < k; j += blockDim.x)
      if (d_windows_overlap<DIM>(bias, cells[j], ck))
        while (ld_acquire_gpu(&st->hill_done[j]) != epoch) {
        }
    __syncthreads();
    double pos[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) pos[d] = centres[(long)k * DIM + d];
    bool dirty;
    // concurrent hills never overlap here (an overlapping later hill waits for this one's release): plain RMWs
    cta_window_pass<DIM, kPassOrdered>(bias, pos, heights[k], red, dirty, s_axis);
    if (dirty) flags[0] = 1;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      if (k == n - 1) st->stamp[10] = global_ns();
      st_release_gpu(&st->hill_done[k], epoch);
    }
  }
}

template <int DIM>
__device__ void round_decide(const GridDesc& hist, const RoundParams& prm, BiasDev* st, const double* __restrict__ centres,
                             const double* __restrict__ heights, const double* __restrict__ ba, edm_hill_event_t* log) {
  // One CTA, a handful of dependent steps: every global load it needs is issued up front (the state words by
  // thread 0 while all threads fetch the integrals), the running sum runs on shared memory, and nothing is read
  // back from global memory afterwards — each dependent round trip here costs as much as the whole scan.
  __shared__ int s_take, s_nb, s_base;
  __shared__ double s_cov;
  __shared__ long long s_steps;
  __shared__ double s_ba[EDM_ROUND_MAX];  // the scan below is one thread's: keep its operands next to it
  const int n = st->n_fast;
  int nb = 0, log_n = 0;
  long long steps = 0, left = 0;
  double cum_bias = 0.0;
  if (threadIdx.x == 0) {
    nb = st->n_plan_b;
    log_n = st->log_n;
    steps = st->steps;
    left = st->left;
    cum_bias = st->cum_bias;
  }
  for (int k = threadIdx.x; k < n; k += blockDim.x) s_ba[k] = __ldcg(&ba[k]);  // other CTAs wrote them: read through L2
  __syncthreads();
  if (threadIdx.x == 0) {
    // Entries [0, take) of the planned list are plain full deposits.  Backlog slots first
    // (flush_bias_buffer, lib/edm_bias.cpp:313-380): a slot is plain unless the drained sum exceeds
    // bias_per_step after it.  Then, only if the backlog emptied (T18), the new hills (the limiter's
    // running sum, lib/edm_bias.cpp:465-474): plain while the sum stays below bias_per_step before
    // and after the hill.  The first entry that is not plain, and everything behind it, is the
    // in-order kernel's.
    const double bps = prm.bias_per_step;
    double cum = 0.0;
    int take = 0;
    bool stop = false;
    while (!stop && take < n) {  // eight entries at a time: the loads and tests overlap, the sum stays serial
      const int cnt = n - take < 8 ? n - take : 8;
      double c[9], lowest = 0.0;
      c[0] = cum;
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const double v = i < cnt ? s_ba[take + i] : 0.0;
        lowest = fmin(lowest, v);
        c[i + 1] = c[i] + v;
      }
      // The usual chunk: no negative entry (undo remainders in drained slots are the only ones) and the sum
      // still below bias_per_step after it — then every prefix is below it too and all eight entries are plain;
      // only the serial adds remain on the critical path.  Any other chunk is examined entry by entry.
      if (lowest >= 0.0 && c[8] < bps && c[0] < bps) {
        cum = c[8];  // entries past cnt added 0.0
        take += cnt;
        continue;
      }
      int first = cnt;
#pragma unroll
      for (int i = 7; i >= 0; i--) {
        const bool slot = take + i < nb;
        const bool plain = slot ? !(c[i + 1] > bps) : (c[i] < bps && c[i + 1] < bps);
        if (i < cnt && !plain) first = i;
      }
#pragma unroll
      for (int i = 0; i <= 8; i++)
        if (i == first) cum = c[i];
      take += first;
      stop = first < cnt;
    }
    if (log_n + take > prm.log_cap) take = 0;
    s_take = take;
    s_nb = nb;
    s_base = log_n;
    s_steps = steps;
    s_cov = cum_bias / prm.total_volume;
    if (take > 0) {
      st->temp_hill_cum = cum;
      st->hills_added = take;
      st->skip = 0;
      st->n_fast = take;
      if (take >= nb)
        st->left = st->right = 0;
      else
        st->left = left + take;
      st->log_n = log_n + take;
      if (take == n) {  // post_add_hill; otherwise the in-order kernel finishes the round
        st->cum_bias = cum_bias + cum;
        st->steps = steps + 1;
        st->round_mode = 2;
        st->rounds_parallel++;
      } else {
        st->round_mode = 3;
      }
    } else {
      st->round_mode = 0;
      st->n_fast = 0;
    }
  }
  __syncthreads();
  const int take = s_take;
  if (take == 0) return;
  const double cov = s_cov;
  const int base = s_base;
  const long long steps_now = s_steps;
  for (int k = threadIdx.x; k < take; k += blockDim.x) {
    edm_hill_event_t& e = log[base + k];
    e.steps = steps_now;
    e.type = k < s_nb ? 'b' : 'h';
    e.hills_added = k + 1;
    double pos[DIM];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      if (d < DIM) pos[d] = centres[(long)k * DIM + d];
      e.pos[d] = d < DIM ? pos[d] : 0.0;
    }
    e.height = heights[k];
    e.bias_added = s_ba[k];
    e.cum_over_vol = cov;
    d_hist_bump<DIM, true>(hist, pos, 1.0);
  }
}

__global__ void reset_mode_kernel(BiasDev* st) {
  pdl_trigger();
  pdl_wait();
  st->round_mode = 0;
}

__global__ void reset_accepted_kernel(BiasDev* st) {
  pdl_trigger();
  pdl_wait();
  st->n_accepted = 0;
  st->accepted_overflow = 0;
  st->accepted_sorted = 0;
}

// hill exchange blocks: double[0] = count, then `count` centres of DIM doubles (key order)
template <int DIM>
__global__ void pack_block_kernel(BiasDev* st, HillAccepted* acc, HillAccepted* tmp, double* block, long cap) {
  pdl_trigger();
  pdl_wait();
  int n = st->n_accepted;
  if (n > cap) {  // never drop hills silently: edm_bias_check / the next host-synchronising call reports it
    if (threadIdx.x == 0) st->accepted_overflow = 1;
    n = (int)cap;
  }
  if (!st->accepted_sorted) cta_sort_accepted(acc, tmp, n);
  if (threadIdx.x == 0) {
    block[0] = (double)n;
    st->accepted_sorted = 1;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    for (int d = 0; d < DIM; d++) block[1 + (long)i * DIM + d] = acc[i].x[d];
}

// The exchange over NVLink peer windows: pack_block_kernel's block goes straight into slot [parity][rank] of EVERY
// rank's window (this one included) with plain stores through the peer mappings, a system-scope release of
// flag[parity][rank] = epoch on each peer publishes it, and the kernel then waits (system-scope acquire) until the
// local window holds every rank's block of this epoch.  When it ends, the local window is the rank-major
// concatenation an all-gather would have produced.  Only `count` records travel, not the block's capacity.
// Two parities suffice: a rank can only be one exchange ahead of the slowest, because finishing exchange e needs
// every rank's block e, which a rank sends after it has consumed exchange e-1 (stream order).
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int DIM>
__global__ void __launch_bounds__(512) pack_push_kernel(BiasDev* st, HillAccepted* acc, HillAccepted* tmp, long cap,
                                                        char* const* __restrict__ peer, int nranks, int rank,
                                                        int parity, int epoch, unsigned long long timeout_ns) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) st->stamp2[2] = global_ns();
  __shared__ unsigned long long s_key[512];
  int n = st->n_accepted;
  if (n > cap) {
    if (threadIdx.x == 0) st->accepted_overflow = 1;
    n = (int)cap;
  }
  const size_t bw = 1 + (size_t)cap * DIM;  // the block stride the round's unpack uses
  const size_t slot = ((size_t)parity * nranks + rank) * bw;
  if (n <= (int)blockDim.x && n <= 512) {
    // the usual few hundred hills: one record per thread, held in registers; its rank among the keys (shared
    // memory, every thread reads the same word at the same time) is where it goes -- in every peer's window and in
    // the local list -- without a pass through global memory in between
    HillAccepted rec;
    const bool have = (int)threadIdx.x < n;
    if (have) {
      rec = acc[threadIdx.x];
      s_key[threadIdx.x] = rec.key;
    }
    __syncthreads();
    if (have) {
      int r = threadIdx.x;
      if (!st->accepted_sorted) {
        r = 0;
        for (int j = 0; j < n; j++) r += (s_key[j] < rec.key);
      }
      for (int p = 0; p < nranks; p++) {
        double* dst = reinterpret_cast<double*>(peer[p] + EDM_PEER_FLAG_BYTES) + slot + 1 + (size_t)r * DIM;
#pragma unroll
        for (int d = 0; d < DIM; d++) dst[d] = rec.x[d];
      }
      acc[r] = rec;  // every record was read before the barrier above
    }
    if (threadIdx.x == 0)
      for (int p = 0; p < nranks; p++) reinterpret_cast<double*>(peer[p] + EDM_PEER_FLAG_BYTES)[slot] = (double)n;
    __syncthreads();
    if (threadIdx.x == 0) st->accepted_sorted = 1;
  } else {
    if (!st->accepted_sorted) cta_sort_accepted(acc, tmp, n);
    if (threadIdx.x == 0) st->accepted_sorted = 1;
    for (int p = 0; p < nranks; p++) {
      double* dst = reinterpret_cast<double*>(peer[p] + EDM_PEER_FLAG_BYTES) + slot;
      if (threadIdx.x == 0) dst[0] = (double)n;
      for (int i = threadIdx.x; i < n * DIM; i += blockDim.x) dst[1 + i] = acc[i / DIM].x[i % DIM];
    }
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < nranks)
    st_release_sys(reinterpret_cast<int*>(peer[threadIdx.x]) + parity * EDM_PEER_MAX_RANKS + rank, epoch);
  if (threadIdx.x == 0) st->stamp2[3] = global_ns();
  if ((int)threadIdx.x < nranks) {
    const int* fl = reinterpret_cast<const int*>(peer[rank]) + parity * EDM_PEER_MAX_RANKS + threadIdx.x;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(fl) != epoch) {
      if (global_ns() - t0 > timeout_ns) {  // a peer died or fell out of step: report instead of hanging the GPU
        st->exchange_timeout = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) st->stamp2[4] = global_ns();
}

template <int DIM>
__global__ void unpack_blocks_kernel(BiasDev* st, HillAccepted* acc, long acc_cap, const double* blocks, int nblocks,
                                     long cap) {
  if (blockIdx.x != 0) return;
  __shared__ int s_off;
  size_t bw = 1 + (size_t)cap * DIM;
  int off = 0;
  for (int b = 0; b < nblocks; b++) {
    const double* blk = blocks + (size_t)b * bw;
    int cnt = (int)blk[0];
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      if (off + i < acc_cap) {
        acc[off + i].key = (unsigned long long)(off + i);  // rank-major order is the canonical order
        for (int d = 0; d < DIM; d++) acc[off + i].x[d] = blk[1 + (size_t)i * DIM + d];
      }
    }
    off += cnt;
  }
  if (threadIdx.x == 0) {
    s_off = off;
    st->n_accepted = s_off;
    st->accepted_sorted = 1;  // keys are the positions in the rank-major concatenation
    if (off > acc_cap) st->accepted_overflow = 1;
  }
}

}  // namespace edm

using namespace edm;

static RoundParams round_params(const edm_bias* b, long long est) {
  RoundParams r;
  const edm_bias_params_t& p = b->prm;
  r.b_tempering = p.b_tempering;
  r.b_targeting = p.b_targeting && b->target;
  r.global_tempering = p.global_tempering;
  r.bias_factor = p.bias_factor;
  r.boltzmann_factor = p.boltzmann_factor;
  r.hill_prefactor = p.hill_prefactor;
  r.bias_per_step = p.bias_per_step;
  r.hill_density = p.hill_density;
  r.expected_target = p.expected_target;
  r.total_volume = p.total_volume;
  r.est_hill_count = est;
  r.accepted_cap = b->accepted_cap;
  r.log_cap = b->log_cap;
  return r;
}

static int ensure_accepted(edm_bias* b, long need) {
  if (need <= b->accepted_cap) return EDM_OK;
  long cap = 4096;
  while (cap < need) cap <<= 1;  // power of two: cta_sort_accepted pads in place
  HillAccepted* p = nullptr;
  EDM_CUDA(cudaMalloc(&p, 2 * (size_t)cap * sizeof(HillAccepted)));
  if (b->d_accepted) {  // candidates of earlier batches of the same round stay (add_hill_batch, select_dev)
    EDM_CUDA(cudaMemcpy(p, b->d_accepted, (size_t)b->accepted_cap * sizeof(HillAccepted), cudaMemcpyDeviceToDevice));
    cudaFree(b->d_accepted);
  }
  b->d_accepted = p;
  b->accepted_cap = cap;
  return EDM_OK;
}

// Room for the candidates a round can accept: `candidates` proposals thinned with probability
// hill_density / est (lib/edm_bias.cpp:543).  With a poor est -- fix edm_pair starts from atom->nmax,
// lammps/fix_edm_pair.cpp:105 -- a round accepts many times hill_density; they must all reach the limiter,
// which then fills the backlog or fails the way the reference does (lib/edm_bias.cpp:503-507).
int edm_bias_size_accepted(edm_bias* b, double candidates, long long est) {
  double expect = candidates;
  if (b->prm.hill_density >= 0) {
    double p = (int)est > 0 ? b->prm.hill_density / (double)(int)est : 1.0;
    if (p > 1.0) p = 1.0;
    expect = candidates * p;
    expect += 6.0 * sqrt(expect) + 64.0;
  }
  if (expect > (double)(1 << 22)) expect = (double)(1 << 22);
  return ensure_accepted(b, (long)expect);
}

int edm_bias_reset_accepted(edm_bias* b, cudaStream_t st) {
  count_launches(1);
  EDM_CUDA(launch_pdl(reset_accepted_kernel, dim3(1), dim3(1), 0, st, b->d_state));
  EDM_CUDA(cudaGetLastError());
  return EDM_OK;
}

// One-shots set by edm_bias_round_commit_on / edm_bias_round_after: from here on the round writes the grid.
static int round_switch_to_commit_stream(edm_bias* b, cudaStream_t& st, bool& moved) {
  moved = false;
  if (b->commit_stream_set) {
    b->commit_stream_set = 0;
    if (b->commit_stream != st) {
      if (!b->ev_round_ready) EDM_CUDA(cudaEventCreateWithFlags(&b->ev_round_ready, cudaEventDisableTiming));
      EDM_CUDA(cudaEventRecord(b->ev_round_ready, st));
      EDM_CUDA(cudaStreamWaitEvent(b->commit_stream, b->ev_round_ready, 0));
      st = b->commit_stream;
      moved = true;
    }
  }
  if (b->round_after) {
    EDM_CUDA(cudaStreamWaitEvent(st, b->round_after, 0));
    b->round_after = nullptr;
  }
  return EDM_OK;
}

template <int DIM>
static int launch_round_dim(edm_bias* b, const RoundParams& rp, const GridDesc& hist, const GridDesc& target,
                            bool fast, cudaStream_t st) {
  HillAccepted* tmp = b->d_accepted + b->accepted_cap;
  const GridDesc& bias = b->bias->d;
  const double* blocks = b->round_blocks;  // candidates arrive as exchange blocks (edm_bias_hills_commit_dev)
  if (blocks && !fast) {                   // the in-order kernel reads the accepted buffer: unpack first
    unpack_blocks_kernel<DIM><<<1, 256, 0, st>>>(b->d_state, b->d_accepted, b->accepted_cap, blocks, b->round_nblocks,
                                                 b->round_block_cap);
    count_launches(1);
  }
  if (fast) {
    const long cap = b->accepted_cap;
    // the plan is sized for the usual few hundred hills; larger rounds take the sequential path
    const long n_max = cap < EDM_ROUND_MAX ? cap : EDM_ROUND_MAX;
    const size_t b_dbl = (size_t)(DIM + 2) * cap * sizeof(double);
    const size_t b_plan = ((size_t)n_max * sizeof(PlanHill<DIM>) + 15) / 16 * 16;
    const size_t b_cells = (size_t)n_max * sizeof(int4);
    // local tempering: (later hill, earlier reaching hill) pairs the plan can list
    const int term_cap = 16384;
    const size_t b_terms = (size_t)term_cap * (1 << DIM) * RecW<DIM>::value * sizeof(double);
    const size_t b_tj = ((size_t)term_cap * sizeof(int) + 15) / 16 * 16;
    const size_t b_nd = ((size_t)n_max * sizeof(int) + 15) / 16 * 16;
    EDM_TRY(b->fast.reserve(b_dbl + b_plan + 2 * b_cells + b_terms + b_tj + b_nd));
    double* centres = b->fast.as<double>();
    double* heights = centres + (size_t)DIM * cap;
    double* ba = heights + cap;
    PlanHill<DIM>* plan = reinterpret_cast<PlanHill<DIM>*>(b->fast.as<char>() + b_dbl);
    int4* cells = reinterpret_cast<int4*>(b->fast.as<char>() + b_dbl + b_plan);
    double* terms = reinterpret_cast<double*>(b->fast.as<char>() + b_dbl + b_plan + b_cells);
    int* term_j = reinterpret_cast<int*>(b->fast.as<char>() + b_dbl + b_plan + b_cells + b_terms);
    int4* cells_folded = reinterpret_cast<int4*>(b->fast.as<char>() + b_dbl + b_plan + b_cells + b_terms + b_tj);
    int* ndep = reinterpret_cast<int*>(b->fast.as<char>() + b_dbl + b_plan + 2 * b_cells + b_terms + b_tj);
    // one cluster of kPlanCtas CTAs (compile-time __cluster_dims__)
    EDM_CUDA(launch_pdl(round_plan_kernel<DIM>, dim3(kPlanCtas), dim3(512), 0, st, bias, target, rp, (int)n_max, b->d_state,
                        b->d_accepted, tmp, centres, heights, plan, cells, cells_folded, ndep, terms, term_j, term_cap,
                        blocks, b->round_nblocks, b->round_block_cap));
    const long nsm2 = 2L * sm_count(b->device);  // two CTAs per SM are resident: more would only queue
    const int nblk = (int)(n_max < nsm2 ? n_max : nsm2);
    EDM_CUDA(launch_pdl(round_integrals_kernel<DIM>, dim3(nblk), dim3(512), 0, st, bias, hist, rp, b->d_state,
                        (const double*)centres, (const double*)heights, ba, b->d_log));
    count_launches(2);
    const bool one_d = DIM == 1 && deposit1d_eligible(b->bias);
    // owner-computes deposit of hills [0, n_fast) in 1-D: staged from the stored values into scratch (reads the
    // grid only), written back below if the round was committed (n_fast = 0 stages the stored values themselves)
    if (one_d) EDM_TRY(deposit1d_stage(b->bias, centres, heights, nullptr, &b->d_state->n_fast, n_max, st));
    // Everything so far only read the grid.  What follows writes it, so whoever else still reads it (this step's
    // force update) must be done first: either the writers move to the stream the force update runs on
    // (commit_stream: they simply follow it there, no cross-stream wake-up on the critical path; the read-only
    // part is linked in by an event that has long fired by then), or this stream waits for the caller's event.
    bool moved = false;
    EDM_TRY(round_switch_to_commit_stream(b, st, moved));
    if (one_d) {
      EDM_TRY(deposit1d_commit_if(b->bias, &b->d_state->round_mode, 2, st));
    } else {
      // one-shot (edm_bias_energy_with_round): the energy sum of the last force update rides in this kernel
      double* e_out = b->energy_with_round;
      b->energy_with_round = nullptr;
      if (moved)  // behind a long force update: launched plainly, or its whole grid would sit resident next to it
        round_deposit_kernel<DIM><<<nblk, 512, 0, st>>>(bias, b->d_state, centres, heights, cells, b->bias->d_flags,
                                                         b->d_energy_partial, b->last_partials, e_out);
      else
        EDM_CUDA(launch_pdl(round_deposit_kernel<DIM>, dim3(nblk), dim3(512), 0, st, bias, b->d_state,
                            (const double*)centres, (const double*)heights, (const int4*)cells, b->bias->d_flags,
                            (const double*)b->d_energy_partial, b->last_partials, e_out));
      count_launches(1);
      if (bias.n_dup) {
        EDM_TRY(edm_grid_dup_boundary_if(b->bias, &b->d_state->round_mode, 2, st));
        count_launches(1);
      }
    }
  } else {
    EDM_CUDA(launch_pdl(reset_mode_kernel, dim3(1), dim3(1), 0, st, b->d_state));
    count_launches(1);
  }
  count_launches(1);
  {  // the in-order kernel writes the grid as well (a no-op if the fast branch above already switched)
    bool moved = false;
    EDM_TRY(round_switch_to_commit_stream(b, st, moved));
  }
  EDM_CUDA(launch_pdl(hill_round_kernel<DIM>, dim3(1), dim3(512), 0, st, bias, hist, target, rp, b->d_state, b->d_accepted,
                      tmp, b->d_log));
  EDM_CUDA(cudaGetLastError());
  if (b->energy_with_round) {  // no deposit kernel took it along (1-D owner-computes path, in-order round)
    double* e_out = b->energy_with_round;
    b->energy_with_round = nullptr;
    EDM_TRY(edm_bias_energy_dev(b, e_out, st));
  }
  return EDM_OK;
}

// launches the hill round over whatever sits in the accepted buffer
int edm_bias_launch_round(edm_bias* b, long long est, cudaStream_t st, bool exchange) {
  // post_add_hill's flush_buffers (lib/edm_bias.cpp:576-577): gather every rank's accepted hills first
  if (exchange && b->comm) return edm_bias_exchange_round(b, est, st);
  RoundParams rp = round_params(b, est);
  GridDesc none;
  memset(&none, 0, sizeof(none));
  const GridDesc& hist = b->hist ? b->hist->d : none;
  const GridDesc& target = b->target ? b->target->d : none;
  const bool local_tempering = b->prm.b_tempering && b->prm.global_tempering < 0;
  static int allow_fast = -1;
  if (allow_fast < 0) allow_fast = getenv("EDM_NO_FAST_ROUND") ? 0 : 1;
  bool fast = allow_fast != 0;
  if (b->prm.dim == 1 && !deposit1d_eligible(b->bias)) fast = false;
  switch (b->prm.dim) {
    case 1: return launch_round_dim<1>(b, rp, hist, target, fast, st);
    case 2: return launch_round_dim<2>(b, rp, hist, target, fast, st);
    default: return launch_round_dim<3>(b, rp, hist, target, fast, st);
  }
}

int edm_bias_check_round(edm_bias* b) {
  BiasDev hdr;
  EDM_CUDA(cudaMemcpy(&hdr, b->d_state, offsetof(BiasDev, overflow), cudaMemcpyDeviceToHost));
  if (hdr.backlog_full) {
    set_error("The bias overflow buffer is full. Too many hills (lib/edm_bias.cpp:503-507)");
    return EDM_ERR_BACKLOG_FULL;
  }
  if (hdr.accepted_overflow) {
    set_error("accepted-hill buffer exhausted");
    return EDM_ERR_CAPACITY;
  }
  if (hdr.exchange_timeout) {
    set_error("hill exchange: a peer's block did not arrive within EDM_B200_PEER_TIMEOUT seconds");
    return EDM_ERR_COMM;
  }
  return EDM_OK;
}

static int select_launch(edm_bias* b, long n, const double* x, long xs, const double* runiform, const int* mask,
                         int apply_mask, long long est, uint64_t seed, uint64_t step, uint64_t first_counter,
                         cudaStream_t st) {
  if (n <= 0) return EDM_OK;
  const edm_bias_params_t& p = b->prm;
  int accept_all = p.hill_density < 0;
  double thresh = accept_all ? 2.0 : p.hill_density / (double)(int)est;  // lib/edm_bias.cpp:543 (est is an int there)
  uint64_t key = uniform_key(seed, step);
  long long blocks = (n + 255) / 256;
  if (blocks > 8LL * sm_count(b->device)) blocks = 8LL * sm_count(b->device);
  count_launches(1);
  switch (p.dim) {
    case 1: EDM_CUDA(launch_pdl(select_kernel<1>, dim3((unsigned)blocks), dim3(256), 0, st, n, x, xs, runiform, mask, apply_mask, thresh, accept_all, key, first_counter, b->d_state, b->d_accepted, b->accepted_cap)); break;
    case 2: EDM_CUDA(launch_pdl(select_kernel<2>, dim3((unsigned)blocks), dim3(256), 0, st, n, x, xs, runiform, mask, apply_mask, thresh, accept_all, key, first_counter, b->d_state, b->d_accepted, b->accepted_cap)); break;
    default: EDM_CUDA(launch_pdl(select_kernel<3>, dim3((unsigned)blocks), dim3(256), 0, st, n, x, xs, runiform, mask, apply_mask, thresh, accept_all, key, first_counter, b->d_state, b->d_accepted, b->accepted_cap)); break;
  }
  EDM_CUDA(cudaGetLastError());
  return EDM_OK;
}

extern "C" {

int edm_bias_create(edm_bias_t** out, edm_grid_t* bias, edm_grid_t* cv_hist, edm_grid_t* target,
                    const edm_bias_params_t* params) {
  EDM_REQUIRE(out && bias && params, "NULL argument");
  EDM_REQUIRE(bias->d.is_gauss && bias->d.dim == params->dim, "bias must be a GaussGrid of params->dim dimensions");
  EDM_TRY(ensure_device(bias->device));
  {  // experiment hook: EDM_L2_FETCH=32|64|128 sets the device's L2 fetch granularity hint (random-gather workloads)
    static const char* l2f = getenv("EDM_L2_FETCH");
    if (l2f) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(l2f));
  }
  edm_bias* b = new edm_bias();
  b->device = bias->device;
  b->prm = *params;
  b->bias = bias;
  b->hist = cv_hist;
  b->target = target;
  *out = b;
  EDM_CUDA(cudaMalloc(&b->d_state, sizeof(BiasDev)));
  EDM_CUDA(cudaMemset(b->d_state, 0, sizeof(BiasDev)));  // T19: the backlog storage starts zero-filled
  b->log_cap = 1 << 16;
  EDM_CUDA(cudaMalloc(&b->d_log, (size_t)b->log_cap * sizeof(edm_hill_event_t)));
  b->n_partial = sm_count(b->device) * 64;
  EDM_CUDA(cudaMalloc(&b->d_energy_partial, (size_t)b->n_partial * sizeof(double)));
  EDM_CUDA(cudaMalloc(&b->d_scalar, 8 * sizeof(double)));
  EDM_TRY(ensure_accepted(b, 4096));
  return EDM_OK;
}

int edm_bias_destroy(edm_bias_t* b) {
  if (!b) return EDM_OK;
  cudaSetDevice(b->device);
  if (b->d_state) cudaFree(b->d_state);
  if (b->d_accepted) cudaFree(b->d_accepted);
  if (b->d_log) cudaFree(b->d_log);
  if (b->d_energy_partial) cudaFree(b->d_energy_partial);
  if (b->d_scalar) cudaFree(b->d_scalar);
  for (int i = 0; i < 3; i++)
    if (b->ev_pair[i]) cudaEventDestroy(b->ev_pair[i]);
  b->io.release();
  b->io2.release();
  b->io3.release();
  b->io4.release();
  b->cells.release();
  b->fast.release();
  b->xchg.release();
  b->list.release();
  b->cand.release();
  if (b->h_pair_flags) cudaFreeHost((void*)b->h_pair_flags);
  if (b->ev_prev) cudaEventDestroy(b->ev_prev);
  if (b->ev_round_ready) cudaEventDestroy(b->ev_round_ready);
  for (int i = 0; i < 5; i++)
    if (b->ev_e2e[i]) cudaEventDestroy(b->ev_e2e[i]);
  if (b->st_main) cudaStreamDestroy(b->st_main);
  if (b->st_copy) cudaStreamDestroy(b->st_copy);
  if (b->ev_f_up) cudaEventDestroy(b->ev_f_up);
  if (b->ev_f_final) cudaEventDestroy(b->ev_f_final);
  if (b->st_side) {
    cudaStreamDestroy(b->st_side);
    cudaEventDestroy(b->ev_fork);
    cudaEventDestroy(b->ev_forces);
    cudaEventDestroy(b->ev_join);
  }
  if (b->st_up) {
    cudaStreamDestroy(b->st_up);
    for (int c = 0; c < edm_bias::kMaxChunks; c++) {
      cudaEventDestroy(b->ev_chunk_up[c]);
      cudaEventDestroy(b->ev_chunk_done[c]);
    }
    cudaFree(b->d_chunk_energy);
  }
  delete b;
  return EDM_OK;
}

int edm_bias_state(edm_bias_t* b, edm_bias_state_t* out) {
  EDM_REQUIRE(b && out, "NULL argument");
  EDM_TRY(ensure_device(b->device));
  BiasDev hdr;
  EDM_CUDA(cudaMemcpy(&hdr, b->d_state, offsetof(BiasDev, overflow), cudaMemcpyDeviceToHost));
  out->cum_bias = hdr.cum_bias;
  out->temp_hill_cum = hdr.temp_hill_cum;
  out->steps = hdr.steps;
  out->hills_added = hdr.hills_added;
  out->skipped = hdr.skip;
  out->backlog_left = (long)hdr.left;
  out->backlog_right = (long)hdr.right;
  out->n_accepted = hdr.n_accepted_last;
  out->log_dropped = hdr.log_dropped;
  return EDM_OK;
}

int edm_bias_set_cum_bias(edm_bias_t* b, double cum_bias) {
  EDM_REQUIRE(b != nullptr, "NULL argument");
  EDM_TRY(ensure_device(b->device));
  EDM_CUDA(cudaMemcpy(&b->d_state->cum_bias, &cum_bias, sizeof(double), cudaMemcpyHostToDevice));
  return EDM_OK;
}

int edm_bias_backlog_get(edm_bias_t* b, long* left, long* right, double* buffer) {
  EDM_REQUIRE(b != nullptr, "NULL argument");
  EDM_TRY(ensure_device(b->device));
  BiasDev hdr;
  EDM_CUDA(cudaMemcpy(&hdr, b->d_state, offsetof(BiasDev, overflow), cudaMemcpyDeviceToHost));
  if (left) *left = (long)hdr.left;
  if (right) *right = (long)hdr.right;
  if (buffer)
    EDM_CUDA(cudaMemcpy(buffer, b->d_state->overflow, EDM_BUFFER_DBLS * sizeof(double), cudaMemcpyDeviceToHost));
  return EDM_OK;
}

int edm_bias_backlog_set(edm_bias_t* b, long left, long right, const double* buffer) {
  EDM_REQUIRE(b && buffer, "NULL argument");
  EDM_REQUIRE(left >= 0 && left <= right && right <= EDM_BUFFER_SLOTS, "bad backlog indices");
  EDM_TRY(ensure_device(b->device));
  long long lr[2] = {left, right};
  EDM_CUDA(cudaMemcpy(&b->d_state->left, lr, sizeof(lr), cudaMemcpyHostToDevice));
  EDM_CUDA(cudaMemcpy(b->d_state->overflow, buffer, EDM_BUFFER_DBLS * sizeof(double), cudaMemcpyHostToDevice));
  return EDM_OK;
}

int edm_bias_log_read(edm_bias_t* b, edm_hill_event_t* out, long cap, long* n) {
  EDM_REQUIRE(b && n, "NULL argument");
  EDM_TRY(ensure_device(b->device));
  BiasDev hdr;
  EDM_CUDA(cudaMemcpy(&hdr, b->d_state, offsetof(BiasDev, overflow), cudaMemcpyDeviceToHost));
  long cnt = hdr.log_n;
  if (out) {
    long take = cnt < cap ? cnt : cap;
    if (take > 0) EDM_CUDA(cudaMemcpy(out, b->d_log, (size_t)take * sizeof(edm_hill_event_t), cudaMemcpyDeviceToHost));
    int zero = 0;
    EDM_CUDA(cudaMemcpy(&b->d_state->log_n, &zero, sizeof(int), cudaMemcpyHostToDevice));
    *n = take;
  } else {
    *n = cnt;
  }
  return EDM_OK;
}

// ------------------------------------------------------------------ update_forces

int edm_bias_update_forces_dev(edm_bias_t* b, long n, const double* x, long xstride, double* f, long fstride,
                               const int* mask, int apply_mask, double* energy, void* stream) {
  EDM_REQUIRE(b && (n == 0 || (x && f)), "NULL argument");
  EDM_REQUIRE(apply_mask < 0 || mask, "apply_mask >= 0 needs a mask");
  EDM_REQUIRE(xstride >= b->prm.dim && fstride >= b->prm.dim, "stride < dim");
  EDM_TRY(ensure_device(b->device));
  cudaStream_t st = (cudaStream_t)stream;
  // short-lived CTAs (a handful of atoms per thread at most): the SM slots they release let the hill round's
  // small kernels, which sit on a higher-priority stream, in while the force update is still running
  long long blocks = (n + 256 * EDM_FORCES_UNROLL - 1) / (256 * EDM_FORCES_UNROLL);
  if (blocks > b->n_partial) blocks = b->n_partial;
  if (blocks < 1) blocks = 1;
  const GridDesc& g = b->bias->d;
  count_launches(1);
  switch (b->prm.dim) {
    case 1: forces_kernel<1><<<(int)blocks, 256, 0, st>>>(g, n, x, xstride, f, fstride, mask, apply_mask, b->d_energy_partial, b->d_state); break;
    case 2: forces_kernel<2><<<(int)blocks, 256, 0, st>>>(g, n, x, xstride, f, fstride, mask, apply_mask, b->d_energy_partial, b->d_state); break;
    default: forces_kernel<3><<<(int)blocks, 256, 0, st>>>(g, n, x, xstride, f, fstride, mask, apply_mask, b->d_energy_partial, b->d_state); break;
  }
  EDM_CUDA(cudaGetLastError());
  if (b->forces_event) {  // one-shot: "the force update has read the grid" (before the energy sum)
    EDM_CUDA(cudaEventRecord(b->forces_event, st));
    b->forces_event = nullptr;
  }
  b->last_partials = (int)blocks;
  if (energy) {
    count_launches(1);
    sum_partials_kernel<<<1, 1024, 0, st>>>((int)blocks, b->d_energy_partial, energy, b->d_state);
    EDM_CUDA(cudaGetLastError());
  }
  return EDM_OK;
}

// The energy of the last edm_bias_update_forces_dev call made with energy = NULL: the per-CTA partials summed in CTA
// order by a one-CTA kernel.  Lets a caller put that kernel BEHIND the hill deposit instead of between the force update
// and the deposit, where it would sit on the step's critical path.
int edm_bias_energy_dev(edm_bias_t* b, double* energy, void* stream) {
  EDM_REQUIRE(b && energy, "NULL argument");
  EDM_REQUIRE(b->last_partials > 0, "no force update on record");
  EDM_TRY(ensure_device(b->device));
  count_launches(1);
  sum_partials_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(b->last_partials, b->d_energy_partial, energy, nullptr);
  EDM_CUDA(cudaGetLastError());
  return EDM_OK;
}

// Host-buffer coordinate step: update_forces and, if asked, the hill round over the same atoms, with
// one upload of the coordinates.  The atoms go through in chunks on three streams so the two PCIe
// directions and the kernels overlap (pinned host buffers make the copies truly asynchronous;
// pageable ones still work, staged by the driver):
//   up:    x0 f0 | x1 f1 | x2 f2 ...
//   main:        | K1(0) select(0) | K1(1) select(1) ...            | hill round
//   down:                          | f0 | f1 ...
static int coords_pipeline(edm_bias* b, long n, const double* x, long xs, double* f, long fs, const int* mask,
                           int apply_mask, int do_hills, const double* runiform, uint64_t seed, uint64_t step,
                           double* energy) {
  if (energy) *energy = 0.0;
  if (!b->st_main) {
    EDM_CUDA(cudaStreamCreateWithFlags(&b->st_main, cudaStreamNonBlocking));
    EDM_CUDA(cudaStreamCreateWithFlags(&b->st_copy, cudaStreamNonBlocking));
    EDM_CUDA(cudaEventCreateWithFlags(&b->ev_f_up, cudaEventDisableTiming));
    EDM_CUDA(cudaEventCreateWithFlags(&b->ev_f_final, cudaEventDisableTiming));
  }
  if (!b->st_up) {
    EDM_CUDA(cudaStreamCreateWithFlags(&b->st_up, cudaStreamNonBlocking));
    for (int c = 0; c < edm_bias::kMaxChunks; c++) {
      EDM_CUDA(cudaEventCreateWithFlags(&b->ev_chunk_up[c], cudaEventDisableTiming));
      EDM_CUDA(cudaEventCreateWithFlags(&b->ev_chunk_done[c], cudaEventDisableTiming));
    }
    EDM_CUDA(cudaMalloc(&b->d_chunk_energy, edm_bias::kMaxChunks * sizeof(double)));
  }
  EDM_CUDA(cudaDeviceSynchronize());  // earlier work of this handle may sit on other streams
  const long long est = edm_job_est(b, n);  // est_hill_count = nlocal, masked or not (lib/edm_bias.cpp:404, T17)
  if (do_hills) {
    EDM_TRY(edm_bias_size_accepted(b, (double)n, est));
    EDM_TRY(edm_bias_reset_accepted(b, b->st_main));
  }
  if (n > 0) {
    const size_t bx = (size_t)n * xs * sizeof(double), bf = (size_t)n * fs * sizeof(double);
    EDM_TRY(b->io.reserve(bx));
    EDM_TRY(b->io2.reserve(bf));
    if (apply_mask >= 0) EDM_TRY(b->io3.reserve((size_t)n * sizeof(int)));
    if (do_hills && runiform) EDM_TRY(b->io4.reserve((size_t)n * sizeof(double)));
    double* dx = b->io.as<double>();
    double* df = b->io2.as<double>();
    int* dm = apply_mask >= 0 ? b->io3.as<int>() : nullptr;
    double* du = (do_hills && runiform) ? b->io4.as<double>() : nullptr;
    long chunk = (n + edm_bias::kMaxChunks - 1) / edm_bias::kMaxChunks;
    if (chunk < (1L << 18)) chunk = 1L << 18;  // enough atoms per launch to fill the GPU
    const int nchunks = (int)((n + chunk - 1) / chunk);
    for (int c = 0; c < nchunks; c++) {
      const long o = (long)c * chunk, cnt = (n - o < chunk) ? n - o : chunk;
      EDM_CUDA(cudaMemcpyAsync(dx + o * xs, x + o * xs, (size_t)cnt * xs * sizeof(double), cudaMemcpyHostToDevice, b->st_up));
      EDM_CUDA(cudaMemcpyAsync(df + o * fs, f + o * fs, (size_t)cnt * fs * sizeof(double), cudaMemcpyHostToDevice, b->st_up));
      if (dm) EDM_CUDA(cudaMemcpyAsync(dm + o, mask + o, (size_t)cnt * sizeof(int), cudaMemcpyHostToDevice, b->st_up));
      if (du) EDM_CUDA(cudaMemcpyAsync(du + o, runiform + o, (size_t)cnt * sizeof(double), cudaMemcpyHostToDevice, b->st_up));
      EDM_CUDA(cudaEventRecord(b->ev_chunk_up[c], b->st_up));
      EDM_CUDA(cudaStreamWaitEvent(b->st_main, b->ev_chunk_up[c], 0));
      EDM_TRY(edm_bias_update_forces_dev(b, cnt, dx + o * xs, xs, df + o * fs, fs, dm ? dm + o : nullptr, apply_mask,
                                         b->d_chunk_energy + c, b->st_main));
      // candidate keys = atom indices
      if (do_hills)
        EDM_TRY(select_launch(b, cnt, dx + o * xs, xs, du ? du + o : nullptr, dm ? dm + o : nullptr, apply_mask, est, seed,
                              step, (uint64_t)o, b->st_main));
      EDM_CUDA(cudaEventRecord(b->ev_chunk_done[c], b->st_main));
      EDM_CUDA(cudaStreamWaitEvent(b->st_copy, b->ev_chunk_done[c], 0));
      EDM_CUDA(cudaMemcpyAsync(f + o * fs, df + o * fs, (size_t)cnt * fs * sizeof(double), cudaMemcpyDeviceToHost, b->st_copy));
    }
    if (do_hills) EDM_TRY(edm_bias_launch_round(b, est, b->st_main));
    double e[edm_bias::kMaxChunks];
    EDM_CUDA(cudaMemcpyAsync(e, b->d_chunk_energy, nchunks * sizeof(double), cudaMemcpyDeviceToHost, b->st_main));
    EDM_CUDA(cudaStreamSynchronize(b->st_main));
    double tot = 0.0;
    for (int c = 0; c < nchunks; c++) tot += e[c];  // chunk order: run-to-run deterministic
    if (energy) *energy = tot;
    EDM_CUDA(cudaStreamSynchronize(b->st_copy));
  } else if (do_hills) {
    EDM_TRY(edm_bias_launch_round(b, est, b->st_main));
    EDM_CUDA(cudaStreamSynchronize(b->st_main));
  }
  return do_hills ? edm_bias_check_round(b) : EDM_OK;
}

int edm_bias_update_forces(edm_bias_t* b, long n, const double* x, long xstride, double* f, long fstride,
                           const int* mask, int apply_mask, double* energy) {
  EDM_REQUIRE(b && (n <= 0 || (x && f)), "NULL argument");
  EDM_REQUIRE(apply_mask < 0 || mask || n <= 0, "apply_mask >= 0 needs a mask");
  EDM_REQUIRE(n <= 0 || (xstride >= b->prm.dim && fstride >= b->prm.dim), "stride < dim");
  EDM_TRY(ensure_device(b->device));
  if (energy) *energy = 0.0;
  if (n <= 0) return EDM_OK;
  return coords_pipeline(b, n, x, xstride, f, fstride, mask, apply_mask, 0, nullptr, 0, 0, energy);
}

int edm_bias_step_coords(edm_bias_t* b, long n, const double* x, long xstride, double* f, long fstride,
                         const int* mask, int apply_mask, int do_hills, const double* runiform, uint64_t seed,
                         uint64_t step, double* energy) {
  EDM_REQUIRE(b && (n <= 0 || (x && f)), "NULL argument");
  EDM_REQUIRE(apply_mask < 0 || mask || n <= 0, "apply_mask >= 0 needs a mask");
  EDM_REQUIRE(n <= 0 || (xstride >= b->prm.dim && fstride >= b->prm.dim), "stride < dim");
  EDM_TRY(ensure_device(b->device));
  return coords_pipeline(b, n < 0 ? 0 : n, x, xstride, f, fstride, mask, apply_mask, do_hills, runiform, seed, step,
                         energy);
}

// ------------------------------------------------------------------ add_hills

int edm_bias_select_dev(edm_bias_t* b, long n, const double* x, long xstride, const double* runiform, const int* mask,
                        int apply_mask, long long est_hill_count, uint64_t seed, uint64_t step, uint64_t first_counter,
                        void* stream) {
  EDM_REQUIRE(b && (n == 0 || x), "NULL argument");
  EDM_TRY(ensure_device(b->device));
  EDM_TRY(edm_bias_size_accepted(b, (double)first_counter + (double)n, est_hill_count));
  return select_launch(b, n, x, xstride, runiform, mask, apply_mask, est_hill_count, seed, step, first_counter,
                       (cudaStream_t)stream);
}

int edm_bias_add_hills_dev(edm_bias_t* b, long n, const double* x, long xstride, const double* runiform,
                           const int* mask, int apply_mask, uint64_t seed, uint64_t step, void* stream) {
  EDM_REQUIRE(b != nullptr, "NULL argument");
  EDM_TRY(ensure_device(b->device));
  cudaStream_t st = (cudaStream_t)stream;
  const long long est = edm_job_est(b, n);  // est_hill_count = nlocal, masked or not (lib/edm_bias.cpp:404, T17)
  EDM_TRY(edm_bias_size_accepted(b, (double)n, est));
  EDM_TRY(edm_bias_reset_accepted(b, st));
  EDM_TRY(select_launch(b, n, x, xstride, runiform, mask, apply_mask, est, seed, step, 0, st));
  return edm_bias_launch_round(b, est, st);
}

int edm_bias_energy_with_round(edm_bias_t* b, double* energy) {
  EDM_REQUIRE(b != nullptr, "NULL argument");
  b->energy_with_round = energy;
  return EDM_OK;
}

int edm_bias_round_commit_on(edm_bias_t* b, void* stream) {
  EDM_REQUIRE(b != nullptr, "NULL argument");
  b->commit_stream = (cudaStream_t)stream;
  b->commit_stream_set = 1;
  return EDM_OK;
}

int edm_bias_round_after(edm_bias_t* b, void* event) {
  EDM_REQUIRE(b != nullptr, "NULL argument");
  b->round_after = (cudaEvent_t)event;
  return EDM_OK;
}

// update_forces and the hill round of one step on device buffers.  The round's selection, plan,
// integrals and decision only read the grid, so they run on a side stream next to the force update;
// the deposit waits for the force update (it reads the start-of-step bias), and the caller's stream
// waits for the round.
int edm_bias_step_coords_dev(edm_bias_t* b, long n, const double* x, long xstride, double* f, long fstride,
                             const int* mask, int apply_mask, int do_hills, const double* runiform, uint64_t seed,
                             uint64_t step, double* energy, void* stream) {
  EDM_REQUIRE(b != nullptr, "NULL argument");
  if (!do_hills) return edm_bias_update_forces_dev(b, n, x, xstride, f, fstride, mask, apply_mask, energy, stream);
  EDM_TRY(ensure_device(b->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (!b->st_side) {
    // highest priority: the round's kernels are tiny next to the force update and must not queue behind its CTAs
    int prio_lo = 0, prio_hi = 0;
    EDM_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    EDM_CUDA(cudaStreamCreateWithPriority(&b->st_side, cudaStreamNonBlocking, prio_hi));
    EDM_CUDA(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
    EDM_CUDA(cudaEventCreateWithFlags(&b->ev_forces, cudaEventDisableTiming));
    EDM_CUDA(cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming));
  }
  EDM_CUDA(cudaEventRecord(b->ev_fork, st));
  EDM_CUDA(cudaStreamWaitEvent(b->st_side, b->ev_fork, 0));
  // The selection goes in FIRST: it is a few microseconds of hashing, but launched behind the force update its CTAs
  // wait ~20 us for SM slots among the force CTAs and everything behind it (exchange, plan, integrals, decision)
  // starts that much later -- on several GPUs, where the round is as long as the force update, that is step time.
  const long long est = edm_job_est(b, n);
  EDM_TRY(edm_bias_size_accepted(b, (double)n, est));
  EDM_TRY(edm_bias_reset_accepted(b, b->st_side));
  EDM_TRY(select_launch(b, n, x, xstride, runiform, mask, apply_mask, est, seed, step, 0, b->st_side));
  b->forces_event = b->ev_forces;
  // the energy sum (a one-CTA kernel) goes behind the round's deposit, not between it and the force update
  if (n > 0) EDM_TRY(edm_bias_update_forces_dev(b, n, x, xstride, f, fstride, mask, apply_mask, nullptr, stream));
  if (b->forces_event) {  // nothing was launched
    EDM_CUDA(cudaEventRecord(b->ev_forces, st));
    b->forces_event = nullptr;
  }
  // the round's writers (deposit, in-order tail) follow the force update on the caller's stream; nothing to join
  b->commit_stream = st;
  b->commit_stream_set = 1;
  if (energy && n > 0) b->energy_with_round = energy;  // summed inside the deposit kernel (or right after the round)
  EDM_TRY(edm_bias_launch_round(b, est, b->st_side));
  return EDM_OK;
}

int edm_bias_add_hills(edm_bias_t* b, long n, const double* x, long xstride, const double* runiform, const int* mask,
                       int apply_mask, uint64_t seed, uint64_t step) {
  EDM_REQUIRE(b && (n == 0 || x), "NULL argument");
  EDM_TRY(ensure_device(b->device));
  const double* dx = nullptr;
  const double* du = nullptr;
  const int* dm = nullptr;
  if (n > 0) {
    size_t bx = (size_t)n * xstride * sizeof(double);
    EDM_TRY(b->io.reserve(bx));
    EDM_CUDA(cudaMemcpyAsync(b->io.p, x, bx, cudaMemcpyHostToDevice, 0));
    dx = b->io.as<double>();
    if (runiform) {
      EDM_TRY(b->io2.reserve((size_t)n * sizeof(double)));
      EDM_CUDA(cudaMemcpyAsync(b->io2.p, runiform, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, 0));
      du = b->io2.as<double>();
    }
    if (apply_mask >= 0) {
      EDM_REQUIRE(mask != nullptr, "apply_mask >= 0 needs a mask");
      EDM_TRY(b->io3.reserve((size_t)n * sizeof(int)));
      EDM_CUDA(cudaMemcpyAsync(b->io3.p, mask, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, 0));
      dm = b->io3.as<int>();
    }
  }
  EDM_TRY(edm_bias_add_hills_dev(b, n, dx, xstride, du, dm, apply_mask, seed, step, nullptr));
  EDM_CUDA(cudaDeviceSynchronize());
  return edm_bias_check_round(b);
}

int edm_bias_pre_add_hill(edm_bias_t* b, int est_hill_count) {
  EDM_REQUIRE(b != nullptr, "NULL argument");
  EDM_TRY(ensure_device(b->device));
  EDM_TRY(edm_bias_reset_accepted(b, 0));
  b->in_round = 1;
  b->round_est = edm_job_est(b, est_hill_count);
  b->round_count = 0;
  return EDM_OK;
}

int edm_bias_add_hill_batch(edm_bias_t* b, long n, const double* x, const double* runiform) {
  EDM_REQUIRE(b && (n == 0 || x), "NULL argument");
  if (!b->in_round) {  // lib/edm_bias.cpp:530-531 aborts
    set_error("Must call pre_add_hill before add_hill");
    return EDM_ERR_STATE;
  }
  EDM_REQUIRE(runiform || b->prm.hill_density < 0, "add_hill needs runiform when hill_density > 0");
  EDM_TRY(ensure_device(b->device));
  if (n <= 0) return EDM_OK;
  int D = b->prm.dim;
  size_t bx = (size_t)n * D * sizeof(double);
  EDM_TRY(b->io.reserve(bx));
  EDM_CUDA(cudaMemcpyAsync(b->io.p, x, bx, cudaMemcpyHostToDevice, 0));
  const double* du = nullptr;
  if (runiform) {
    EDM_TRY(b->io2.reserve((size_t)n * sizeof(double)));
    EDM_CUDA(cudaMemcpyAsync(b->io2.p, runiform, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, 0));
    du = b->io2.as<double>();
  }
  EDM_TRY(edm_bias_size_accepted(b, (double)b->round_count + (double)n, b->round_est));
  EDM_TRY(select_launch(b, n, b->io.as<double>(), D, du, nullptr, -1, b->round_est, 0, 0, b->round_count, 0));
  b->round_count += (unsigned long long)n;
  EDM_CUDA(cudaDeviceSynchronize());
  return EDM_OK;
}

int edm_bias_post_add_hill(edm_bias_t* b) {
  EDM_REQUIRE(b != nullptr, "NULL argument");
  if (!b->in_round) {
    set_error("Must call pre_add_hill before post_add_hill");
    return EDM_ERR_STATE;
  }
  EDM_TRY(ensure_device(b->device));
  b->in_round = 0;
  EDM_TRY(edm_bias_launch_round(b, b->round_est, 0));
  EDM_CUDA(cudaDeviceSynchronize());
  return edm_bias_check_round(b);
}

// %globaltimer stamps of the last hill round in microseconds relative to the plan's first instruction:
// [0..6] plan phases (start, mode known, order known, entries planned, counts, lists, serial part done),
// [7,8] decision begin/end, [9,10] first deposit taken / last deposit done, [11,12] in-order kernel begin/end.
int edm_bias_round_times_us(edm_bias_t* b, double* out13) {
  EDM_REQUIRE(b && out13, "NULL argument");
  EDM_TRY(ensure_device(b->device));
  BiasDev hdr;
  EDM_CUDA(cudaMemcpy(&hdr, b->d_state, offsetof(BiasDev, overflow), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 16; i++) out13[i] = ((double)hdr.stamp[i] - (double)hdr.stamp[0]) * 1e-3;
  return EDM_OK;
}

// Selection and exchange of the last round, us relative to the plan's begin (the same origin as
// edm_bias_round_times_us): [0] selection began, [1] its last CTA left, [2] exchange kernel past its predecessor,
// [3] own block delivered to every peer, [4] every peer's block in.  out holds 5 doubles.
int edm_bias_exchange_times_us(edm_bias_t* b, double* out5) {
  EDM_REQUIRE(b && out5, "NULL argument");
  EDM_TRY(ensure_device(b->device));
  BiasDev hdr;
  EDM_CUDA(cudaMemcpy(&hdr, b->d_state, offsetof(BiasDev, overflow), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 5; i++) out5[i] = ((double)hdr.stamp2[i] - (double)hdr.stamp[0]) * 1e-3;
  return EDM_OK;
}

int edm_bias_round_info(edm_bias_t* b, long long* parallel, long long* split, long long* in_order) {
  EDM_REQUIRE(b != nullptr, "NULL argument");
  EDM_TRY(ensure_device(b->device));
  BiasDev hdr;
  EDM_CUDA(cudaMemcpy(&hdr, b->d_state, offsetof(BiasDev, overflow), cudaMemcpyDeviceToHost));
  if (parallel) *parallel = hdr.rounds_parallel;
  if (split) *split = hdr.rounds_split;
  if (in_order) *in_order = hdr.rounds_in_order;
  return EDM_OK;
}

int edm_bias_set_profiling(edm_bias_t* b, int on) {
  EDM_REQUIRE(b != nullptr, "NULL argument");
  EDM_TRY(ensure_device(b->device));
  if (on && !b->ev_pair[0]) {
    EDM_CUDA(cudaEventCreate(&b->ev_pair[0]));
    EDM_CUDA(cudaEventCreate(&b->ev_pair[1]));
    EDM_CUDA(cudaEventCreate(&b->ev_pair[2]));
  }
  b->profiling = on;
  return EDM_OK;
}

int edm_bias_profile_ms(edm_bias_t* b, double* pair_kernel_ms) {
  EDM_REQUIRE(b && pair_kernel_ms && b->ev_pair[0], "profiling was never enabled");
  EDM_TRY(ensure_device(b->device));
  EDM_CUDA(cudaEventSynchronize(b->ev_pair[1]));
  float ms = 0;
  EDM_CUDA(cudaEventElapsedTime(&ms, b->ev_pair[0], b->ev_pair[1]));
  *pair_kernel_ms = ms;
  return EDM_OK;
}

int edm_bias_profile_pair_ms(edm_bias_t* b, double* search_ms, double* eval_ms) {
  EDM_REQUIRE(b && search_ms && eval_ms && b->ev_pair[0], "profiling was never enabled");
  EDM_TRY(ensure_device(b->device));
  EDM_CUDA(cudaEventSynchronize(b->ev_pair[1]));
  float a = 0, c = 0;
  // ev_pair[2] sits between the block search and the block evaluation; it is not recorded when the
  // generic search ran instead
  if (cudaEventElapsedTime(&a, b->ev_pair[0], b->ev_pair[2]) != cudaSuccess ||
      cudaEventElapsedTime(&c, b->ev_pair[2], b->ev_pair[1]) != cudaSuccess) {
    cudaGetLastError();
    a = c = 0;
  }
  *search_ms = a;
  *eval_ms = c;
  return EDM_OK;
}

// Where the last profiled edm_pair_step_cells (host buffers) spent its time on the device: upload of the
// positions, everything from there to the step's last kernel (binning, pair kernels, hill round, report),
// the download of the forces (which overlaps the hill round), and the whole span from the first copy to
// the later of the two ends.  Synchronises on the recorded events.
int edm_bias_profile_e2e_ms(edm_bias_t* b, double* x_up_ms, double* kernels_ms, double* f_down_ms, double* span_ms) {
  EDM_REQUIRE(b && b->e2e_valid, "no profiled host-buffer pair step (edm_bias_set_profiling, then edm_pair_step_cells)");
  EDM_TRY(ensure_device(b->device));
  EDM_CUDA(cudaEventSynchronize(b->ev_e2e[2]));
  EDM_CUDA(cudaEventSynchronize(b->ev_e2e[4]));
  float up = 0, k = 0, dn = 0, s1 = 0, s2 = 0;
  EDM_CUDA(cudaEventElapsedTime(&up, b->ev_e2e[0], b->ev_e2e[1]));
  EDM_CUDA(cudaEventElapsedTime(&k, b->ev_e2e[1], b->ev_e2e[2]));
  EDM_CUDA(cudaEventElapsedTime(&dn, b->ev_e2e[3], b->ev_e2e[4]));
  EDM_CUDA(cudaEventElapsedTime(&s1, b->ev_e2e[0], b->ev_e2e[2]));
  EDM_CUDA(cudaEventElapsedTime(&s2, b->ev_e2e[0], b->ev_e2e[4]));
  if (x_up_ms) *x_up_ms = up;
  if (kernels_ms) *kernels_ms = k;
  if (f_down_ms) *f_down_ms = dn;
  if (span_ms) *span_ms = s1 > s2 ? s1 : s2;
  return EDM_OK;
}

// ------------------------------------------------------------------ multi-GPU exchange

size_t edm_hill_block_doubles(int dim, long cap) { return 1 + (size_t)cap * dim; }

int edm_bias_hills_pack_dev(edm_bias_t* b, double* block, long cap, void* stream) {
  EDM_REQUIRE(b && block && cap > 0, "bad argument");
  EDM_TRY(ensure_device(b->device));
  cudaStream_t st = (cudaStream_t)stream;
  HillAccepted* tmp = b->d_accepted + b->accepted_cap;
  count_launches(1);
  switch (b->prm.dim) {
    case 1: EDM_CUDA(launch_pdl(pack_block_kernel<1>, dim3(1), dim3(512), 0, st, b->d_state, b->d_accepted, tmp, block, cap)); break;
    case 2: EDM_CUDA(launch_pdl(pack_block_kernel<2>, dim3(1), dim3(512), 0, st, b->d_state, b->d_accepted, tmp, block, cap)); break;
    default: EDM_CUDA(launch_pdl(pack_block_kernel<3>, dim3(1), dim3(512), 0, st, b->d_state, b->d_accepted, tmp, block, cap)); break;
  }
  EDM_CUDA(cudaGetLastError());
  return EDM_OK;
}

}  // extern "C"

// Internal (edm_comm.cu): the peer-window exchange of this rank's accepted hills; *blocks_out = the rank-major
// concatenation in the local window once the kernel has finished.
int edm_bias_hills_push_dev(edm_bias* b, edm_comm* c, long cap, cudaStream_t st, const double** blocks_out) {
  const int parity = c->epoch & 1;
  const int epoch = ++c->epoch;
  HillAccepted* tmp = b->d_accepted + b->accepted_cap;
  static const double timeout_s = getenv("EDM_B200_PEER_TIMEOUT") ? atof(getenv("EDM_B200_PEER_TIMEOUT")) : 10.0;
  const unsigned long long timeout_ns = (unsigned long long)(timeout_s * 1e9);
  count_launches(1);
  switch (b->prm.dim) {
    case 1: EDM_CUDA(launch_pdl(pack_push_kernel<1>, dim3(1), dim3(512), 0, st, b->d_state, b->d_accepted, tmp, cap, (char* const*)c->d_peer, c->nranks, c->rank, parity, epoch, timeout_ns)); break;
    case 2: EDM_CUDA(launch_pdl(pack_push_kernel<2>, dim3(1), dim3(512), 0, st, b->d_state, b->d_accepted, tmp, cap, (char* const*)c->d_peer, c->nranks, c->rank, parity, epoch, timeout_ns)); break;
    default: EDM_CUDA(launch_pdl(pack_push_kernel<3>, dim3(1), dim3(512), 0, st, b->d_state, b->d_accepted, tmp, cap, (char* const*)c->d_peer, c->nranks, c->rank, parity, epoch, timeout_ns)); break;
  }
  EDM_CUDA(cudaGetLastError());
  const size_t bw = edm_hill_block_doubles(b->prm.dim, cap);
  *blocks_out = reinterpret_cast<const double*>(c->win + EDM_PEER_FLAG_BYTES) + (size_t)parity * c->nranks * bw;
  return EDM_OK;
}

extern "C" {

int edm_bias_hills_commit_dev(edm_bias_t* b, const double* blocks, int nblocks, long cap, long long est_hill_count,
                              void* stream) {
  EDM_REQUIRE(b && blocks && nblocks > 0 && cap > 0, "bad argument");
  EDM_TRY(ensure_device(b->device));
  EDM_REQUIRE(nblocks <= 64, "at most 64 exchange blocks");
  cudaStream_t st = (cudaStream_t)stream;
  EDM_TRY(ensure_accepted(b, (long)nblocks * cap));
  // the round's plan reads the blocks itself (its first phase is the unpack)
  b->round_blocks = blocks;
  b->round_nblocks = nblocks;
  b->round_block_cap = cap;
  const int rc = edm_bias_launch_round(b, est_hill_count, st, false);
  b->round_blocks = nullptr;
  return rc;
}

int edm_bias_check(edm_bias_t* b) {
  EDM_REQUIRE(b != nullptr, "NULL argument");
  EDM_TRY(ensure_device(b->device));
  return edm_bias_check_round(b);
}

}  // extern "C"
