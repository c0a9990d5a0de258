// Pair-distance collective variable of fix edm_pair (lammps/fix_edm_pair.cpp:139-256) on the
// device: cell binning, half-shell pair search, 1-D bias evaluation at r, force scatter and the
// two-per-pair hill proposals.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "edm_host.h"

int edm_bias_reset_accepted(edm_bias* b, cudaStream_t st);
int edm_bias_launch_round(edm_bias* b, long long est, cudaStream_t st);
int edm_bias_check_round(edm_bias* b);

namespace edm {

struct CellGrid {
  int nc[3];
  double cs[3];
  double box[3];
  int ncell;
};

struct PairParams {
  int itype, jtype, use_types;
  int do_hills, accept_all;
  double thresh;
  uint64_t thresh_bits;  // ceil(thresh * 2^32), saturated
  int dbg;
  uint64_t key;
  double rc2;
  long natoms;
  long acc_cap;
};

__device__ __forceinline__ int cell_coord(double x, double cs, int nc) {
  int c = (int)floor(x / cs);
  return c < 0 ? 0 : (c >= nc ? nc - 1 : c);
}

__global__ void cell_count_kernel(long n, const double* __restrict__ x, CellGrid cg, int* __restrict__ cell_of,
                                  int* __restrict__ count) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cx = cell_coord(x[3 * i + 0], cg.cs[0], cg.nc[0]);
  int cy = cell_coord(x[3 * i + 1], cg.cs[1], cg.nc[1]);
  int cz = cell_coord(x[3 * i + 2], cg.cs[2], cg.nc[2]);
  int c = (cz * cg.nc[1] + cy) * cg.nc[0] + cx;
  cell_of[i] = c;
  atomicAdd(&count[c], 1);
}

// exclusive scan of count[0..n) into start[0..n], two passes over tiles of 2048 cells:
// (1) per-tile totals, (2) every tile sums the totals before it and scans its own cells
constexpr int kScanTile = 2048;  // 256 threads x 8 cells

__global__ void __launch_bounds__(256) cell_scan_totals_kernel(int n, const int* __restrict__ count,
                                                               int* __restrict__ tile_total) {
  __shared__ int wsum[8];
  const int base = blockIdx.x * kScanTile + threadIdx.x * 8;
  int s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++)
    if (base + k < n) s += count[base + k];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < 8; i++) t += wsum[i];
    tile_total[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) cell_scan_kernel(int n, const int* __restrict__ count,
                                                        const int* __restrict__ tile_total, int* __restrict__ start) {
  __shared__ int wsum[8];
  __shared__ int s_offset;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // offset of this tile: sum of the totals of the tiles before it
  int part = 0;
  for (int t = threadIdx.x; t < (int)blockIdx.x; t += blockDim.x) part += tile_total[t];
  for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
  if (lane == 0) wsum[wid] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < 8; i++) t += wsum[i];
    s_offset = t;
  }
  __syncthreads();
  const int base = blockIdx.x * kScanTile + threadIdx.x * 8;
  int v[8], s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    v[k] = (base + k < n) ? count[base + k] : 0;
    s += v[k];
  }
  int incl = s;  // inclusive scan of the per-thread sums across the warp
  for (int o = 1; o < 32; o <<= 1) {
    int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  __syncthreads();
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  int woff = 0;
  for (int i = 0; i < wid; i++) woff += wsum[i];
  int run = s_offset + woff + incl - s;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    if (base + k < n) start[base + k] = run;
    run += v[k];
  }
  if (base <= n - 1 && n - 1 < base + 8) start[n] = run;  // the thread that owns the last cell closes the array
}

__global__ void cell_fill_kernel(long n, const int* __restrict__ cell_of, const int* __restrict__ start,
                                 int* __restrict__ fill, int* __restrict__ order) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = cell_of[i];
  int slot = start[c] + atomicAdd(&fill[c], 1);
  order[slot] = (int)i;
}

// canonical order inside a cell (ascending atom index) + gather of positions/types into slot order,
// so the result does not depend on the order the atomics above happened to resolve in
__global__ void cell_sort_gather_kernel(CellGrid cg, const int* __restrict__ start, int* __restrict__ order,
                                        const double* __restrict__ x, const int* __restrict__ type,
                                        double* __restrict__ xs, float* __restrict__ xs32, int* __restrict__ ts) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cg.ncell) return;
  int lo = start[c], hi = start[c + 1];
  for (int a = lo + 1; a < hi; a++) {
    int v = order[a];
    int b = a - 1;
    while (b >= lo && order[b] > v) {
      order[b + 1] = order[b];
      b--;
    }
    order[b + 1] = v;
  }
  // cell-relative fp32 copy for the pair kernel's prefilter (error ~cs * 6e-8, box-size independent)
  const double cx = (c % cg.nc[0]) * cg.cs[0], cy = ((c / cg.nc[0]) % cg.nc[1]) * cg.cs[1];
  const double cz = (c / (cg.nc[0] * cg.nc[1])) * cg.cs[2];
  for (int a = lo; a < hi; a++) {
    int i = order[a];
    const double px = x[3 * (long)i + 0], py = x[3 * (long)i + 1], pz = x[3 * (long)i + 2];
    xs[3 * (long)a + 0] = px;
    xs[3 * (long)a + 1] = py;
    xs[3 * (long)a + 2] = pz;
    xs32[3 * (long)a + 0] = (float)(px - cx);
    xs32[3 * (long)a + 1] = (float)(py - cy);
    xs32[3 * (long)a + 2] = (float)(pz - cz);
    if (ts) ts[a] = type ? type[i] : 0;
  }
}

// 1-D bias evaluation at r with update_force's sign (lib/edm_bias.cpp:297-311): returns V(r),
// force = -dV/dr.
__device__ __forceinline__ double pair_eval(const GridDesc& g, double r, double& force) {
  double der[1];
  double rr[1] = {r};
  double v = d_eval_point<1>(g, rr, der, g.b_interp != 0);
  force = -der[0];
  return v;
}

// The two hill proposals of a pair (lammps/fix_edm_pair.cpp:230-236) draw their uniforms from one
// hash: u_which = 32-bit half of pair_bits() * 2^-32 (edm_uniform_pair).  u < thresh is decided on
// the integers: bits32 < ceil(thresh * 2^32).
__device__ __forceinline__ void propose_hills(const PairParams& pp, unsigned long long pairkey, double r, BiasDev* st,
                                              HillAccepted* acc) {
  const uint64_t bits = pair_bits(pp.key, pairkey);
#pragma unroll
  for (int which = 0; which < 2; which++) {
    const uint64_t half = which == 0 ? (bits >> 32) : (bits & 0xffffffffULL);
    if (pp.accept_all || half < pp.thresh_bits) {
      int slot = atomicAdd(&st->n_accepted, 1);
      if (slot < pp.acc_cap) {
        acc[slot].key = 2ULL * pairkey + which;
        acc[slot].x[0] = r;
        acc[slot].x[1] = 0.0;
        acc[slot].x[2] = 0.0;
      } else {
        st->accepted_overflow = 1;
      }
    }
  }
}

// v1: one thread per atom (slot order), half shell of 13 forward cells + own cell.
__global__ void __launch_bounds__(128) pair_cells_kernel(GridDesc g, CellGrid cg, PairParams pp,
                                                         const int* __restrict__ start, const int* __restrict__ order,
                                                         const double* __restrict__ xs, const int* __restrict__ ts,
                                                         double* __restrict__ f, double* __restrict__ partial,
                                                         BiasDev* st, HillAccepted* acc) {
  __shared__ double red[33];
  double e = 0.0;
  unsigned long long npairs = 0;
  long stride = (long)gridDim.x * blockDim.x;
  for (long a = (long)blockIdx.x * blockDim.x + threadIdx.x; a < pp.natoms; a += stride) {
    const double xi = xs[3 * a + 0], yi = xs[3 * a + 1], zi = xs[3 * a + 2];
    const int ti = pp.use_types ? ts[a] : 0;
    if (pp.use_types && ti != pp.itype && ti != pp.jtype) continue;
    const int oi = order[a];
    int cx = cell_coord(xi, cg.cs[0], cg.nc[0]);
    int cy = cell_coord(yi, cg.cs[1], cg.nc[1]);
    int cz = cell_coord(zi, cg.cs[2], cg.nc[2]);
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int nb = 0; nb < 14; nb++) {
      // nb 0 = own cell; 1..13 = offsets with (dz,dy,dx) lexicographically positive
      int t = nb + 13;  // 13..26 in the 3x3x3 enumeration, 13 = centre
      int ox = t % 3 - 1, oy = (t / 3) % 3 - 1, oz = t / 9 - 1;
      int qx = cx + ox, qy = cy + oy, qz = cz + oz;
      double sx = 0.0, sy = 0.0, sz = 0.0;
      if (qx >= cg.nc[0]) { qx -= cg.nc[0]; sx = cg.box[0]; } else if (qx < 0) { qx += cg.nc[0]; sx = -cg.box[0]; }
      if (qy >= cg.nc[1]) { qy -= cg.nc[1]; sy = cg.box[1]; } else if (qy < 0) { qy += cg.nc[1]; sy = -cg.box[1]; }
      if (qz >= cg.nc[2]) { qz -= cg.nc[2]; sz = cg.box[2]; } else if (qz < 0) { qz += cg.nc[2]; sz = -cg.box[2]; }
      int q = (qz * cg.nc[1] + qy) * cg.nc[0] + qx;
      long jlo = start[q], jhi = start[q + 1];
      if (nb == 0) jlo = a + 1;
      for (long j = jlo; j < jhi; j++) {
        // separation exactly as the oracle forms it: (x_i - x_j) - image shift, squares summed in x,y,z order
        double dx = __dsub_rn(__dsub_rn(xi, xs[3 * j + 0]), sx);
        double dy = __dsub_rn(__dsub_rn(yi, xs[3 * j + 1]), sy);
        double dz = __dsub_rn(__dsub_rn(zi, xs[3 * j + 2]), sz);
        double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (!(d2 < pp.rc2)) continue;
        if (pp.use_types) {
          int tj = ts[j];
          bool match = (ti == pp.itype) ? (tj == pp.jtype) : (tj == pp.itype);
          if (!match) continue;
        }
        const int oj = order[j];
        // the pair is oriented (i = lower atom index) as in a half list sorted by (i, j)
        double sgn = (oi < oj) ? 1.0 : -1.0;
        double r = sqrt(d2);
        double rinv = 1.0 / r;
        double force;
        e += pair_eval(g, r, force);
        npairs++;
        double px = dx * rinv * force, py = dy * rinv * force, pz = dz * rinv * force;
        fx += px;
        fy += py;
        fz += pz;
        atomicAdd(&f[3 * (long)oj + 0], -px);
        atomicAdd(&f[3 * (long)oj + 1], -py);
        atomicAdd(&f[3 * (long)oj + 2], -pz);
        (void)sgn;
        if (pp.do_hills) {
          unsigned long long lo = oi < oj ? oi : oj, hi = oi < oj ? oj : oi;
          propose_hills(pp, lo * (unsigned long long)pp.natoms + hi, r, st, acc);
        }
      }
    }
    atomicAdd(&f[3 * (long)oi + 0], fx);
    atomicAdd(&f[3 * (long)oi + 1], fy);
    atomicAdd(&f[3 * (long)oi + 2], fz);
  }
  double tot = block_sum(e, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
  // pair count: integer, so the atomic order is irrelevant
  for (int o = 16; o > 0; o >>= 1) npairs += __shfl_down_sync(0xffffffffu, npairs, o);
  if ((threadIdx.x & 31) == 0 && npairs) atomicAdd(&st->n_pairs, npairs);
}

// ---- v4: warp per home cell, fp32 prefilter, warp-level compaction ---------------------------
//
// Only 1 in 6.5 tested pairs lies inside the cutoff, so a thread-per-atom loop (v1 above) runs the
// heavy per-pair path (square root, interpolation, hashing, force scatter) with ~15 % of its lanes
// active.  Here a warp owns one home cell:
//   * STAGE: the atoms of the own cell and of the 13 forward neighbour cells are copied, x-row by
//     x-row (cells adjacent in x are contiguous in slot order), into shared memory as fp32
//     coordinates relative to the corner of the 3x3x3 neighbourhood (built from the cell-relative
//     fp32 copy the binning kernel leaves, so no image shift and no dependence on the box size);
//   * TEST: for each home atom (warp-uniform) the lanes sweep the staged atoms 32 at a time and
//     compare an fp32 distance against a 1e-4 enlarged cutoff (full-rate fp32); accepted (i, j)
//     index pairs are compacted into a per-warp queue with a ballot;
//   * HEAVY: whenever 32 pairs are queued each lane takes one: it reloads both atoms in fp64, forms
//     the separation exactly in the oracle's operation order, applies the exact cutoff test,
//     evaluates the bias at r, hashes the two hill proposals and scatters the forces.
// The prefilter only decides what reaches the exact test, so results do not depend on it.
// Home-atom forces: segmented shuffle scan over the batch (entries of one atom are contiguous),
// the last lane of each run adds the run total to a shared-memory accumulator; neighbour-atom
// forces leave as fp64 REDs.
#ifndef EDM_PAIR_MINBLOCKS
#define EDM_PAIR_MINBLOCKS 5
#endif
constexpr int kPairWarps = 4;
constexpr int kJCap = 192;
constexpr int kICap = 64;
constexpr int kQCap = 64;

struct PairWarpSmem {
  double fi[kICap][3];
  float jx[kJCap], jy[kJCap], jz[kJCap];
  float ix[kICap], iy[kICap], iz[kICap];
  int jslot[kJCap];
  int queue[kQCap];            // (ii << 16) | jj
  unsigned char jcode[kJCap];  // image shift code, 2 bits per dim
};

struct PairCtx {
  GridDesc g;
  CellGrid cg;
  PairParams pp;
  const int* start;
  const int* order;
  const double* xs;
  const float* xs32;  // cell-relative fp32 copy of xs
  const int* ts;
  double* f;
  double* partial;
  BiasDev* st;
  HillAccepted* acc;
};

__device__ __noinline__ double pair_eval_slow(const GridDesc& g, double r, double& force) {
  return pair_eval(g, r, force);
}

// lean 1-D evaluation for the pair path: the arithmetic of d_eval_point<1> for an interpolated
// non-periodic grid; everything else (remap, periodic wrap) goes through the generic routine
__device__ __forceinline__ double pair_eval_fast(const GridDesc& g, double r, double& force) {
  force = 0.0;
  if (r < g.bmin[0] || r > g.bmax[0] || g.periodic[0] || !g.b_interp) return pair_eval_slow(g, r, force);
  if (r < g.min[0] || r >= g.upper[0]) return 0.0;
  const double t = __dsub_rn(r, g.min[0]);
  int idx = (int)(t * g.inv_dx[0]);  // a one-off at a cell edge is harmless: the interpolant is C1 there
  const int hi = g.n[0] - 2;
  idx = idx < 0 ? 0 : (idx > hi ? hi : idx);
  const double where = __dsub_rn(t, __dmul_rn((double)idx, g.dx[0]));
  const double X = where * g.inv_dx[0];
  const double2 r0 = *reinterpret_cast<const double2*>(g.rec + (long)idx * 2);
  const double2 r1 = *reinterpret_cast<const double2*>(g.rec + (long)idx * 2 + 2);
  const double Y = fabs(X - 1.0);
  const double X2 = X * X, X3 = X2 * X, Y2 = Y * Y, Y3 = Y2 * Y;
  const double td0 = !(fabs(r0.x) < kInterpZero) ? r0.y : 0.0;
  const double td1 = !(fabs(r1.x) < kInterpZero) ? r1.y : 0.0;
  double f = r0.x * (1.0 - 3.0 * X2 + 2.0 * X3) + td0 * (X - 2.0 * X2 + X3) * g.dx[0];
  double der = r0.x * (-6.0 * X + 6.0 * X2) * g.inv_dx[0] + td0 * (1.0 - 4.0 * X + 3.0 * X2);
  f += r1.x * (1.0 - 3.0 * Y2 + 2.0 * Y3) - td1 * (Y - 2.0 * Y2 + Y3) * g.dx[0];
  der += -r1.x * (-6.0 * Y + 6.0 * Y2) * g.inv_dx[0] + td1 * (1.0 - 4.0 * Y + 3.0 * Y2);
  force = -der;
  return f;
}

// hill proposals of one pair; the exactly rounded sqrt is taken only for an accepted proposal
__device__ __forceinline__ void propose_hills_d2(const PairParams& pp, unsigned long long pairkey, double d2,
                                                 BiasDev* st, HillAccepted* acc) {
  const uint64_t bits = pair_bits(pp.key, pairkey);
  const bool t0 = pp.accept_all || (bits >> 32) < pp.thresh_bits;
  const bool t1 = pp.accept_all || (bits & 0xffffffffULL) < pp.thresh_bits;
  if (t0 || t1) {
    const double r = sqrt(d2);
#pragma unroll
    for (int which = 0; which < 2; which++) {
      if (which == 0 ? t0 : t1) {
        int slot = atomicAdd(&st->n_accepted, 1);
        if (slot < pp.acc_cap) {
          acc[slot].key = 2ULL * pairkey + which;
          acc[slot].x[0] = r;
          acc[slot].x[1] = 0.0;
          acc[slot].x[2] = 0.0;
        } else {
          st->accepted_overflow = 1;
        }
      }
    }
  }
}

__device__ __noinline__ void pair_heavy_batch(const PairCtx& c, PairWarpSmem& w, int first, int count, int ilo,
                                              double& e, unsigned long long& npairs) {
  const int lane = threadIdx.x & 31;
  bool on = lane < count;
  int ii = -1;
  double px = 0.0, py = 0.0, pz = 0.0;
  if (on) {
    const int q = w.queue[first + lane];
    ii = q >> 16;
    const int jj = q & 0xffff;
    const long si = ilo + ii, sj = w.jslot[jj];
    const int code = w.jcode[jj];
    const double sx = (code & 1) ? c.cg.box[0] : ((code & 2) ? -c.cg.box[0] : 0.0);
    const double sy = (code & 4) ? c.cg.box[1] : ((code & 8) ? -c.cg.box[1] : 0.0);
    const double sz = (code & 16) ? c.cg.box[2] : ((code & 32) ? -c.cg.box[2] : 0.0);
    // exactly the oracle's separation: (x_i - x_j) - image shift, squares summed in x, y, z order
    const double dx = __dsub_rn(__dsub_rn(c.xs[3 * si + 0], c.xs[3 * sj + 0]), sx);
    const double dy = __dsub_rn(__dsub_rn(c.xs[3 * si + 1], c.xs[3 * sj + 1]), sy);
    const double dz = __dsub_rn(__dsub_rn(c.xs[3 * si + 2], c.xs[3 * sj + 2]), sz);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    on = d2 < c.pp.rc2;
    if (on) {
      const int oi = c.order[si], oj = c.order[sj];
      const double rinv = rsqrt(d2);
      const double r = d2 * rinv;  // within 2 ulp of sqrt(d2): enough for V(r); hills take the exact root
      double force = 1.0;
      if (!(c.pp.dbg & 2)) e += pair_eval_fast(c.g, r, force);
      npairs++;
      const double s = rinv * force;
      px = dx * s;
      py = dy * s;
      pz = dz * s;
      if (c.pp.do_hills && !(c.pp.dbg & 4)) {
        const unsigned long long lo = oi < oj ? oi : oj, hi = oi < oj ? oj : oi;
        propose_hills_d2(c.pp, lo * (unsigned long long)c.pp.natoms + hi, d2, c.st, c.acc);
      }
      if (!(c.pp.dbg & 1)) {
      atomicAdd(&c.f[3 * (long)oj + 0], -px);
      atomicAdd(&c.f[3 * (long)oj + 1], -py);
      atomicAdd(&c.f[3 * (long)oj + 2], -pz);
      }
    }
  }
  if (c.pp.dbg & 16) {
    if (on) {
      const long o = 3 * (long)c.order[ilo + ii];
      atomicAdd(&c.f[o + 0], px);
      atomicAdd(&c.f[o + 1], py);
      atomicAdd(&c.f[o + 2], pz);
    }
    return;
  }
  // home-atom side: inclusive segmented scan over runs of equal ii (contiguous in the queue)
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int iu = __shfl_up_sync(0xffffffffu, ii, o);
    const double ux = __shfl_up_sync(0xffffffffu, px, o);
    const double uy = __shfl_up_sync(0xffffffffu, py, o);
    const double uz = __shfl_up_sync(0xffffffffu, pz, o);
    if (lane >= o && iu == ii) {
      px += ux;
      py += uy;
      pz += uz;
    }
  }
  const int inext = __shfl_down_sync(0xffffffffu, ii, 1);
  if (ii >= 0 && (lane == 31 || inext != ii)) {
    w.fi[ii][0] += px;
    w.fi[ii][1] += py;
    w.fi[ii][2] += pz;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kPairWarps * 32, EDM_PAIR_MINBLOCKS) pair_cells_v4_kernel(const __grid_constant__ PairCtx c) {
  __shared__ PairWarpSmem wsm[kPairWarps];
  __shared__ double red[33];
  PairWarpSmem& w = wsm[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kPairWarps + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * kPairWarps;
  const CellGrid& cg = c.cg;
  const PairParams& pp = c.pp;
  double e = 0.0;
  unsigned long long npairs = 0;
  const float rc2m = (float)(pp.rc2 * 1.0001) + 1e-6f;  // prefilter radius; the fp32 error is ~1e-6 relative
  const float csx = (float)cg.cs[0], csy = (float)cg.cs[1], csz = (float)cg.cs[2];

  for (int cell = gw; cell < cg.ncell; cell += nwarps) {
    const int clo = c.start[cell], chi = c.start[cell + 1];
    if (clo == chi) continue;
    const int cx = cell % cg.nc[0], cy = (cell / cg.nc[0]) % cg.nc[1], cz = cell / (cg.nc[0] * cg.nc[1]);
    for (int ilo = clo; ilo < chi; ilo += kICap) {
      const int ni = min(kICap, chi - ilo);
      for (int a = lane; a < ni; a += 32) {
        w.ix[a] = c.xs32[3 * (long)(ilo + a) + 0] + csx;
        w.iy[a] = c.xs32[3 * (long)(ilo + a) + 1] + csy;
        w.iz[a] = c.xs32[3 * (long)(ilo + a) + 2] + csz;
        w.fi[a][0] = 0.0;
        w.fi[a][1] = 0.0;
        w.fi[a][2] = 0.0;
      }
      int nj = 0, nq = 0;
      // own-cell atoms are always staged at the head of the list: entries [0, self_hi) hold slots
      // self_slot0, self_slot0 + 1, ...; of those only partners with a larger slot than i count
      int self_hi = 0, self_slot0 = 0;
      __syncwarp();

      // TEST (+ HEAVY on full batches) over the staged list; drains the queue at the end because
      // queue entries index the staged list
      auto run_tests = [&]() {
        __syncwarp();
        for (int ii = 0; ii < ni; ii++) {
          const int si = ilo + ii;
          int ti = 0;
          if (pp.use_types) {
            ti = c.ts[si];
            if (ti != pp.itype && ti != pp.jtype) continue;
          }
          const float xi = w.ix[ii], yi = w.iy[ii], zi = w.iz[ii];
          const int selfcut = si - self_slot0;
          // four staged atoms per lane and trip: independent fp32 chains hide the LDS latency and
          // the loop/queue bookkeeping is paid once per 128 tests
          for (int jb = 0; jb < nj; jb += 128) {
            bool take[4];
            int jjs[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const int jj = jb + 32 * u + lane;
              jjs[u] = jj;
              take[u] = false;
              if (jj < nj) {
                const float dx = xi - w.jx[jj], dy = yi - w.jy[jj], dz = zi - w.jz[jj];
                const float d2 = dx * dx + dy * dy + dz * dz;
                take[u] = (d2 < rc2m) && (jj >= self_hi || jj > selfcut);
                if (take[u] && pp.use_types) {
                  const int tj = c.ts[w.jslot[jj]];
                  take[u] = (ti == pp.itype) ? (tj == pp.jtype) : (tj == pp.itype);
                }
              }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const unsigned m = __ballot_sync(0xffffffffu, take[u]);
              if (m == 0u) continue;
              if (take[u]) w.queue[nq + __popc(m & ((1u << lane) - 1u))] = (ii << 16) | jjs[u];
              nq += __popc(m);
              if (nq >= 32) {
                __syncwarp();
                nq -= 32;
                if (!(pp.dbg & 8)) pair_heavy_batch(c, w, nq, 32, ilo, e, npairs);
              }
            }
          }
        }
        __syncwarp();
        if (nq && !(pp.dbg & 8)) pair_heavy_batch(c, w, 0, nq, ilo, e, npairs);
        nq = 0;
      };

      // half shell as 5 x-rows: (oy,oz) = (0,0) with ox = 0..1 (own cell first), then
      // (1,0), (-1,1), (0,1), (1,1) with ox = -1..1
      for (int row = 0; row < 5; row++) {
        const int oy = (row == 0) ? 0 : (row == 1 ? 1 : row - 3);
        const int oz = (row < 2) ? 0 : 1;
        const int ox0 = (row == 0) ? 0 : -1;
        int qy = cy + oy, qz = cz + oz, ycode = 0;
        if (qy >= cg.nc[1]) { qy -= cg.nc[1]; ycode |= 4; } else if (qy < 0) { qy += cg.nc[1]; ycode |= 8; }
        if (qz >= cg.nc[2]) { qz -= cg.nc[2]; ycode |= 16; } else if (qz < 0) { qz += cg.nc[2]; ycode |= 32; }
        const float offy = (float)(oy + 1) * csy, offz = (float)(oz + 1) * csz;
        const int rowbase = (qz * cg.nc[1] + qy) * cg.nc[0];
        for (int ox = ox0; ox <= 1; ox++) {
          int qx = cx + ox, code = ycode;
          if (qx >= cg.nc[0]) { qx -= cg.nc[0]; code |= 1; } else if (qx < 0) { qx += cg.nc[0]; code |= 2; }
          // merge the following cells of the row while they stay contiguous (no wrap in between)
          int ncell = 1;
          while (ox + ncell <= 1 && qx + ncell < cg.nc[0] && !(code & 2)) ncell++;
          if ((code & 2)) ncell = 1;  // the wrapped-low cell stands alone; the rest restarts at qx = 0
          const int q0 = rowbase + qx;
          const int b0 = c.start[q0];
          const int b1 = c.start[q0 + 1];
          const int b2 = (ncell > 1) ? c.start[q0 + 2] : b1;
          const int b3 = (ncell > 2) ? c.start[q0 + 3] : b2;
          const float offx0 = (float)(ox + 1) * csx;
          const bool own = (row == 0 && ox == 0);
          int lo = b0;
          const int hi = b3;
          while (lo < hi) {
            if (nj == kJCap) {
              run_tests();
              nj = 0;
              self_hi = 0;
              __syncwarp();
            }
            const int takeN = min(kJCap - nj, hi - lo);
            if (own && lo < b1) {  // nj == 0 here: the own cell opens the list, also after a refill
              self_slot0 = lo;
              self_hi = min(takeN, b1 - lo);
            }
            for (int a = lane; a < takeN; a += 32) {
              const int sl = lo + a;
              const float offx = offx0 + (float)((sl >= b1) + (sl >= b2)) * csx;
              w.jx[nj + a] = c.xs32[3 * (long)sl + 0] + offx;
              w.jy[nj + a] = c.xs32[3 * (long)sl + 1] + offy;
              w.jz[nj + a] = c.xs32[3 * (long)sl + 2] + offz;
              w.jslot[nj + a] = sl;
              w.jcode[nj + a] = (unsigned char)code;
            }
            nj += takeN;
            lo += takeN;
          }
          ox += ncell - 1;
        }
      }
      if (nj) run_tests();
      __syncwarp();
      for (int a = lane; a < ni; a += 32) {
        const long o = 3 * (long)c.order[ilo + a];
        atomicAdd(&c.f[o + 0], w.fi[a][0]);
        atomicAdd(&c.f[o + 1], w.fi[a][1]);
        atomicAdd(&c.f[o + 2], w.fi[a][2]);
      }
      __syncwarp();
    }
  }
  double tot = block_sum(e, red);
  if (threadIdx.x == 0) c.partial[blockIdx.x] = tot;
  for (int o = 16; o > 0; o >>= 1) npairs += __shfl_down_sync(0xffffffffu, npairs, o);
  if (lane == 0 && npairs) atomicAdd(&c.st->n_pairs, npairs);
}

// ---- v5: the same search split in two kernels -------------------------------------------------
//
// pair_find_kernel   STAGE + TEST of v4; instead of a shared-memory queue feeding an in-kernel heavy
//                    path it appends the candidate pairs (slot_i, slot_j | shift code) to a global
//                    list, in chunks a warp allocates with one atomic per 256 candidates.
// pair_eval_kernel   one lane per candidate, fully populated warps, no shared memory: exact fp64
//                    separation + cutoff, bias at r, hill proposals, force scatter.  Candidates of a
//                    home atom are contiguous in a chunk, so the home side is a segmented shuffle scan
//                    with one RED triple per run; the neighbour side is one RED triple per pair.
// Splitting lets each kernel run at its own occupancy (the search needs shared memory and few
// registers, the evaluation the opposite) at the price of 8 B written + read per candidate.
constexpr int kChunk = 256;

struct FindWarpSmem {
  float jx[kJCap], jy[kJCap], jz[kJCap];
  float ix[kICap], iy[kICap], iz[kICap];
  int jpack[kJCap];  // slot | code << 26
};

struct CandList {
  int2* items;                  // x = slot_i (-1: padding), y = slot_j | code << 26
  unsigned long long cap;       // entries
  unsigned long long* count;    // entries allocated so far (multiple of kChunk)
  int* overflow;
};

__global__ void __launch_bounds__(kPairWarps * 32, 8) pair_find_kernel(const __grid_constant__ PairCtx c, CandList cl) {
  __shared__ FindWarpSmem wsm[kPairWarps];
  FindWarpSmem& w = wsm[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kPairWarps + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * kPairWarps;
  const CellGrid& cg = c.cg;
  const PairParams& pp = c.pp;
  const float rc2m = (float)(pp.rc2 * 1.0001) + 1e-6f;
  const float csx = (float)cg.cs[0], csy = (float)cg.cs[1], csz = (float)cg.cs[2];
  long long chunk_base = -1;  // -1: candidates are dropped (list full)
  int used = kChunk;          // forces an allocation on first use

  for (int cell = gw; cell < cg.ncell; cell += nwarps) {
    const int clo = c.start[cell], chi = c.start[cell + 1];
    if (clo == chi) continue;
    const int cx = cell % cg.nc[0], cy = (cell / cg.nc[0]) % cg.nc[1], cz = cell / (cg.nc[0] * cg.nc[1]);
    for (int ilo = clo; ilo < chi; ilo += kICap) {
      const int ni = min(kICap, chi - ilo);
      for (int a = lane; a < ni; a += 32) {
        w.ix[a] = c.xs32[3 * (long)(ilo + a) + 0] + csx;
        w.iy[a] = c.xs32[3 * (long)(ilo + a) + 1] + csy;
        w.iz[a] = c.xs32[3 * (long)(ilo + a) + 2] + csz;
      }
      int nj = 0;
      int self_hi = 0, self_slot0 = 0;
      __syncwarp();

      auto run_tests = [&]() {
        __syncwarp();
        for (int ii = 0; ii < ni; ii++) {
          const int si = ilo + ii;
          int ti = 0;
          if (pp.use_types) {
            ti = c.ts[si];
            if (ti != pp.itype && ti != pp.jtype) continue;
          }
          const float xi = w.ix[ii], yi = w.iy[ii], zi = w.iz[ii];
          const int selfcut = si - self_slot0;
          for (int jb = 0; jb < nj; jb += 128) {
            bool take[4];
            int pk[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const int jj = jb + 32 * u + lane;
              take[u] = false;
              pk[u] = 0;
              if (jj < nj) {
                const float dx = xi - w.jx[jj], dy = yi - w.jy[jj], dz = zi - w.jz[jj];
                const float d2 = dx * dx + dy * dy + dz * dz;
                take[u] = (d2 < rc2m) && (jj >= self_hi || jj > selfcut);
                pk[u] = w.jpack[jj];
                if (take[u] && pp.use_types) {
                  const int tj = c.ts[pk[u] & 0x3ffffff];
                  take[u] = (ti == pp.itype) ? (tj == pp.jtype) : (tj == pp.itype);
                }
              }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const unsigned m = __ballot_sync(0xffffffffu, take[u]);
              if (m == 0u) continue;
              const int k = __popc(m);
              if (used + k > kChunk) {  // close the chunk with padding and take a new one
                if (chunk_base >= 0)
                  for (int a = used + lane; a < kChunk; a += 32) cl.items[chunk_base + a] = make_int2(-1, 0);
                unsigned long long nb = 0;
                if (lane == 0) nb = atomicAdd(cl.count, (unsigned long long)kChunk);
                nb = __shfl_sync(0xffffffffu, nb, 0);
                if (nb + kChunk > cl.cap) {
                  chunk_base = -1;
                  if (lane == 0) *cl.overflow = 1;
                } else {
                  chunk_base = (long long)nb;
                }
                used = 0;
              }
              if (take[u] && chunk_base >= 0)
                cl.items[chunk_base + used + __popc(m & ((1u << lane) - 1u))] = make_int2(si, pk[u]);
              used += k;
            }
          }
        }
      };

      for (int row = 0; row < 5; row++) {
        const int oy = (row == 0) ? 0 : (row == 1 ? 1 : row - 3);
        const int oz = (row < 2) ? 0 : 1;
        const int ox0 = (row == 0) ? 0 : -1;
        int qy = cy + oy, qz = cz + oz, ycode = 0;
        if (qy >= cg.nc[1]) { qy -= cg.nc[1]; ycode |= 4; } else if (qy < 0) { qy += cg.nc[1]; ycode |= 8; }
        if (qz >= cg.nc[2]) { qz -= cg.nc[2]; ycode |= 16; } else if (qz < 0) { qz += cg.nc[2]; ycode |= 32; }
        const float offy = (float)(oy + 1) * csy, offz = (float)(oz + 1) * csz;
        const int rowbase = (qz * cg.nc[1] + qy) * cg.nc[0];
        for (int ox = ox0; ox <= 1; ox++) {
          int qx = cx + ox, code = ycode;
          if (qx >= cg.nc[0]) { qx -= cg.nc[0]; code |= 1; } else if (qx < 0) { qx += cg.nc[0]; code |= 2; }
          int ncell = 1;
          while (ox + ncell <= 1 && qx + ncell < cg.nc[0] && !(code & 2)) ncell++;
          if ((code & 2)) ncell = 1;
          const int q0 = rowbase + qx;
          const int b0 = c.start[q0];
          const int b1 = c.start[q0 + 1];
          const int b2 = (ncell > 1) ? c.start[q0 + 2] : b1;
          const int b3 = (ncell > 2) ? c.start[q0 + 3] : b2;
          const float offx0 = (float)(ox + 1) * csx;
          const bool own = (row == 0 && ox == 0);
          int lo = b0;
          const int hi = b3;
          while (lo < hi) {
            if (nj == kJCap) {
              run_tests();
              nj = 0;
              self_hi = 0;
              __syncwarp();
            }
            const int takeN = min(kJCap - nj, hi - lo);
            if (own && lo < b1) {
              self_slot0 = lo;
              self_hi = min(takeN, b1 - lo);
            }
            for (int a = lane; a < takeN; a += 32) {
              const int sl = lo + a;
              const float offx = offx0 + (float)((sl >= b1) + (sl >= b2)) * csx;
              w.jx[nj + a] = c.xs32[3 * (long)sl + 0] + offx;
              w.jy[nj + a] = c.xs32[3 * (long)sl + 1] + offy;
              w.jz[nj + a] = c.xs32[3 * (long)sl + 2] + offz;
              w.jpack[nj + a] = sl | (code << 26);
            }
            nj += takeN;
            lo += takeN;
          }
          ox += ncell - 1;
        }
      }
      if (nj) run_tests();
      __syncwarp();
    }
  }
  if (chunk_base >= 0)
    for (int a = used + lane; a < kChunk; a += 32) cl.items[chunk_base + a] = make_int2(-1, 0);
}

__global__ void __launch_bounds__(256) pair_eval_kernel(const __grid_constant__ PairCtx c, CandList cl) {
  __shared__ double red[33];
  const int lane = threadIdx.x & 31;
  unsigned long long total = *cl.count;
  if (total > cl.cap) total = cl.cap - cl.cap % kChunk;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  double e = 0.0;
  unsigned long long npairs = 0;
  for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < total;
       base += stride) {
    const int2 it = cl.items[base + lane];  // total is a multiple of 32
    int si = it.x;
    int oi = -1;
    bool on = si >= 0;
    double px = 0.0, py = 0.0, pz = 0.0;
    if (on) {
      const long sj = it.y & 0x3ffffff;
      const int code = (int)((unsigned)it.y >> 26);
      const double sx = (code & 1) ? c.cg.box[0] : ((code & 2) ? -c.cg.box[0] : 0.0);
      const double sy = (code & 4) ? c.cg.box[1] : ((code & 8) ? -c.cg.box[1] : 0.0);
      const double sz = (code & 16) ? c.cg.box[2] : ((code & 32) ? -c.cg.box[2] : 0.0);
      // exactly the oracle's separation: (x_i - x_j) - image shift, squares summed in x, y, z order
      const double dx = __dsub_rn(__dsub_rn(c.xs[3 * (long)si + 0], c.xs[3 * sj + 0]), sx);
      const double dy = __dsub_rn(__dsub_rn(c.xs[3 * (long)si + 1], c.xs[3 * sj + 1]), sy);
      const double dz = __dsub_rn(__dsub_rn(c.xs[3 * (long)si + 2], c.xs[3 * sj + 2]), sz);
      const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      on = d2 < c.pp.rc2;
      oi = c.order[si];
      if (on) {
        const int oj = c.order[sj];
        const double rinv = rsqrt(d2);
        const double r = d2 * rinv;
        double force;
        e += pair_eval_fast(c.g, r, force);
        npairs++;
        const double s = rinv * force;
        px = dx * s;
        py = dy * s;
        pz = dz * s;
        if (c.pp.do_hills) {
          const unsigned long long lo = oi < oj ? oi : oj, hi = oi < oj ? oj : oi;
          propose_hills_d2(c.pp, lo * (unsigned long long)c.pp.natoms + hi, d2, c.st, c.acc);
        }
        atomicAdd(&c.f[3 * (long)oj + 0], -px);
        atomicAdd(&c.f[3 * (long)oj + 1], -py);
        atomicAdd(&c.f[3 * (long)oj + 2], -pz);
      }
    }
    // home side: segmented inclusive scan (head flags: a slot_i may reappear later in the list when a
    // neighbourhood was staged in two parts), one RED triple per run
    const int iprev = __shfl_up_sync(0xffffffffu, si, 1);
    int head = (lane == 0 || iprev != si) ? 1 : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int uh = __shfl_up_sync(0xffffffffu, head, o);
      const double ux = __shfl_up_sync(0xffffffffu, px, o);
      const double uy = __shfl_up_sync(0xffffffffu, py, o);
      const double uz = __shfl_up_sync(0xffffffffu, pz, o);
      if (lane >= o && !head) {
        px += ux;
        py += uy;
        pz += uz;
        head = uh;
      }
    }
    const int inext = __shfl_down_sync(0xffffffffu, si, 1);
    if (si >= 0 && (lane == 31 || inext != si)) {
      atomicAdd(&c.f[3 * (long)oi + 0], px);
      atomicAdd(&c.f[3 * (long)oi + 1], py);
      atomicAdd(&c.f[3 * (long)oi + 2], pz);
    }
  }
  double tot = block_sum(e, red);
  if (threadIdx.x == 0) c.partial[blockIdx.x] = tot;
  for (int o = 16; o > 0; o >>= 1) npairs += __shfl_down_sync(0xffffffffu, npairs, o);
  if (lane == 0 && npairs) atomicAdd(&c.st->n_pairs, npairs);
}

__global__ void reset_cand_kernel(unsigned long long* count, int* overflow) {
  *count = 0;
  *overflow = 0;
}

// Neighbour-list form: one thread per listed i-row (lammps/fix_edm_pair.cpp:177-240).
__global__ void __launch_bounds__(128) pair_list_kernel(GridDesc g, PairParams pp, long nlocal, long inum,
                                                        const int* __restrict__ ilist, const long* __restrict__ first,
                                                        const int* __restrict__ jlist, const double* __restrict__ x,
                                                        const int* __restrict__ type, const double* __restrict__ runiform,
                                                        double* __restrict__ f, double* __restrict__ partial,
                                                        BiasDev* st, HillAccepted* acc, unsigned long long* ncalls) {
  __shared__ double red[33];
  double e = 0.0;
  unsigned long long npairs = 0, calls = 0;
  long stride = (long)gridDim.x * blockDim.x;
  for (long ii = (long)blockIdx.x * blockDim.x + threadIdx.x; ii < inum; ii += stride) {
    const int i = ilist[ii];
    int type_ind = 0;
    if (pp.use_types) {
      int it = type[i];
      if (it == pp.itype) type_ind = 1;
      else if (it == pp.jtype) type_ind = 0;
      else continue;
    }
    const double xi = x[3 * (long)i + 0], yi = x[3 * (long)i + 1], zi = x[3 * (long)i + 2];
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (long k = first[ii]; k < first[ii + 1]; k++) {
      const int j = jlist[k];
      if (pp.use_types) {
        int jt = type[j];
        if (type_ind && jt != pp.jtype) continue;
        if (!type_ind && jt != pp.itype) continue;
      }
      double dx = __dsub_rn(xi, x[3 * (long)j + 0]);
      double dy = __dsub_rn(yi, x[3 * (long)j + 1]);
      double dz = __dsub_rn(zi, x[3 * (long)j + 2]);
      double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      double r = sqrt(d2);
      double rinv = 1.0 / r;
      double force;
      e += pair_eval(g, r, force);
      npairs++;
      double px = dx * rinv * force, py = dy * rinv * force, pz = dz * rinv * force;
      fx += px;
      fy += py;
      fz += pz;
      const bool jlocal = j < nlocal;
      if (jlocal) {
        atomicAdd(&f[3 * (long)j + 0], -px);
        atomicAdd(&f[3 * (long)j + 1], -py);
        atomicAdd(&f[3 * (long)j + 2], -pz);
      }
      if (pp.do_hills) {
        int nprop = jlocal ? 2 : 1;
        calls += nprop;
        for (int which = 0; which < nprop; which++) {
          unsigned long long kk = 2ULL * (unsigned long long)k + which;
          double u = runiform ? runiform[kk] : uniform_from_key(pp.key, kk);
          if (pp.accept_all || u < pp.thresh) {
            int slot = atomicAdd(&st->n_accepted, 1);
            if (slot < pp.acc_cap) {
              acc[slot].key = kk;
              acc[slot].x[0] = r;
              acc[slot].x[1] = 0.0;
              acc[slot].x[2] = 0.0;
            } else {
              st->accepted_overflow = 1;
            }
          }
        }
      }
    }
    atomicAdd(&f[3 * (long)i + 0], fx);
    atomicAdd(&f[3 * (long)i + 1], fy);
    atomicAdd(&f[3 * (long)i + 2], fz);
  }
  double tot = block_sum(e, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
  for (int o = 16; o > 0; o >>= 1) {
    npairs += __shfl_down_sync(0xffffffffu, npairs, o);
    calls += __shfl_down_sync(0xffffffffu, calls, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (npairs) atomicAdd(&st->n_pairs, npairs);
    if (calls) atomicAdd(ncalls, calls);
  }
}

__global__ void reset_pairs_kernel(BiasDev* st, unsigned long long* ncalls) {
  st->n_pairs = 0;
  if (ncalls) *ncalls = 0;
}

__global__ void sum_partials2_kernel(int n, const double* __restrict__ partial, double* out) {
  __shared__ double red[33];
  double e = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) e += partial[i];
  double tot = block_sum(e, red);
  if (threadIdx.x == 0) out[0] = tot;
}

}  // namespace edm

using namespace edm;

static PairParams pair_params(const edm_bias* b, const int* type, int itype, int jtype, int do_hills, long long est,
                              uint64_t seed, uint64_t step, double cutoff, long natoms) {
  PairParams pp;
  pp.itype = itype;
  pp.jtype = jtype;
  pp.use_types = type != nullptr;
  pp.do_hills = do_hills;
  pp.accept_all = b->prm.hill_density < 0;
  pp.thresh = pp.accept_all ? 2.0 : b->prm.hill_density / (double)(int)est;
  {
    double tb = ceil(pp.thresh * 4294967296.0);
    pp.thresh_bits = tb >= 4294967296.0 ? 4294967296ULL : (tb <= 0.0 ? 0ULL : (uint64_t)tb);
  }
  pp.key = uniform_key(seed, step);
  pp.dbg = getenv("EDM_DBG") ? atoi(getenv("EDM_DBG")) : 0;
  pp.rc2 = cutoff * cutoff;
  pp.natoms = natoms;
  pp.acc_cap = b->accepted_cap;
  return pp;
}

// binning + pair kernel on device pointers; leaves the energy in b->d_scalar[0]
static int pair_cells_launch(edm_bias* b, long natoms, const double* x, double* f, const int* type, int itype,
                             int jtype, const double* box, double cutoff, int do_hills, long long est, uint64_t seed,
                             uint64_t step, double* energy_dev, cudaStream_t st) {
  EDM_REQUIRE(b->prm.dim == 1, "Pairwise distance must be 1 dimension in EDM input file");  // fix_edm_pair.cpp:52-53
  EDM_REQUIRE(natoms > 0 && natoms < 2000000000L, "bad atom count");
  CellGrid cg;
  long long ncell = 1;
  for (int d = 0; d < 3; d++) {
    cg.box[d] = box[d];
    cg.nc[d] = (int)floor(box[d] / cutoff);
    EDM_REQUIRE(cg.nc[d] >= 3, "box must hold at least 3 cutoffs per side for the half-shell cell search");
    cg.cs[d] = box[d] / cg.nc[d];
    ncell *= cg.nc[d];
  }
  EDM_REQUIRE(ncell < 2000000000LL, "too many cells");
  cg.ncell = (int)ncell;
  // scratch: cell_of[n], order[n], ts[n], count[ncell+1], start[ncell+1], xs[3n]
  size_t n = (size_t)natoms, nc1 = (size_t)ncell + 1;
  size_t off_cell = 0, off_order = off_cell + n * 4, off_ts = off_order + n * 4, off_count = off_ts + n * 4;
  size_t off_start = off_count + nc1 * 4;
  size_t off_tiles = off_start + nc1 * 4;
  size_t off_xs = (off_tiles + ((size_t)ncell / 2048 + 2) * 4 + 255) / 256 * 256;
  size_t off_xs32 = off_xs + 3 * n * sizeof(double);
  size_t total = off_xs32 + 3 * n * sizeof(float);
  EDM_TRY(b->cells.reserve(total));
  char* base = b->cells.as<char>();
  int* cell_of = reinterpret_cast<int*>(base + off_cell);
  int* order = reinterpret_cast<int*>(base + off_order);
  int* ts = reinterpret_cast<int*>(base + off_ts);
  int* count = reinterpret_cast<int*>(base + off_count);
  int* start = reinterpret_cast<int*>(base + off_start);
  int* tile_total = reinterpret_cast<int*>(base + off_tiles);
  double* xs = reinterpret_cast<double*>(base + off_xs);
  float* xs32 = reinterpret_cast<float*>(base + off_xs32);

  EDM_CUDA(cudaMemsetAsync(count, 0, nc1 * 4, st));
  unsigned nb = (unsigned)((natoms + 255) / 256);
  cell_count_kernel<<<nb, 256, 0, st>>>(natoms, x, cg, cell_of, count);
  const int ntiles = (cg.ncell + kScanTile - 1) / kScanTile;
  cell_scan_totals_kernel<<<ntiles, 256, 0, st>>>(cg.ncell, count, tile_total);
  cell_scan_kernel<<<ntiles, 256, 0, st>>>(cg.ncell, count, tile_total, start);
  EDM_CUDA(cudaMemsetAsync(count, 0, nc1 * 4, st));
  cell_fill_kernel<<<nb, 256, 0, st>>>(natoms, cell_of, start, count, order);
  cell_sort_gather_kernel<<<(cg.ncell + 127) / 128, 128, 0, st>>>(cg, start, order, x, type, xs, xs32, type ? ts : nullptr);
  EDM_CUDA(cudaGetLastError());

  PairParams pp = pair_params(b, type, itype, jtype, do_hills, est, seed, step, cutoff, natoms);
  reset_pairs_kernel<<<1, 1, 0, st>>>(b->d_state, nullptr);
  // EDM_PAIR_MODE=0 selects the v1 thread-per-atom kernel (kept for A/B measurements)
  static int mode = -1;
  if (mode < 0) {
    const char* ev = getenv("EDM_PAIR_MODE");
    mode = ev ? atoi(ev) : 2;
  }
  long long blocks;
  if (b->profiling) EDM_CUDA(cudaEventRecord(b->ev_pair[0], st));
  if (mode == 0) {
    blocks = (natoms + 127) / 128;
    if (blocks > b->n_partial) blocks = b->n_partial;
    pair_cells_kernel<<<(int)blocks, 128, 0, st>>>(b->bias->d, cg, pp, start, order, xs, ts, f, b->d_energy_partial,
                                                   b->d_state, b->d_accepted);
  } else {
    blocks = 148 * 8;  // one wave of 8 resident CTAs (32 warps) per SM; warps stride over the cells
    if (blocks > (cg.ncell + kPairWarps - 1) / kPairWarps) blocks = (cg.ncell + kPairWarps - 1) / kPairWarps;
    if (blocks > b->n_partial) blocks = b->n_partial;
    PairCtx ctx;
    ctx.g = b->bias->d;
    ctx.cg = cg;
    ctx.pp = pp;
    ctx.start = start;
    ctx.order = order;
    ctx.xs = xs;
    ctx.xs32 = xs32;
    ctx.ts = ts;
    ctx.f = f;
    ctx.partial = b->d_energy_partial;
    ctx.st = b->d_state;
    ctx.acc = b->d_accepted;
    if (mode == 1) {
      pair_cells_v4_kernel<<<(int)blocks, kPairWarps * 32, 0, st>>>(ctx);
    } else {
      // candidate list: expected pairs at uniform density x 1.5 + chunk slack; an overflow is flagged
      double vol = box[0] * box[1] * box[2];
      double expect = 0.5 * (double)natoms * ((double)natoms / vol) * (4.0 / 3.0) * M_PI * cutoff * cutoff * cutoff;
      unsigned long long cap = (unsigned long long)(1.5 * expect) + 8ULL * 148 * 8 * kPairWarps * kChunk / 8 + (1u << 16);
      cap = (cap + kChunk - 1) / kChunk * kChunk;
      EDM_TRY(b->cand.reserve(cap * sizeof(int2) + 64));
      CandList cl;
      cl.count = b->cand.as<unsigned long long>();
      cl.overflow = reinterpret_cast<int*>(b->cand.as<char>() + 8);
      cl.items = reinterpret_cast<int2*>(b->cand.as<char>() + 64);
      cl.cap = cap;
      reset_cand_kernel<<<1, 1, 0, st>>>(cl.count, cl.overflow);
      pair_find_kernel<<<(int)blocks, kPairWarps * 32, 0, st>>>(ctx, cl);
      int eblocks = 148 * 8;
      if (eblocks > b->n_partial) eblocks = b->n_partial;
      blocks = eblocks;
      pair_eval_kernel<<<eblocks, 256, 0, st>>>(ctx, cl);
      count_launches(2);
    }
  }
  if (b->profiling) EDM_CUDA(cudaEventRecord(b->ev_pair[1], st));
  count_launches(8);
  sum_partials2_kernel<<<1, 256, 0, st>>>((int)blocks, b->d_energy_partial, energy_dev);
  EDM_CUDA(cudaGetLastError());
  return EDM_OK;
}

static int read_pair_result(edm_bias* b, edm_pair_result_t* result, const unsigned long long* ncalls_dev) {
  if (!result) return EDM_OK;
  double e;
  unsigned long long np;
  EDM_CUDA(cudaMemcpy(&e, b->d_scalar, sizeof(double), cudaMemcpyDeviceToHost));
  EDM_CUDA(cudaMemcpy(&np, &b->d_state->n_pairs, sizeof(np), cudaMemcpyDeviceToHost));
  result->energy = e;
  result->n_pairs = (long long)np;
  result->n_calls = 2 * (long long)np;
  if (ncalls_dev) {
    unsigned long long nc;
    EDM_CUDA(cudaMemcpy(&nc, ncalls_dev, sizeof(nc), cudaMemcpyDeviceToHost));
    result->n_calls = (long long)nc;
  }
  return EDM_OK;
}

extern "C" {

int edm_pair_select_cells_dev(edm_bias_t* b, long natoms, const double* x, double* f, const int* type, int itype,
                              int jtype, const double* box, double cutoff, long long est_hill_count, uint64_t seed,
                              uint64_t step, double* energy_dev, void* stream) {
  EDM_REQUIRE(b && x && f && box, "NULL argument");
  EDM_TRY(ensure_device(b->device));
  cudaStream_t st = (cudaStream_t)stream;
  EDM_TRY(edm_bias_reset_accepted(b, st));
  return pair_cells_launch(b, natoms, x, f, type, itype, jtype, box, cutoff, 1, est_hill_count, seed, step,
                           energy_dev ? energy_dev : b->d_scalar, st);
}

int edm_pair_step_cells_dev(edm_bias_t* b, long natoms, const double* x, double* f, const int* type, int itype,
                            int jtype, const double* box, double cutoff, int do_hills, long long est_hill_count,
                            uint64_t seed, uint64_t step, edm_pair_result_t* result, void* stream) {
  EDM_REQUIRE(b && x && f && box, "NULL argument");
  EDM_TRY(ensure_device(b->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (do_hills) EDM_TRY(edm_bias_reset_accepted(b, st));
  EDM_TRY(pair_cells_launch(b, natoms, x, f, type, itype, jtype, box, cutoff, do_hills, est_hill_count, seed, step,
                            b->d_scalar, st));
  if (do_hills) EDM_TRY(edm_bias_launch_round(b, est_hill_count, st));
  if (result) {
    EDM_CUDA(cudaStreamSynchronize(st));
    EDM_TRY(read_pair_result(b, result, nullptr));
    if (do_hills) EDM_TRY(edm_bias_check_round(b));
  }
  return EDM_OK;
}

int edm_pair_step_cells(edm_bias_t* b, long natoms, const double* x, double* f, const int* type, int itype, int jtype,
                        const double* box, double cutoff, int do_hills, long long est_hill_count, uint64_t seed,
                        uint64_t step, edm_pair_result_t* result) {
  EDM_REQUIRE(b && x && f && box && natoms > 0, "bad argument");
  EDM_TRY(ensure_device(b->device));
  size_t bx = (size_t)natoms * 3 * sizeof(double);
  EDM_TRY(b->io.reserve(bx));
  EDM_TRY(b->io2.reserve(bx));
  const int* dt = nullptr;
  if (type) {
    EDM_TRY(b->io3.reserve((size_t)natoms * sizeof(int)));
    EDM_CUDA(cudaMemcpyAsync(b->io3.p, type, (size_t)natoms * sizeof(int), cudaMemcpyHostToDevice, 0));
    dt = b->io3.as<int>();
  }
  EDM_CUDA(cudaMemcpyAsync(b->io.p, x, bx, cudaMemcpyHostToDevice, 0));
  EDM_CUDA(cudaMemcpyAsync(b->io2.p, f, bx, cudaMemcpyHostToDevice, 0));
  edm_pair_result_t local;
  EDM_TRY(edm_pair_step_cells_dev(b, natoms, b->io.as<double>(), b->io2.as<double>(), dt, itype, jtype, box, cutoff,
                                  do_hills, est_hill_count, seed, step, &local, nullptr));
  EDM_CUDA(cudaMemcpy(f, b->io2.p, bx, cudaMemcpyDeviceToHost));
  if (result) *result = local;
  return EDM_OK;
}

int edm_pair_step_list(edm_bias_t* b, long nall, long nlocal, const double* x, double* f, const int* type, int itype,
                       int jtype, long inum, const int* ilist, const long* first, const int* jlist, int do_hills,
                       long long est_hill_count, const double* runiform, uint64_t seed, uint64_t step,
                       edm_pair_result_t* result) {
  EDM_REQUIRE(b && x && f && ilist && first && (jlist || first[inum] == 0), "NULL argument");
  EDM_REQUIRE(b->prm.dim == 1, "Pairwise distance must be 1 dimension in EDM input file");
  EDM_TRY(ensure_device(b->device));
  long nlisted = first[inum];
  size_t bx = (size_t)nall * 3 * sizeof(double);
  EDM_TRY(b->io.reserve(bx));
  EDM_TRY(b->io2.reserve(bx));
  // list scratch: ilist[inum] | first[inum+1] | jlist[nlisted] | type[nall] | ncalls | uniforms
  size_t o_il = 0, o_first = (o_il + (size_t)inum * 4 + 7) / 8 * 8, o_jl = o_first + (size_t)(inum + 1) * 8;
  size_t o_ty = (o_jl + (size_t)nlisted * 4 + 7) / 8 * 8, o_nc = (o_ty + (size_t)nall * 4 + 7) / 8 * 8;
  size_t o_u = o_nc + 8;
  size_t total = o_u + (runiform ? (size_t)nlisted * 2 * sizeof(double) : 0);
  EDM_TRY(b->io4.reserve(total));
  char* base = b->io4.as<char>();
  EDM_CUDA(cudaMemcpyAsync(base + o_il, ilist, (size_t)inum * 4, cudaMemcpyHostToDevice, 0));
  EDM_CUDA(cudaMemcpyAsync(base + o_first, first, (size_t)(inum + 1) * 8, cudaMemcpyHostToDevice, 0));
  if (nlisted) EDM_CUDA(cudaMemcpyAsync(base + o_jl, jlist, (size_t)nlisted * 4, cudaMemcpyHostToDevice, 0));
  if (type) EDM_CUDA(cudaMemcpyAsync(base + o_ty, type, (size_t)nall * 4, cudaMemcpyHostToDevice, 0));
  if (runiform)
    EDM_CUDA(cudaMemcpyAsync(base + o_u, runiform, (size_t)nlisted * 2 * sizeof(double), cudaMemcpyHostToDevice, 0));
  EDM_CUDA(cudaMemcpyAsync(b->io.p, x, bx, cudaMemcpyHostToDevice, 0));
  EDM_CUDA(cudaMemcpyAsync(b->io2.p, f, bx, cudaMemcpyHostToDevice, 0));
  unsigned long long* ncalls = reinterpret_cast<unsigned long long*>(base + o_nc);
  if (do_hills) EDM_TRY(edm_bias_reset_accepted(b, 0));
  PairParams pp = pair_params(b, type, itype, jtype, do_hills, est_hill_count, seed, step, 0.0, nall);
  reset_pairs_kernel<<<1, 1>>>(b->d_state, ncalls);
  long long blocks = (inum + 127) / 128;
  if (blocks > b->n_partial) blocks = b->n_partial;
  if (blocks < 1) blocks = 1;
  pair_list_kernel<<<(int)blocks, 128>>>(b->bias->d, pp, nlocal, inum, reinterpret_cast<int*>(base + o_il),
                                         reinterpret_cast<long*>(base + o_first), reinterpret_cast<int*>(base + o_jl),
                                         b->io.as<double>(), type ? reinterpret_cast<int*>(base + o_ty) : nullptr,
                                         runiform ? reinterpret_cast<double*>(base + o_u) : nullptr, b->io2.as<double>(),
                                         b->d_energy_partial, b->d_state, b->d_accepted, ncalls);
  sum_partials2_kernel<<<1, 256>>>((int)blocks, b->d_energy_partial, b->d_scalar);
  count_launches(3);
  EDM_CUDA(cudaGetLastError());
  if (do_hills) EDM_TRY(edm_bias_launch_round(b, est_hill_count, 0));
  EDM_CUDA(cudaMemcpy(f, b->io2.p, bx, cudaMemcpyDeviceToHost));
  EDM_TRY(read_pair_result(b, result, ncalls));
  if (do_hills) EDM_TRY(edm_bias_check_round(b));
  return EDM_OK;
}

}  // extern "C"
