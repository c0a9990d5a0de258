// Pair-distance collective variable of fix edm_pair (lammps/fix_edm_pair.cpp:139-256) on the
// device: cell binning, half-shell pair search, 1-D bias evaluation at r, force scatter and the
// two-per-pair hill proposals.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "edm_host.h"

namespace edm {

struct CellGrid {
  int nc[3];
  double cs[3];
  double box[3];   // extent of the cell grid per dimension (the period, in periodic dimensions)
  double lo[3];    // its lower corner
  int periodic[3]; // 0: the cell grid ends there — whatever lies beyond reaches this rank as ghost atoms
  int ncell;
};

struct PairParams {
  int itype, jtype, use_types;
  int do_hills, accept_all;
  int lean;  // interpolated, non-periodic grid and boundary: pair_eval_fast applies
  double thresh;
  uint64_t thresh_bits;  // ceil(thresh * 2^32), saturated
  int dbg;
  uint64_t key;
  double rc2;
  float rc2m;  // fp32 prefilter radius^2: cutoff^2 enlarged by 1e-4 (the fp32 error is ~1e-5 absolute)
  long natoms;
  long nlocal;  // atoms [nlocal, natoms) are ghosts: no force, one hill proposal, ghost-ghost pairs skipped
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
  int cx = cell_coord(x[3 * i + 0] - cg.lo[0], cg.cs[0], cg.nc[0]);
  int cy = cell_coord(x[3 * i + 1] - cg.lo[1], cg.cs[1], cg.nc[1]);
  int cz = cell_coord(x[3 * i + 2] - cg.lo[2], cg.cs[2], cg.nc[2]);
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

// 1-D bias evaluation at r with update_force's sign (lib/edm_bias.cpp:297-311): returns V(r),
// force = -dV/dr.
__device__ __forceinline__ double pair_eval(const GridDesc& g, double r, double& force) {
  double der[1];
  double rr[1] = {r};
  double v = d_eval_point<1>(g, rr, der, g.b_interp != 0);
  force = -der[0];
  return v;
}

__device__ __noinline__ double2 pair_eval_slow(const GridDesc& g, double r) {
  double force;
  const double v = pair_eval(g, r, force);
  return make_double2(v, force);
}

// lean 1-D evaluation for the pair path: the arithmetic of d_eval_point<1> for an interpolated
// non-periodic Gaussian grid with a non-periodic boundary (what fix edm_pair sets up,
// lammps/fix_edm_pair.cpp:96-104): outside the boundary remap changes nothing and the bias is 0
// (lib/gaussian_grid.h:128-135).  Anything else goes through the generic routine.
__device__ __forceinline__ double pair_eval_fast(const GridDesc& g, const double* __restrict__ cellrec, bool lean,
                                                 double r, double& force) {
  if (!lean) {
    const double2 vf = pair_eval_slow(g, r);
    force = vf.y;
    return vf.x;
  }
  force = 0.0;
  if (r < g.bmin[0] || r > g.bmax[0] || r < g.min[0] || r >= g.upper[0]) return 0.0;
  const double t = __dsub_rn(r, g.min[0]);
  int idx = (int)(t * g.inv_dx[0]);  // a one-off at a cell edge is harmless: the interpolant is C1 there
  const int hi = g.n[0] - 2;
  idx = idx < 0 ? 0 : (idx > hi ? hi : idx);
  const double where = __dsub_rn(t, __dmul_rn((double)idx, g.dx[0]));
  const double X = where * g.inv_dx[0];
  // both corners of the cell in one aligned 32 B record: one sector, one 256-bit load
  double2 r0, r1;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
      : "=d"(r0.x), "=d"(r0.y), "=d"(r1.x), "=d"(r1.y)
      : "l"(cellrec + (long)idx * 4));
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
                                                 BiasDev* st, HillAccepted* acc, int nprop = 2) {
  const uint64_t bits = pair_bits(pp.key, pairkey);
  const bool t0 = pp.accept_all || (bits >> 32) < pp.thresh_bits;
  const bool t1 = nprop > 1 && (pp.accept_all || (bits & 0xffffffffULL) < pp.thresh_bits);
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

// Atoms in cell order ("slots"): one 32 B record per atom, one aligned sector per gather.
struct __align__(32) AtomRec {
  double x, y, z;
  long long tag;  // original atom index; staged copies add the image-shift code in bits 32..37
};

// slot = start[cell] + (count[cell]-- - 1): reuses the histogram as the fill counter
__global__ void cell_fill_kernel(long n, const int* __restrict__ cell_of, const int* __restrict__ start,
                                 int* __restrict__ count, int* __restrict__ order) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = cell_of[i];
  int slot = start[c] + atomicSub(&count[c], 1) - 1;
  order[slot] = (int)i;
}

// Canonical order inside a cell (ascending atom index, so the slot order does not depend on how the
// atomics above resolved) + gather into slot order.  One thread per SLOT of the unordered fill: it finds the atom
// the fill put there, ranks it among its cell mates (the number of mates with a smaller index) and writes its record
// at that rank.  Threads of a warp share a handful of cells, so their rank loops read the same few words and their
// writes land in the same few hundred bytes; the coordinates are a gather from x, which the counting pass has just
// streamed through L2 (one thread per ATOM instead, with its scattered 48 B of writes, took 40 us of the step).
// xs32 = position relative to the cell corner in fp32 (error ~cs * 6e-8, box-size independent) with the type in
// .w, for the search prefilter.
__global__ void cell_rank_gather_kernel(long n, CellGrid cg, const int* __restrict__ cell_of,
                                        const int* __restrict__ start, const int* __restrict__ order,
                                        const double* __restrict__ x, const int* __restrict__ type,
                                        AtomRec* __restrict__ arec, float4* __restrict__ xs32) {
  const long slot = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n) return;
  const int i = order[slot];
  const int c = cell_of[i];
  const int lo = start[c], hi = start[c + 1];
  int rank = 0;
  for (int a = lo; a < hi; a++) rank += (order[a] < i);
  const long s = lo + rank;
  const double px = x[3L * i + 0], py = x[3L * i + 1], pz = x[3L * i + 2];
  const double cx = cg.lo[0] + (c % cg.nc[0]) * cg.cs[0], cy = cg.lo[1] + ((c / cg.nc[0]) % cg.nc[1]) * cg.cs[1];
  const double cz = cg.lo[2] + (c / (cg.nc[0] * cg.nc[1])) * cg.cs[2];
  AtomRec r;
  r.x = px;
  r.y = py;
  r.z = pz;
  r.tag = (long long)i;
  arec[s] = r;
  xs32[s] = make_float4((float)(px - cx), (float)(py - cy), (float)(pz - cz), __int_as_float(type ? type[i] : 0));
}

struct PairCtx {
  GridDesc g;
  CellGrid cg;
  PairParams pp;
  const int* start;
  const AtomRec* arec;
  const float4* xs32;
  double* f;
  BiasDev* st;
  HillAccepted* acc;
  int* fallback;  // set by the block search when a region does not fit: the direct search takes the step
  const double* cellrec;                // per grid cell {V_k, V'_k, V_k+1, V'_k+1}, 32 B aligned (pair_prep_kernel)
  const unsigned long long* fmax_bits;  // bit patterns of bounds on |dV/dr| ([0]) and |V| ([1]) over the grid
};

// Per step: the 1-D bias grid regrouped by cell, so the interpolation's two corners are one aligned
// 32 B gather, and a bound on |dV/dr|: inside a cell the Hermite derivative is
// dV/dx * 6X(1-X) + V'_k (1-4X+3X^2) + V'_k+1 (3X^2-2X), at most 1.5 |dV|/dx + max(|V'_k|, |V'_k+1|).
// The bound scales the fixed-point force accumulators of block_eval_kernel.
__global__ void __launch_bounds__(256) pair_prep_kernel(GridDesc g, double* __restrict__ cellrec,
                                                        unsigned long long* fmax_bits) {
  const int n = g.n[0];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  double bound = 0.0, vbound = 0.0;
  if (k < n) {
    const int k1 = (k + 1 < n) ? k + 1 : (g.periodic[0] ? 0 : k);
    const double v0 = g.rec[2 * (long)k], d0 = g.rec[2 * (long)k + 1];
    const double v1 = g.rec[2 * (long)k1], d1 = g.rec[2 * (long)k1 + 1];
    if (k + 1 < n) {
      cellrec[4 * (long)k + 0] = v0;
      cellrec[4 * (long)k + 1] = d0;
      cellrec[4 * (long)k + 2] = v1;
      cellrec[4 * (long)k + 3] = d1;
    }
    bound = 1.5 * fabs(v1 - v0) * g.inv_dx[0] + fmax(fabs(d0), fabs(d1));
    // |V| inside the cell: the Hermite value weights sum to 1, the slope weights stay below 4/27 each
    vbound = fmax(fabs(v0), fabs(v1)) + 0.3 * g.dx[0] * (fabs(d0) + fabs(d1));
  }
  for (int o = 16; o > 0; o >>= 1) {
    bound = fmax(bound, __shfl_down_sync(0xffffffffu, bound, o));
    vbound = fmax(vbound, __shfl_down_sync(0xffffffffu, vbound, o));
  }
  // non-negative doubles order like their bit patterns
  if ((threadIdx.x & 31) == 0 && bound > 0.0) atomicMax(fmax_bits, (unsigned long long)__double_as_longlong(bound));
  if ((threadIdx.x & 31) == 0 && vbound > 0.0) atomicMax(fmax_bits + 1, (unsigned long long)__double_as_longlong(vbound));
}

__device__ __forceinline__ double shift_of(int code, int d, const CellGrid& cg) {
  return (code & (1 << (2 * d))) ? cg.box[d] : ((code & (2 << (2 * d))) ? -cg.box[d] : 0.0);
}

// One candidate pair, exactly: separation as the oracle forms it ((x_i - x_j) - image shift, squares
// summed in x, y, z order), exact cutoff, bias at r, two hill proposals.  Returns false when the
// pair is outside the cutoff; p = force on i (the force on j is -p).
__device__ __forceinline__ bool pair_exact(const PairCtx& c, const AtomRec& ri, const AtomRec& rj, int code, double& e,
                                           double& px, double& py, double& pz) {
  double dx = __dsub_rn(ri.x, rj.x), dy = __dsub_rn(ri.y, rj.y), dz = __dsub_rn(ri.z, rj.z);
  if (code) {  // only partners reached across the box boundary; subtracting a zero shift changes nothing
    dx = __dsub_rn(dx, shift_of(code, 0, c.cg));
    dy = __dsub_rn(dy, shift_of(code, 1, c.cg));
    dz = __dsub_rn(dz, shift_of(code, 2, c.cg));
  }
  const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  if (!(d2 < c.pp.rc2)) return false;
  // lammps/fix_edm_pair.cpp:177 loops over LOCAL atoms i only: a pair of two ghosts is not in the list; a pair with
  // one ghost is, once (from its local atom): no force on the ghost (:223), one hill proposal instead of two (:233)
  const bool gi = (ri.tag & 0xffffffffLL) >= c.pp.nlocal, gj = (rj.tag & 0xffffffffLL) >= c.pp.nlocal;
  if (gi && gj) return false;
  const double rinv = rsqrt(d2);
  const double r = d2 * rinv;  // within 2 ulp of sqrt(d2): enough for V(r); hills take the exact root
  double force;
  e += pair_eval_fast(c.g, c.cellrec, c.pp.lean != 0, r, force);
  const double s = rinv * force;
  px = dx * s;
  py = dy * s;
  pz = dz * s;
  if (gi || gj) atomicAdd(&c.st->n_pairs_ghost, 1ULL);  // rare (halo pairs only)
  if (c.pp.do_hills) {
    const unsigned long long oi = (unsigned long long)(ri.tag & 0xffffffffLL), oj = (unsigned long long)(rj.tag & 0xffffffffLL);
    const unsigned long long lo = oi < oj ? oi : oj, hi = oi < oj ? oj : oi;
    propose_hills_d2(c.pp, lo * (unsigned long long)c.pp.natoms + hi, d2, c.st, c.acc, (gi || gj) ? 1 : 2);
  }
  return true;
}

// Segmented inclusive scan over runs of equal keys (contiguous lanes); afterwards the last lane of
// each run holds the run total.  Returns true on those lanes.
template <typename T> __device__ __forceinline__ bool run_totals(int key, T& px, T& py, T& pz) {
  const int lane = threadIdx.x & 31;
  const int kprev = __shfl_up_sync(0xffffffffu, key, 1);
  int head = (lane == 0 || kprev != key) ? 1 : 0;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int uh = __shfl_up_sync(0xffffffffu, head, o);
    const T ux = __shfl_up_sync(0xffffffffu, px, o);
    const T uy = __shfl_up_sync(0xffffffffu, py, o);
    const T uz = __shfl_up_sync(0xffffffffu, pz, o);
    if (lane >= o && !head) {
      px += ux;
      py += uy;
      pz += uz;
      head = uh;
    }
  }
  const int knext = __shfl_down_sync(0xffffffffu, key, 1);
  return key >= 0 && (lane == 31 || knext != key);
}


// ---- direct search: the last resort -----------------------------------------------------------------
//
// One thread per home atom, half shell of 13 forward cells + own cell straight from global memory,
// pair_exact on every pair inside the fp32 prefilter, all forces as fp64 REDs.  No capacity of any
// kind, so it takes the step whatever the density looks like; it runs (returns at once otherwise)
// only when the block search below gave up, and is several times slower than it.
__global__ void __launch_bounds__(128) pair_direct_kernel(const __grid_constant__ PairCtx c,
                                                          double* __restrict__ partial) {
  __shared__ double red[33];
  if (*c.fallback == 0) {
    if (threadIdx.x == 0) partial[blockIdx.x] = 0.0;
    return;
  }
  const CellGrid& cg = c.cg;
  const PairParams& pp = c.pp;
  double e = 0.0;
  unsigned long long npairs = 0;
  const long stride = (long)gridDim.x * blockDim.x;
  for (long a = (long)blockIdx.x * blockDim.x + threadIdx.x; a < pp.natoms; a += stride) {
    const AtomRec ri = c.arec[a];
    const int ti = __float_as_int(c.xs32[a].w);
    if (pp.use_types && ti != pp.itype && ti != pp.jtype) continue;
    const int cx = cell_coord(ri.x - cg.lo[0], cg.cs[0], cg.nc[0]);
    const int cy = cell_coord(ri.y - cg.lo[1], cg.cs[1], cg.nc[1]);
    const int cz = cell_coord(ri.z - cg.lo[2], cg.cs[2], cg.nc[2]);
    const bool ghost_i = (ri.tag & 0xffffffffLL) >= pp.nlocal;
    double fx = 0.0, fy = 0.0, fz = 0.0;
    for (int nb = 0; nb < 14; nb++) {
      // nb 0 = own cell; 1..13 = offsets with (dz,dy,dx) lexicographically positive
      const int t = nb + 13;  // 13..26 in the 3x3x3 enumeration, 13 = centre
      int qx = cx + t % 3 - 1, qy = cy + (t / 3) % 3 - 1, qz = cz + t / 9 - 1, code = 0;
      if (qx >= cg.nc[0]) { qx -= cg.nc[0]; code |= 1; } else if (qx < 0) { qx += cg.nc[0]; code |= 2; }
      if (qy >= cg.nc[1]) { qy -= cg.nc[1]; code |= 4; } else if (qy < 0) { qy += cg.nc[1]; code |= 8; }
      if (qz >= cg.nc[2]) { qz -= cg.nc[2]; code |= 16; } else if (qz < 0) { qz += cg.nc[2]; code |= 32; }
      // beyond a non-periodic end of the cell grid there is nothing (ghost atoms already sit inside it)
      if (((code & 3) && !cg.periodic[0]) || ((code & 12) && !cg.periodic[1]) || ((code & 48) && !cg.periodic[2])) continue;
      const int q = (qz * cg.nc[1] + qy) * cg.nc[0] + qx;
      long jlo = c.start[q];
      const long jhi = c.start[q + 1];
      if (nb == 0) jlo = a + 1;
      for (long j = jlo; j < jhi; j++) {
        if (pp.use_types) {
          const int tj = __float_as_int(c.xs32[j].w);
          if ((ti == pp.itype) ? (tj != pp.jtype) : (tj != pp.itype)) continue;
        }
        const AtomRec rj = c.arec[j];
        double px, py, pz;
        if (!pair_exact(c, ri, rj, code, e, px, py, pz)) continue;
        npairs++;
        fx += px;
        fy += py;
        fz += pz;
        if ((rj.tag & 0xffffffffLL) < pp.nlocal) {
          atomicAdd(&c.f[3 * rj.tag + 0], -px);
          atomicAdd(&c.f[3 * rj.tag + 1], -py);
          atomicAdd(&c.f[3 * rj.tag + 2], -pz);
        }
      }
    }
    if (!ghost_i) {
      atomicAdd(&c.f[3 * ri.tag + 0], fx);
      atomicAdd(&c.f[3 * ri.tag + 1], fy);
      atomicAdd(&c.f[3 * ri.tag + 2], fz);
    }
  }
  double tot = block_sum(e, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
  for (int o = 16; o > 0; o >>= 1) npairs += __shfl_down_sync(0xffffffffu, npairs, o);
  if ((threadIdx.x & 31) == 0 && npairs) atomicAdd(&c.st->n_pairs, npairs);
}

constexpr int kChunk = 256;

// ---- block search: a CTA owns a brick of home cells ---------------------------------------------
//
// A pair evaluation that gathers both atoms from global memory and returns both forces as REDs is
// bound by scattered traffic, not arithmetic: per pair two position gathers, a 32 B grid gather and
// three fp64 REDs, all through L1/L2 (ncu on the previous list-based kernel: L1 84 %, L2 68 % busy,
// 84 M RED requests per step).  Here a CTA owns a brick of bd[0] x bd[1] x bd[2] home
// cells and keeps everything those cells touch — the brick plus its forward half-shell halo, the
// REGION — in shared memory:
//   block_find_kernel  region positions as fp32 relative to the region corner.  A warp takes a home
//                      cell: each lane keeps up to 6 partner atoms in registers (the 14 neighbour
//                      cells are 5 contiguous runs of region indices), then for each home atom one
//                      broadcast LDS.128 and 6 distance tests per lane.  Accepted pairs go out as
//                      4 B items (li << 16 | lj, region indices) into chunks inside the block's own
//                      slice of the candidate buffer.
//   block_eval_kernel  region atoms as fp64 records + fp64 force accumulators in shared memory.
//                      One lane per candidate: both atoms from shared memory, pair_exact, partner
//                      force into the accumulator (shared-memory CAS add), home force through the
//                      segmented scan; one RED triple per region atom when the block is done —
//                      about 0.4 REDs per pair instead of 3.3.
// Both kernels number the region atoms the same way (region cells x fastest, slot order inside a
// cell), from the same start[] array.
constexpr int kBlkCap = 1920;    // region atoms: 56 B each in block_eval_kernel -> two CTAs per SM
constexpr int kRegCells = 256;   // region cells
constexpr int kFindThreads = 256;
constexpr int kEvalThreads = 512;
constexpr unsigned kPad = 0xffffffffu;

struct BlockGeom {
  int bd[3];       // home cells per brick side
  int nb[3];       // bricks per dimension
  unsigned capb;   // candidate items per brick (multiple of kChunk)
};

struct Region {
  int cstart[kRegCells + 1];  // region index of the first atom of each region cell
  int cslot[kRegCells];       // its slot
  unsigned char ccode[kRegCells];
  int hn[3], rd[3];           // home cells / region cells per side of this brick
  int nrc, nreg;
};

// false (for the whole CTA) when the region holds more than kBlkCap atoms
__device__ __forceinline__ bool region_setup(const CellGrid& cg, const BlockGeom& bg, const int* __restrict__ start,
                                             Region& R) {
  int b = blockIdx.x;
  const int bx = b % bg.nb[0];
  b /= bg.nb[0];
  const int by = b % bg.nb[1], bz = b / bg.nb[1];
  const int h0[3] = {bx * bg.bd[0], by * bg.bd[1], bz * bg.bd[2]};
  int hn[3], rd[3];
#pragma unroll
  for (int d = 0; d < 3; d++) {
    hn[d] = min(bg.bd[d], cg.nc[d] - h0[d]);
    rd[d] = hn[d] + (d < 2 ? 2 : 1);
  }
  const int nrc = rd[0] * rd[1] * rd[2];
  if (threadIdx.x == 0) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
      R.hn[d] = hn[d];
      R.rd[d] = rd[d];
    }
    R.nrc = nrc;
  }
  for (int rc = threadIdx.x; rc < nrc; rc += blockDim.x) {
    const int rx = rc % rd[0], ry = (rc / rd[0]) % rd[1], rz = rc / (rd[0] * rd[1]);
    int qx = h0[0] - 1 + rx, qy = h0[1] - 1 + ry, qz = h0[2] + rz, code = 0;
    if (qx >= cg.nc[0]) { qx -= cg.nc[0]; code |= 1; } else if (qx < 0) { qx += cg.nc[0]; code |= 2; }
    if (qy >= cg.nc[1]) { qy -= cg.nc[1]; code |= 4; } else if (qy < 0) { qy += cg.nc[1]; code |= 8; }
    if (qz >= cg.nc[2]) { qz -= cg.nc[2]; code |= 16; }
    // beyond a non-periodic end of the cell grid: an empty cell (ghost atoms already sit inside the grid)
    const bool beyond = ((code & 3) && !cg.periodic[0]) || ((code & 12) && !cg.periodic[1]) || ((code & 48) && !cg.periodic[2]);
    const int q = (qz * cg.nc[1] + qy) * cg.nc[0] + qx;
    const int s0 = start[q];
    R.cslot[rc] = s0;
    R.cstart[rc] = beyond ? 0 : start[q + 1] - s0;  // count for now
    R.ccode[rc] = (unsigned char)code;
  }
  __syncthreads();
  if (threadIdx.x < 32) {  // exclusive scan of up to 256 counts, 8 per lane
    const int lane = threadIdx.x, base = lane * 8;
    int v[8], s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      v[k] = (base + k < nrc) ? R.cstart[base + k] : 0;
      s += v[k];
    }
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += u;
    }
    __syncwarp();
    int run = incl - s;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (base + k < nrc) R.cstart[base + k] = run;
      run += v[k];
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (lane == 0) {
      R.cstart[nrc] = total;
      R.nreg = total;
    }
  }
  __syncthreads();
  return R.nreg <= kBlkCap;
}

// region cell holding region index l
__device__ __forceinline__ int region_cell_of(const Region& R, int l) {
  int lo = 0, hi = R.nrc;  // cstart[lo] <= l < cstart[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (R.cstart[mid] <= l) lo = mid; else hi = mid;
  }
  return lo;
}

// Pads the rest of the open chunk and takes the next one of the brick's slice.  When the slice is
// exhausted the step falls back to the generic search, so later items may land anywhere inside it.
__device__ __noinline__ unsigned* next_chunk(unsigned* chunk, int used, unsigned* slice, int* nchunks, int maxchunks,
                                             bool& dropped) {
  const int lane = threadIdx.x & 31;
  for (int a = used + lane; a < kChunk; a += 32) chunk[a] = kPad;
  int idx = 0;
  if (lane == 0) idx = atomicAdd(nchunks, 1);
  idx = __shfl_sync(0xffffffffu, idx, 0);
  if (idx >= maxchunks) {
    dropped = true;
    return slice;
  }
  return slice + (size_t)idx * kChunk;
}

struct FindSmem {
  Region R;
  float4 p[kBlkCap];  // region-relative fp32 position, type in .w
  int nchunks;
  int next_home;  // next home cell a warp of the CTA takes
};

template <bool TYPES>
__global__ void __launch_bounds__(kFindThreads, 4) block_find_kernel(const __grid_constant__ PairCtx c,
                                                                  const __grid_constant__ BlockGeom bg,
                                                                  unsigned* __restrict__ items,
                                                                  int* __restrict__ nchunks_out) {
  __shared__ FindSmem S;
  Region& R = S.R;
  const CellGrid& cg = c.cg;
  const PairParams& pp = c.pp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    S.nchunks = 0;
    S.next_home = kFindThreads / 32;
  }
  if (!region_setup(cg, bg, c.start, R)) {
    if (threadIdx.x == 0) {
      *c.fallback = 1;
      nchunks_out[blockIdx.x] = 0;
    }
    return;
  }
  const float csx = (float)cg.cs[0], csy = (float)cg.cs[1], csz = (float)cg.cs[2];
  const int rd0 = R.rd[0], rd1 = R.rd[1];
  for (int l = threadIdx.x; l < R.nreg; l += kFindThreads) {
    const int rc = region_cell_of(R, l);
    float4 v = c.xs32[R.cslot[rc] + (l - R.cstart[rc])];
    v.x += (float)(rc % rd0) * csx;
    v.y += (float)((rc / rd0) % rd1) * csy;
    v.z += (float)(rc / (rd0 * rd1)) * csz;
    S.p[l] = v;
  }
  __syncthreads();

  const float rc2m = pp.rc2m;
  const int maxchunks = (int)(bg.capb / kChunk);
  unsigned* const slice = items + (size_t)blockIdx.x * bg.capb;
  unsigned* chunk = slice;
  int used = kChunk;  // forces an allocation on first use
  bool dropped = false;
  const int nhome = R.hn[0] * R.hn[1] * R.hn[2];

  // Home cells are handed out dynamically after the first round: 36 cells over 8 warps leave half the warps a
  // cell short, and occupancies vary (12.5 +- 3.5 atoms); the CTA's final barrier was 15 % of the stall samples.
  // Which warp searched which cell changes only the order of the candidate items, which nothing depends on.
  auto next_home = [&]() {
    int hn = 0;
    if (lane == 0) hn = atomicAdd(&S.next_home, 1);
    return __shfl_sync(0xffffffffu, hn, 0);
  };
  for (int h = warp; h < nhome; h = next_home()) {
    const int hx = h % R.hn[0], hy = (h / R.hn[0]) % R.hn[1], hz = h / (R.hn[0] * R.hn[1]);
    const int rx = hx + 1, ry = hy + 1, rz = hz;
    const int c00 = (rz * rd1 + ry) * rd0 + rx;
    const int ilo = R.cstart[c00], ihi = R.cstart[c00 + 1];
    if (ilo == ihi) continue;
    // the 14 cells of the half shell as 5 runs of region indices
    const int c1 = c00 + rd0, c2 = c00 + rd0 * rd1 - rd0, c3 = c00 + rd0 * rd1, c4 = c3 + rd0;
    const int s0 = ilo, s1 = R.cstart[c1 - 1], s2 = R.cstart[c2 - 1], s3 = R.cstart[c3 - 1], s4 = R.cstart[c4 - 1];
    const int o1 = R.cstart[c00 + 2] - s0;
    const int o2 = o1 + R.cstart[c1 + 2] - s1;
    const int o3 = o2 + R.cstart[c2 + 2] - s2;
    const int o4 = o3 + R.cstart[c3 + 2] - s3;
    const int nj = o4 + R.cstart[c4 + 2] - s4;

    for (int p0 = 0; p0 < nj; p0 += 192) {
      // partner u of this lane: position and ljc = region index << 2 | type class (bit 0: is itype,
      // bit 1: is jtype).  Every partner outside the own cell has a larger region index than any
      // home atom, so "li < lj" is exactly the own-cell rule (count each pair of cell mates once).
      float jx[6], jy[6], jz[6];
      int ljc[6];
#pragma unroll
      for (int u = 0; u < 6; u++) {
        const int k = p0 + 32 * u + lane;
        const int l = k < o1 ? s0 + k : (k < o2 ? s1 + (k - o1) : (k < o3 ? s2 + (k - o2) : (k < o4 ? s3 + (k - o3) : s4 + (k - o4))));
        ljc[u] = 0;
        jx[u] = 1e30f;
        jy[u] = 0.f;
        jz[u] = 0.f;
        if (k < nj) {
          const float4 v = S.p[l];
          jx[u] = v.x;
          jy[u] = v.y;
          jz[u] = v.z;
          int cls = 3;
          if (TYPES) {
            const int tj = __float_as_int(v.w);
            cls = (tj == pp.itype ? 1 : 0) | (tj == pp.jtype ? 2 : 0);
          }
          ljc[u] = (l << 2) | cls;
        }
      }
      for (int li = ilo; li < ihi; li++) {
        const float4 pi = S.p[li];
        int want = 3;
        if (TYPES) {  // an itype atom pairs with jtype partners and the other way round
          const int ti = __float_as_int(pi.w);
          if (ti != pp.itype && ti != pp.jtype) continue;
          want = (ti == pp.itype) ? 2 : 1;
        }
        const int lic = (li << 2) | 3;
        // each lane keeps its own accept flags; one warp scan of the per-lane counts places the items
        // (their order inside a home atom's run does not matter to the evaluation)
        unsigned flags = 0;
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < 6; u++) {
          const float dx = pi.x - jx[u], dy = pi.y - jy[u], dz = pi.z - jz[u];
          const float d2 = dx * dx + dy * dy + dz * dz;
          bool take = (d2 < rc2m) && (lic < ljc[u]);
          if (TYPES) take = take && (ljc[u] & want);
          flags |= (take ? 1u : 0u) << u;
          cnt += take ? 1 : 0;
        }
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += v;
        }
        const int tot = __shfl_sync(0xffffffffu, inc, 31);
        if (tot == 0) continue;
        if (used + tot > kChunk) {  // at most 192 per home atom and pass, so a fresh chunk always fits
          chunk = next_chunk(chunk, used, slice, &S.nchunks, maxchunks, dropped);
          used = 0;
        }
        const unsigned hi16 = (unsigned)li << 16;
        int pos = used + inc - cnt;
#pragma unroll
        for (int u = 0; u < 6; u++)
          if ((flags >> u) & 1u) chunk[pos++] = hi16 | (unsigned)(ljc[u] >> 2);
        used += tot;
      }
    }
  }
  for (int a = used + lane; a < kChunk; a += 32) chunk[a] = kPad;
  if (dropped && lane == 0) *c.fallback = 1;
  __syncthreads();
  if (threadIdx.x == 0) nchunks_out[blockIdx.x] = min(S.nchunks, maxchunks);
}

struct EvalSmem {
  Region R;
  double red[33];
  double2 pa[kBlkCap];   // {x, y}
  double2 pb[kBlkCap];   // {z, bits: original index | image-shift code << 32}
  // fixed-point force accumulators, 64 bits as (lo, hi) words with native 32-bit shared atomics
  unsigned lo[3][kBlkCap];
  int hi[3][kBlkCap];
};

// acc += q (two's complement, 64 bits split in two words): add the low word, carry into the high one
__device__ __forceinline__ void fx_add(unsigned* lo, int* hi, long long q) {
  const unsigned ql = (unsigned)q;
  int qh = (int)(q >> 32);
  const unsigned old = atomicAdd(lo, ql);
  qh += (old + ql < old) ? 1 : 0;
  if (qh) atomicAdd(hi, qh);
}

__global__ void __launch_bounds__(kEvalThreads, 2) block_eval_kernel(const __grid_constant__ PairCtx c,
                                                                     const __grid_constant__ BlockGeom bg,
                                                                     const unsigned* __restrict__ items,
                                                                     const int* __restrict__ nchunks_in,
                                                                     double* __restrict__ partial) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  EvalSmem& S = *reinterpret_cast<EvalSmem*>(smem_raw);
  Region& R = S.R;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (*c.fallback != 0) {
    if (threadIdx.x == 0) partial[blockIdx.x] = 0.0;
    return;
  }
  region_setup(c.cg, bg, c.start, R);  // fits: block_find_kernel checked the same numbers
  for (int l = threadIdx.x; l < R.nreg; l += kEvalThreads) {
    const int rc = region_cell_of(R, l);
    const AtomRec r = c.arec[R.cslot[rc] + (l - R.cstart[rc])];
    S.pa[l] = make_double2(r.x, r.y);
    S.pb[l] = make_double2(r.z, __longlong_as_double(r.tag | ((long long)R.ccode[rc] << 32)));
#pragma unroll
    for (int k = 0; k < 3; k++) {
      S.lo[k][l] = 0u;
      S.hi[k][l] = 0;
    }
  }
  // Fixed point: a pair force component is at most fmax < 2^ex in magnitude and a region atom has
  // fewer than kBlkCap < 2^11 partners, so with scale 2^(51-ex) no sum reaches 2^62; one unit is
  // 2^-51 of the largest possible pair force, the resolution fp64 itself has there.  Integer sums
  // do not depend on the order the atomics resolve in.
  const double fmax_grid = __longlong_as_double((long long)c.fmax_bits[0]);
  const double scale = fmax_grid > 0.0 ? scalbn(1.0, 51 - (ilogb(fmax_grid) + 1)) : 1.0;
  // The energy too: |V| < 2^ev, so with scale 2^(44-ev) a lane's integer sum (fewer than 2^18 pairs) stays
  // below 2^62; the CTA adds the lanes in 128 bits.  The sum does not depend on which lane met which pair
  // (the order chunks were handed out in); one unit is 2^-44 of the largest possible pair energy.
  const double vmax_grid = __longlong_as_double((long long)c.fmax_bits[1]);
  const double escale = vmax_grid > 0.0 ? scalbn(1.0, 44 - (ilogb(vmax_grid) + 1)) : 1.0;
  __syncthreads();

  const unsigned* const slice = items + (size_t)blockIdx.x * bg.capb;
  const int nbatch = nchunks_in[blockIdx.x] * (kChunk / 32);
  long long eq = 0;
  unsigned npairs = 0;
  unsigned it_next = warp < nbatch ? slice[(size_t)warp * 32 + lane] : kPad;
  for (int bt = warp; bt < nbatch; bt += kEvalThreads / 32) {
    const unsigned it = it_next;
    const int bn = bt + kEvalThreads / 32;
    it_next = bn < nbatch ? slice[(size_t)bn * 32 + lane] : kPad;  // in flight while this batch is evaluated
    if (__ballot_sync(0xffffffffu, it != kPad) == 0u) continue;
    int li = -1;
    double px = 0.0, py = 0.0, pz = 0.0;
    if (it != kPad) {
      li = (int)(it >> 16);
      const int lj = (int)(it & 0xffffu);
      const double2 ia = S.pa[li], ib = S.pb[li], ja = S.pa[lj], jb = S.pb[lj];
      AtomRec ri, rj;
      ri.x = ia.x; ri.y = ia.y; ri.z = ib.x; ri.tag = __double_as_longlong(ib.y);
      rj.x = ja.x; rj.y = ja.y; rj.z = jb.x; rj.tag = __double_as_longlong(jb.y);
      double e = 0.0;
      if (pair_exact(c, ri, rj, (int)(rj.tag >> 32), e, px, py, pz)) {
        npairs++;
        eq += __double2ll_rn(e * escale);
      }
    }
    // the partner takes -q, the home atom the run's sum of the same integers: the bias force on
    // the block's atoms adds up to exactly zero
    const long long qx = __double2ll_rn(px * scale), qy = __double2ll_rn(py * scale), qz = __double2ll_rn(pz * scale);
    if (qx | qy | qz) {
      const int lj = (int)(it & 0xffffu);
      fx_add(&S.lo[0][lj], &S.hi[0][lj], -qx);
      fx_add(&S.lo[1][lj], &S.hi[1][lj], -qy);
      fx_add(&S.lo[2][lj], &S.hi[2][lj], -qz);
    }
    // (a warp-reduce per run mask, REDUX on 21-bit limbs, was measured 30 % slower than this scan:
    // with several runs in a batch the masks differ per lane and the reduce serialises per mask)
    long long sx = qx, sy = qy, sz = qz;
    if (run_totals(li, sx, sy, sz)) {
      fx_add(&S.lo[0][li], &S.hi[0][li], sx);
      fx_add(&S.lo[1][li], &S.hi[1][li], sy);
      fx_add(&S.lo[2][li], &S.hi[2][li], sz);
    }
  }
  __syncthreads();
  const double inv_scale = 1.0 / scale;
  for (int l = threadIdx.x; l < R.nreg; l += kEvalThreads) {
    long long q[3];
#pragma unroll
    for (int k = 0; k < 3; k++) q[k] = ((long long)S.hi[k][l] << 32) + (long long)S.lo[k][l];
    const long long tag = __double_as_longlong(S.pb[l].y) & 0xffffffffLL;
    if ((q[0] | q[1] | q[2]) && tag < c.pp.nlocal) {  // a ghost takes no force here: its owner computes the pair too
      const long long o = 3 * tag;
      atomicAdd(&c.f[o + 0], (double)q[0] * inv_scale);
      atomicAdd(&c.f[o + 1], (double)q[1] * inv_scale);
      atomicAdd(&c.f[o + 2], (double)q[2] * inv_scale);
    }
  }
  // integer sum over the CTA: exact, so any order will do
  __int128 wq = eq;
  for (int o = 16; o > 0; o >>= 1) {
    const long long hi = __shfl_down_sync(0xffffffffu, (long long)(wq >> 64), o);
    const unsigned long long lo = __shfl_down_sync(0xffffffffu, (unsigned long long)wq, o);
    wq += ((__int128)hi << 64) | (__int128)lo;
  }
  unsigned long long* redq = reinterpret_cast<unsigned long long*>(S.red);  // 16 warps x (lo, hi) in the 33-double scratch
  __syncthreads();
  if (lane == 0) {
    redq[2 * warp] = (unsigned long long)wq;
    redq[2 * warp + 1] = (unsigned long long)(wq >> 64);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __int128 t = 0;
    for (int i = 0; i < kEvalThreads / 32; i++) t += ((__int128)(long long)redq[2 * i + 1] << 64) | (__int128)redq[2 * i];
    partial[blockIdx.x] = (double)t / escale;
  }
  for (int o = 16; o > 0; o >>= 1) npairs += __shfl_down_sync(0xffffffffu, npairs, o);
  if (lane == 0 && npairs) atomicAdd(&c.st->n_pairs, (unsigned long long)npairs);
}

__global__ void pair_reset_kernel(BiasDev* st, int* fallback, int fallback_init, unsigned long long* fmax_bits) {
  st->n_pairs = 0;
  st->n_pairs_ghost = 0;
  *fallback = fallback_init;
  fmax_bits[0] = 0ull;
  fmax_bits[1] = 0ull;
}

// ---- neighbour-list form, pair-parallel (lammps/fix_edm_pair.cpp:177-240) ------------------------------
//
// The list a caller hands over (LAMMPS rebuilds it every ten steps or so) stays on the device between
// calls, flattened: row_of[k] is the list row of entry k.  One thread per listed pair: consecutive
// threads share the row atom i, so its force is a segmented warp scan and one RED per run; the
// partner's force goes out as fp64 REDs (x, 24 B per atom, lives in L2).
__global__ void list_rows_kernel(long inum, const long* __restrict__ first, int* __restrict__ row_of) {
  const long ii = (long)blockIdx.x * blockDim.y + threadIdx.y;  // a warp per row
  if (ii >= inum) return;
  for (long k = first[ii] + threadIdx.x; k < first[ii + 1]; k += 32) row_of[k] = (int)ii;
}

__global__ void __launch_bounds__(256) pair_list_flat_kernel(GridDesc g, const double* __restrict__ cellrec,
                                                             PairParams pp, long nlocal, long nlisted,
                                                             const int* __restrict__ ilist,
                                                             const int* __restrict__ row_of,
                                                             const int* __restrict__ jlist,
                                                             const double* __restrict__ x,
                                                             const int* __restrict__ type,
                                                             const double* __restrict__ runiform,
                                                             double* __restrict__ f, double* __restrict__ partial,
                                                             BiasDev* st, HillAccepted* acc,
                                                             unsigned long long* ncalls) {
  __shared__ double red[33];
  double e = 0.0;
  unsigned long long npairs = 0, calls = 0;
  const long stride = (long)gridDim.x * blockDim.x;
  const long nround = (nlisted + 31) / 32 * 32;  // whole warps stay together for the scan
  for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < nround; k += stride) {
    int row = -1, i = 0, j = 0;
    bool on = k < nlisted;
    if (on) {
      row = row_of[k];
      i = ilist[row];
      j = jlist[k];
      if (pp.use_types) {  // lammps/fix_edm_pair.cpp:179-203
        const int it = type[i], jt = type[j];
        if (it == pp.itype) on = jt == pp.jtype;
        else if (it == pp.jtype) on = jt == pp.itype;
        else on = false;
      }
    }
    double px = 0.0, py = 0.0, pz = 0.0;
    if (on) {
      const double dx = __dsub_rn(x[3 * (long)i + 0], x[3 * (long)j + 0]);
      const double dy = __dsub_rn(x[3 * (long)i + 1], x[3 * (long)j + 1]);
      const double dz = __dsub_rn(x[3 * (long)i + 2], x[3 * (long)j + 2]);
      const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      const double rinv = rsqrt(d2);
      const double r = d2 * rinv;  // within 2 ulp of sqrt(d2): enough for V(r); hills take the exact root
      double force;
      e += pair_eval_fast(g, cellrec, pp.lean != 0, r, force);
      npairs++;
      const double sc = rinv * force;
      px = dx * sc;
      py = dy * sc;
      pz = dz * sc;
      const bool jlocal = j < nlocal;
      if (jlocal) {  // fix_edm_pair.cpp:223
        atomicAdd(&f[3 * (long)j + 0], -px);
        atomicAdd(&f[3 * (long)j + 1], -py);
        atomicAdd(&f[3 * (long)j + 2], -pz);
      }
      if (pp.do_hills) {  // fix_edm_pair.cpp:230-236
        const int nprop = jlocal ? 2 : 1;
        calls += nprop;
        for (int which = 0; which < nprop; which++) {
          const unsigned long long kk = 2ULL * (unsigned long long)k + which;
          const double u = runiform ? runiform[kk] : uniform_from_key(pp.key, kk);
          if (pp.accept_all || u < pp.thresh) {
            const int slot = atomicAdd(&st->n_accepted, 1);
            if (slot < pp.acc_cap) {
              acc[slot].key = kk;
              acc[slot].x[0] = sqrt(d2);
              acc[slot].x[1] = 0.0;
              acc[slot].x[2] = 0.0;
            } else {
              st->accepted_overflow = 1;
            }
          }
        }
      }
    }
    if (run_totals(on ? row : -1 - (int)(threadIdx.x & 31), px, py, pz)) {
      atomicAdd(&f[3 * (long)i + 0], px);
      atomicAdd(&f[3 * (long)i + 1], py);
      atomicAdd(&f[3 * (long)i + 2], pz);
    }
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

__global__ void add_forces_kernel(long n, const double* __restrict__ d, double* __restrict__ f) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) f[i] += d[i];
}

__global__ void reset_pairs_kernel(BiasDev* st, unsigned long long* ncalls) {
  st->n_pairs = 0;
  st->n_pairs_ghost = 0;
  if (ncalls) *ncalls = 0;
}

// closes a pair step: energy = sum of the per-CTA partials; reports a fallback to the host-mapped flag
__global__ void sum_partials2_kernel(int n, const double* __restrict__ partial, double* out, const int* fallback,
                                     volatile int* host_flag) {
  __shared__ double red[33];
  double e = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) e += partial[i];
  double tot = block_sum(e, red);
  if (threadIdx.x == 0) {
    out[0] = tot;
    if (host_flag && fallback && *fallback) {
      *host_flag = 1;
      __threadfence_system();
    }
  }
}

// last kernel of a host-buffer step: the scalars the host wants, into host-mapped memory
__global__ void step_report_kernel(const BiasDev* st, const double* energy, const unsigned long long* ncalls,
                                   HostReport* rep) {
  rep->energy = energy ? *energy : 0.0;
  rep->n_pairs = st->n_pairs;
  rep->n_calls = ncalls ? *ncalls : 2ULL * st->n_pairs - st->n_pairs_ghost;
  rep->backlog_full = st->backlog_full;
  rep->accepted_overflow = st->accepted_overflow;
  __threadfence_system();
}

}  // namespace edm

using namespace edm;

int edm_host_report_ensure(edm_bias* b) {
  if (b->h_pair_flags) return EDM_OK;
  static_assert(sizeof(HostReport) <= 64, "HostReport outgrew its allocation");
  EDM_CUDA(cudaHostAlloc((void**)&b->h_pair_flags, 64, cudaHostAllocMapped));
  memset((void*)b->h_pair_flags, 0, 64);
  EDM_CUDA(cudaHostGetDevicePointer((void**)&b->d_pair_flags, (void*)b->h_pair_flags, 0));
  return EDM_OK;
}

// streams and events of the host-buffer pair steps; orders both streams after whatever the caller
// enqueued on the default stream before this call (uploads, _dev entry points), without stalling the host
static int pair_host_streams(edm_bias* b) {
  if (!b->st_main) {
    EDM_CUDA(cudaStreamCreateWithFlags(&b->st_main, cudaStreamNonBlocking));
    EDM_CUDA(cudaStreamCreateWithFlags(&b->st_copy, cudaStreamNonBlocking));
    EDM_CUDA(cudaEventCreateWithFlags(&b->ev_f_up, cudaEventDisableTiming));
    EDM_CUDA(cudaEventCreateWithFlags(&b->ev_f_final, cudaEventDisableTiming));
  }
  if (!b->ev_prev) EDM_CUDA(cudaEventCreateWithFlags(&b->ev_prev, cudaEventDisableTiming));
  EDM_CUDA(cudaEventRecord(b->ev_prev, 0));
  EDM_CUDA(cudaStreamWaitEvent(b->st_main, b->ev_prev, 0));
  EDM_CUDA(cudaStreamWaitEvent(b->st_copy, b->ev_prev, 0));
  if (b->profiling && !b->ev_e2e[0])
    for (int i = 0; i < 5; i++) EDM_CUDA(cudaEventCreate(&b->ev_e2e[i]));
  return edm_host_report_ensure(b);
}

// after the stream that ran step_report_kernel has been synchronised
static int pair_host_result(edm_bias* b, edm_pair_result_t* result, int do_hills) {
  const HostReport* rep = reinterpret_cast<const HostReport*>(const_cast<const int*>(b->h_pair_flags));
  if (result) {
    result->energy = rep->energy;
    result->n_pairs = (long long)rep->n_pairs;
    result->n_calls = (long long)rep->n_calls;
  }
  if (do_hills) {
    if (rep->backlog_full) {
      set_error("The bias overflow buffer is full. Too many hills (lib/edm_bias.cpp:503-507)");
      return EDM_ERR_BACKLOG_FULL;
    }
    if (rep->accepted_overflow) {
      set_error("accepted-hill buffer exhausted");
      return EDM_ERR_CAPACITY;
    }
  }
  return EDM_OK;
}

static PairParams pair_params(const edm_bias* b, const int* type, int itype, int jtype, int do_hills, long long est,
                              uint64_t seed, uint64_t step, double cutoff, long natoms) {
  PairParams pp;
  pp.itype = itype;
  pp.jtype = jtype;
  pp.use_types = type != nullptr;
  pp.do_hills = do_hills;
  pp.accept_all = b->prm.hill_density < 0;
  {
    const GridDesc& g = b->bias->d;
    pp.lean = g.is_gauss && g.b_interp && g.b_deriv && !g.periodic[0] && !g.bper[0];
  }
  pp.thresh = pp.accept_all ? 2.0 : b->prm.hill_density / (double)(int)est;
  {
    double tb = ceil(pp.thresh * 4294967296.0);
    pp.thresh_bits = tb >= 4294967296.0 ? 4294967296ULL : (tb <= 0.0 ? 0ULL : (uint64_t)tb);
  }
  pp.key = uniform_key(seed, step);
  static const int dbg_env = getenv("EDM_DBG") ? atoi(getenv("EDM_DBG")) : 0;  // read once, not per step
  pp.dbg = dbg_env;
  pp.rc2 = cutoff * cutoff;
  pp.rc2m = (float)(pp.rc2 * 1.0001) + 1e-6f;
  pp.natoms = natoms;
  pp.nlocal = natoms;  // every atom local unless the caller says otherwise
  pp.acc_cap = b->accepted_cap;
  return pp;
}

// bricks of home cells for the block search: the largest home/region ratio whose region fits the
// shared-memory budget with a margin over the mean occupancy (10 % + 6 sigma of a Poisson count).
// `scale` inflates the assumed density: it grows after a step that had to fall back (non-uniform
// systems), which shrinks the bricks.
static bool choose_bricks(const CellGrid& cg, long natoms, double cutoff, double scale, BlockGeom& bg) {
  const double m = scale * (double)natoms / (double)cg.ncell;
  double best = 0.0;
  for (int bz = 1; bz <= 8 && bz <= cg.nc[2]; bz++)
    for (int by = 1; by <= 8 && by <= cg.nc[1]; by++)
      for (int bx = 1; bx <= 8 && bx <= cg.nc[0]; bx++) {
        const int rcells = (bx + 2) * (by + 2) * (bz + 1);
        if (rcells > kRegCells) continue;
        const double mean = rcells * m;
        if (1.1 * mean + 6.0 * sqrt(mean) > kBlkCap) continue;
        const double score = (double)(bx * by * bz) / rcells + 1e-6 * bx;
        if (score > best) {
          best = score;
          bg.bd[0] = bx;
          bg.bd[1] = by;
          bg.bd[2] = bz;
        }
      }
  if (best == 0.0) return false;
  for (int d = 0; d < 3; d++) bg.nb[d] = (cg.nc[d] + bg.bd[d] - 1) / bg.bd[d];
  const double vol = cg.box[0] * cg.box[1] * cg.box[2];
  const double per_cell = 0.5 * m * (scale * (double)natoms / vol) * (4.0 / 3.0) * M_PI * cutoff * cutoff * cutoff;
  const double expect = per_cell * bg.bd[0] * bg.bd[1] * bg.bd[2];
  unsigned long long cap = (unsigned long long)(1.5 * expect + 8.0 * sqrt(expect + 1.0)) + (kFindThreads / 32 + 1) * kChunk;
  cap = (cap + kChunk - 1) / kChunk * kChunk;
  const unsigned long long total = cap * (unsigned long long)bg.nb[0] * bg.nb[1] * bg.nb[2];
  if (cap > (1ull << 30) || total * sizeof(unsigned) > (24ull << 30)) return false;
  bg.capb = (unsigned)cap;
  return true;
}

// binning + pair kernels on device pointers; leaves the energy in *energy_dev
static int pair_cells_launch(edm_bias* b, long natoms, const double* x, double* f, const int* type, int itype,
                             int jtype, const double* box, double cutoff, int do_hills, long long est, uint64_t seed,
                             uint64_t step, double* energy_dev, cudaStream_t st, cudaEvent_t forces_ready = nullptr,
                             cudaEvent_t forces_done = nullptr, const edm_pair_domain_t* dom = nullptr) {
  // dom == NULL: a periodic box [0, box)^3 with every atom local.  Otherwise one rank's share of a larger system
  // (LAMMPS' picture): local atoms first, then ghosts, inside [lo, hi); dimensions flagged periodic are wrapped here,
  // the others end at lo / hi and whatever lies beyond has been handed over as ghost atoms.
  // forces_ready: waited for before the first kernel that touches f (the search before it only reads
  // x); forces_done: recorded once f is final (before any hill work the caller appends)
  EDM_REQUIRE(b->prm.dim == 1, "Pairwise distance must be 1 dimension in EDM input file");  // fix_edm_pair.cpp:52-53
  EDM_REQUIRE(natoms > 0 && natoms < 2000000000L, "bad atom count");
  // two proposals per pair; the pair count is not known before the search: ~100 partners per atom is generous
  // for a liquid at this cutoff, and a round that outgrows the buffer is reported, never truncated
  if (do_hills) EDM_TRY(edm_bias_size_accepted(b, 200.0 * (double)natoms, est));
  CellGrid cg;
  long long ncell = 1;
  for (int d = 0; d < 3; d++) {
    cg.lo[d] = dom ? dom->lo[d] : 0.0;
    cg.periodic[d] = dom ? (dom->periodic[d] != 0) : 1;
    cg.box[d] = dom ? dom->hi[d] - dom->lo[d] : box[d];
    EDM_REQUIRE(cg.box[d] > 0, "empty domain");
    cg.nc[d] = (int)floor(cg.box[d] / cutoff);
    if (cg.periodic[d])
      EDM_REQUIRE(cg.nc[d] >= 3, "box must hold at least 3 cutoffs per side for the half-shell cell search");
    else if (cg.nc[d] < 1)
      cg.nc[d] = 1;
    cg.cs[d] = cg.box[d] / cg.nc[d];
    ncell *= cg.nc[d];
  }
  EDM_REQUIRE(ncell < 2000000000LL, "too many cells");
  cg.ncell = (int)ncell;

  // feedback from earlier steps (device-written, host-mapped): a fallback shrinks the bricks
  EDM_TRY(edm_host_report_ensure(b));
  if (b->h_pair_flags[0]) {
    b->h_pair_flags[0] = 0;
    b->pair_fallbacks++;
    if (b->brick_scale < 64.0) b->brick_scale *= 1.5;
  }
  // EDM_PAIR_MODE=0 forces the direct search (A/B measurements, tests of the last resort)
  static int mode = -1;
  if (mode < 0) {
    const char* ev = getenv("EDM_PAIR_MODE");
    mode = ev ? atoi(ev) : 6;
  }
  BlockGeom bg;
  const bool bricks = (mode != 0) && choose_bricks(cg, natoms, cutoff, b->brick_scale, bg);
  const int nblocks = bricks ? bg.nb[0] * bg.nb[1] * bg.nb[2] : 0;
  for (int d = 0; d < 3; d++) b->brick_dims[d] = bricks ? bg.bd[d] : 0;
  const int dblocks = sm_count(b->device) * 8;  // direct search: grid-stride over the atoms

  // scratch: cell_of[n] order[n] | count[ncell+1] start[ncell+1] tiles | flags | nchunks[nblocks] |
  //          partial[nblocks + dblocks] | arec[n] | xs32[n]
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t n = (size_t)natoms, nc1 = (size_t)ncell + 1;
  const size_t off_cell = 0, off_order = up(off_cell + n * 4), off_count = up(off_order + n * 4);
  const size_t off_start = up(off_count + nc1 * 4), off_tiles = up(off_start + nc1 * 4);
  const size_t off_flags = up(off_tiles + ((size_t)ncell / kScanTile + 2) * 4);
  const size_t off_nch = up(off_flags + 64), off_part = up(off_nch + (size_t)nblocks * 4 + 4);
  const size_t off_rec = up(off_part + ((size_t)nblocks + dblocks) * 8);
  const size_t off_xs32 = up(off_rec + n * sizeof(AtomRec));
  const size_t off_crec = up(off_xs32 + n * sizeof(float4));
  const size_t total = off_crec + (size_t)b->bias->d.n[0] * 4 * sizeof(double);
  EDM_TRY(b->cells.reserve(total));
  char* base = b->cells.as<char>();
  int* cell_of = reinterpret_cast<int*>(base + off_cell);
  int* order = reinterpret_cast<int*>(base + off_order);
  int* count = reinterpret_cast<int*>(base + off_count);
  int* start = reinterpret_cast<int*>(base + off_start);
  int* tile_total = reinterpret_cast<int*>(base + off_tiles);
  int* fallback = reinterpret_cast<int*>(base + off_flags);
  int* nchunks = reinterpret_cast<int*>(base + off_nch);
  double* partial = reinterpret_cast<double*>(base + off_part);
  AtomRec* arec = reinterpret_cast<AtomRec*>(base + off_rec);
  float4* xs32 = reinterpret_cast<float4*>(base + off_xs32);
  double* cellrec = reinterpret_cast<double*>(base + off_crec);
  unsigned long long* fmax_bits = reinterpret_cast<unsigned long long*>(base + off_flags + 8);
  if (bricks) EDM_TRY(b->cand.reserve((size_t)nblocks * bg.capb * sizeof(unsigned)));

  EDM_CUDA(cudaMemsetAsync(count, 0, nc1 * 4, st));
  const unsigned nb = (unsigned)((natoms + 255) / 256);
  cell_count_kernel<<<nb, 256, 0, st>>>(natoms, x, cg, cell_of, count);
  const int ntiles = (cg.ncell + kScanTile - 1) / kScanTile;
  cell_scan_totals_kernel<<<ntiles, 256, 0, st>>>(cg.ncell, count, tile_total);
  cell_scan_kernel<<<ntiles, 256, 0, st>>>(cg.ncell, count, tile_total, start);
  cell_fill_kernel<<<nb, 256, 0, st>>>(natoms, cell_of, start, count, order);
  cell_rank_gather_kernel<<<nb, 256, 0, st>>>(natoms, cg, cell_of, start, order, x, type, arec, xs32);
  pair_reset_kernel<<<1, 1, 0, st>>>(b->d_state, fallback, bricks ? 0 : 1, fmax_bits);
  pair_prep_kernel<<<(b->bias->d.n[0] + 255) / 256, 256, 0, st>>>(b->bias->d, cellrec, fmax_bits);
  EDM_CUDA(cudaGetLastError());

  PairCtx ctx;
  ctx.g = b->bias->d;
  ctx.cg = cg;
  ctx.pp = pair_params(b, type, itype, jtype, do_hills, est, seed, step, cutoff, natoms);
  if (dom) {
    EDM_REQUIRE(dom->nlocal >= 0 && dom->nlocal <= natoms, "nlocal out of range");
    ctx.pp.nlocal = dom->nlocal;
  }
  ctx.start = start;
  ctx.arec = arec;
  ctx.xs32 = xs32;
  ctx.f = f;
  ctx.st = b->d_state;
  ctx.acc = b->d_accepted;
  ctx.fallback = fallback;
  ctx.cellrec = cellrec;
  ctx.fmax_bits = fmax_bits;

  if (b->profiling) EDM_CUDA(cudaEventRecord(b->ev_pair[0], st));
  int launched = 7;
  if (bricks) {
    if (!b->eval_attr_set) {  // per device: one process may drive several
      EDM_CUDA(cudaFuncSetAttribute(block_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(EvalSmem)));
      b->eval_attr_set = 1;
    }
    if (type)
      block_find_kernel<true><<<nblocks, kFindThreads, 0, st>>>(ctx, bg, b->cand.as<unsigned>(), nchunks);
    else
      block_find_kernel<false><<<nblocks, kFindThreads, 0, st>>>(ctx, bg, b->cand.as<unsigned>(), nchunks);
    if (b->profiling) EDM_CUDA(cudaEventRecord(b->ev_pair[2], st));
    if (forces_ready) EDM_CUDA(cudaStreamWaitEvent(st, forces_ready, 0));
    block_eval_kernel<<<nblocks, kEvalThreads, sizeof(EvalSmem), st>>>(ctx, bg, b->cand.as<unsigned>(), nchunks, partial);
    launched += 2;
  } else {
    if (b->profiling) EDM_CUDA(cudaEventRecord(b->ev_pair[2], st));
    if (forces_ready) EDM_CUDA(cudaStreamWaitEvent(st, forces_ready, 0));
  }
  // returns at once unless the fallback flag is up
  pair_direct_kernel<<<dblocks, 128, 0, st>>>(ctx, partial + nblocks);
  if (forces_done) EDM_CUDA(cudaEventRecord(forces_done, st));
  if (b->profiling) EDM_CUDA(cudaEventRecord(b->ev_pair[1], st));
  sum_partials2_kernel<<<1, 256, 0, st>>>(nblocks + dblocks, partial, energy_dev, fallback, bricks ? b->d_pair_flags : nullptr);
  count_launches(launched + 2);
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
  unsigned long long ng = 0;
  EDM_CUDA(cudaMemcpy(&ng, &b->d_state->n_pairs_ghost, sizeof(ng), cudaMemcpyDeviceToHost));
  result->n_calls = 2 * (long long)np - (long long)ng;  // a pair with a ghost proposes once, fix_edm_pair.cpp:233
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
  est_hill_count = edm_job_est(b, est_hill_count);
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
  est_hill_count = edm_job_est(b, est_hill_count);
  // Two streams so the copies hide behind the kernels that do not need them (pinned host buffers
  // make the copies truly asynchronous; pageable ones still work, staged by the driver):
  //   main: x up | binning, search                | evaluation | hill round | report
  //   copy:      | f up (needed by the evaluation) |            | f down (final once the evaluation ends)
  // One host synchronisation per stream at the end; the scalars come back through host-mapped memory.
  EDM_TRY(pair_host_streams(b));
  const size_t bx = (size_t)natoms * 3 * sizeof(double);
  EDM_TRY(b->io.reserve(bx));
  EDM_TRY(b->io2.reserve(bx));
  const bool prof = b->profiling != 0;
  if (prof) EDM_CUDA(cudaEventRecord(b->ev_e2e[0], b->st_main));
  const int* dt = nullptr;
  if (type) {
    EDM_TRY(b->io3.reserve((size_t)natoms * sizeof(int)));
    EDM_CUDA(cudaMemcpyAsync(b->io3.p, type, (size_t)natoms * sizeof(int), cudaMemcpyHostToDevice, b->st_main));
    dt = b->io3.as<int>();
  }
  EDM_CUDA(cudaMemcpyAsync(b->io.p, x, bx, cudaMemcpyHostToDevice, b->st_main));
  // the old forces follow the positions over the link (two concurrent uploads would share it and delay the
  // positions, which the binning and the search are waiting for; the forces are not needed before the evaluation)
  EDM_CUDA(cudaEventRecord(b->ev_f_final, b->st_main));  // reused here as "x is up"
  EDM_CUDA(cudaStreamWaitEvent(b->st_copy, b->ev_f_final, 0));
  if (prof) EDM_CUDA(cudaEventRecord(b->ev_e2e[1], b->st_main));
  EDM_CUDA(cudaMemcpyAsync(b->io2.p, f, bx, cudaMemcpyHostToDevice, b->st_copy));
  EDM_CUDA(cudaEventRecord(b->ev_f_up, b->st_copy));
  if (do_hills) EDM_TRY(edm_bias_reset_accepted(b, b->st_main));
  EDM_TRY(pair_cells_launch(b, natoms, b->io.as<double>(), b->io2.as<double>(), dt, itype, jtype, box, cutoff, do_hills,
                            est_hill_count, seed, step, b->d_scalar, b->st_main, b->ev_f_up, b->ev_f_final));
  EDM_CUDA(cudaStreamWaitEvent(b->st_copy, b->ev_f_final, 0));
  if (prof) EDM_CUDA(cudaEventRecord(b->ev_e2e[3], b->st_copy));
  EDM_CUDA(cudaMemcpyAsync(f, b->io2.p, bx, cudaMemcpyDeviceToHost, b->st_copy));
  if (prof) EDM_CUDA(cudaEventRecord(b->ev_e2e[4], b->st_copy));
  if (do_hills) EDM_TRY(edm_bias_launch_round(b, est_hill_count, b->st_main));
  step_report_kernel<<<1, 1, 0, b->st_main>>>(b->d_state, b->d_scalar, nullptr, reinterpret_cast<HostReport*>(b->d_pair_flags));
  count_launches(1);
  if (prof) EDM_CUDA(cudaEventRecord(b->ev_e2e[2], b->st_main));
  b->e2e_valid = prof ? 1 : 0;
  EDM_CUDA(cudaStreamSynchronize(b->st_main));
  const int rc = pair_host_result(b, result, do_hills);
  EDM_CUDA(cudaStreamSynchronize(b->st_copy));
  return rc;
}

// ---- one rank's share of a larger system: local atoms + ghosts inside a sub-box (LAMMPS' own picture) ----

int edm_pair_select_cells_domain_dev(edm_bias_t* b, long nall, const double* x, double* f, const int* type, int itype,
                                     int jtype, const edm_pair_domain_t* dom, double cutoff, long long est_hill_count,
                                     uint64_t seed, uint64_t step, double* energy_dev, void* stream) {
  EDM_REQUIRE(b && x && f && dom, "NULL argument");
  EDM_TRY(ensure_device(b->device));
  cudaStream_t st = (cudaStream_t)stream;
  EDM_TRY(edm_bias_reset_accepted(b, st));
  return pair_cells_launch(b, nall, x, f, type, itype, jtype, nullptr, cutoff, 1, est_hill_count, seed, step,
                           energy_dev ? energy_dev : b->d_scalar, st, nullptr, nullptr, dom);
}

int edm_pair_step_cells_domain_dev(edm_bias_t* b, long nall, const double* x, double* f, const int* type, int itype,
                                   int jtype, const edm_pair_domain_t* dom, double cutoff, int do_hills,
                                   long long est_hill_count, uint64_t seed, uint64_t step, edm_pair_result_t* result,
                                   void* stream) {
  EDM_REQUIRE(b && x && f && dom, "NULL argument");
  EDM_TRY(ensure_device(b->device));
  cudaStream_t st = (cudaStream_t)stream;
  est_hill_count = edm_job_est(b, est_hill_count);
  if (do_hills) EDM_TRY(edm_bias_reset_accepted(b, st));
  EDM_TRY(pair_cells_launch(b, nall, x, f, type, itype, jtype, nullptr, cutoff, do_hills, est_hill_count, seed, step,
                            b->d_scalar, st, nullptr, nullptr, dom));
  if (do_hills) EDM_TRY(edm_bias_launch_round(b, est_hill_count, st));
  if (result) {
    EDM_CUDA(cudaStreamSynchronize(st));
    EDM_TRY(read_pair_result(b, result, nullptr));
    if (do_hills) EDM_TRY(edm_bias_check_round(b));
  }
  return EDM_OK;
}

int edm_pair_step_cells_domain(edm_bias_t* b, long nall, const double* x, double* f, const int* type, int itype, int jtype,
                               const edm_pair_domain_t* dom, double cutoff, int do_hills, long long est_hill_count,
                               uint64_t seed, uint64_t step, edm_pair_result_t* result) {
  EDM_REQUIRE(b && x && f && dom && nall > 0, "bad argument");
  EDM_TRY(ensure_device(b->device));
  est_hill_count = edm_job_est(b, est_hill_count);
  EDM_TRY(pair_host_streams(b));
  const size_t bx = (size_t)nall * 3 * sizeof(double);
  const size_t bl = (size_t)dom->nlocal * 3 * sizeof(double);  // forces travel for the local atoms only
  EDM_TRY(b->io.reserve(bx));
  EDM_TRY(b->io2.reserve(bx));
  const int* dt = nullptr;
  if (type) {
    EDM_TRY(b->io3.reserve((size_t)nall * sizeof(int)));
    EDM_CUDA(cudaMemcpyAsync(b->io3.p, type, (size_t)nall * sizeof(int), cudaMemcpyHostToDevice, b->st_main));
    dt = b->io3.as<int>();
  }
  EDM_CUDA(cudaMemcpyAsync(b->io.p, x, bx, cudaMemcpyHostToDevice, b->st_main));
  EDM_CUDA(cudaEventRecord(b->ev_f_final, b->st_main));  // "x is up": the forces follow the positions over the link
  EDM_CUDA(cudaStreamWaitEvent(b->st_copy, b->ev_f_final, 0));
  if (bl) EDM_CUDA(cudaMemcpyAsync(b->io2.p, f, bl, cudaMemcpyHostToDevice, b->st_copy));
  EDM_CUDA(cudaEventRecord(b->ev_f_up, b->st_copy));
  if (do_hills) EDM_TRY(edm_bias_reset_accepted(b, b->st_main));
  EDM_TRY(pair_cells_launch(b, nall, b->io.as<double>(), b->io2.as<double>(), dt, itype, jtype, nullptr, cutoff, do_hills,
                            est_hill_count, seed, step, b->d_scalar, b->st_main, b->ev_f_up, b->ev_f_final, dom));
  EDM_CUDA(cudaStreamWaitEvent(b->st_copy, b->ev_f_final, 0));
  if (bl) EDM_CUDA(cudaMemcpyAsync(f, b->io2.p, bl, cudaMemcpyDeviceToHost, b->st_copy));
  if (do_hills) EDM_TRY(edm_bias_launch_round(b, est_hill_count, b->st_main));
  step_report_kernel<<<1, 1, 0, b->st_main>>>(b->d_state, b->d_scalar, nullptr, reinterpret_cast<HostReport*>(b->d_pair_flags));
  count_launches(1);
  b->e2e_valid = 0;
  EDM_CUDA(cudaStreamSynchronize(b->st_main));
  const int rc = pair_host_result(b, result, do_hills);
  EDM_CUDA(cudaStreamSynchronize(b->st_copy));
  return rc;
}

int edm_pair_search_info(edm_bias_t* b, int* brick_dims, double* density_scale, long long* fallbacks) {
  EDM_REQUIRE(b != nullptr, "NULL argument");
  if (b->h_pair_flags && b->h_pair_flags[0]) {  // a fallback reported since the last launch
    b->h_pair_flags[0] = 0;
    b->pair_fallbacks++;
    if (b->brick_scale < 64.0) b->brick_scale *= 1.5;
  }
  if (brick_dims)
    for (int d = 0; d < 3; d++) brick_dims[d] = b->brick_dims[d];
  if (density_scale) *density_scale = b->brick_scale;
  if (fallbacks) *fallbacks = b->pair_fallbacks;
  return EDM_OK;
}

int edm_pair_list_set(edm_bias_t* b, long inum, const int* ilist, const long* first, const int* jlist) {
  EDM_REQUIRE(b && inum >= 0 && (inum == 0 || (ilist && first)), "NULL argument");
  EDM_TRY(ensure_device(b->device));
  const long nlisted = inum ? first[inum] : 0;
  EDM_REQUIRE(nlisted == 0 || jlist, "NULL argument");
  // ilist[inum] | first[inum+1] | jlist[nlisted] | row_of[nlisted]
  const size_t o_first = ((size_t)inum * 4 + 7) / 8 * 8, o_jl = o_first + (size_t)(inum + 1) * 8;
  const size_t o_row = (o_jl + (size_t)nlisted * 4 + 7) / 8 * 8;
  EDM_TRY(b->list.reserve(o_row + (size_t)nlisted * 4 + 8));
  char* base = b->list.as<char>();
  b->list_inum = inum;
  b->list_nlisted = nlisted;
  b->list_ilist = reinterpret_cast<int*>(base);
  b->list_first = reinterpret_cast<long*>(base + o_first);
  b->list_jlist = reinterpret_cast<int*>(base + o_jl);
  b->list_row = reinterpret_cast<int*>(base + o_row);
  if (inum) {
    EDM_CUDA(cudaMemcpyAsync(b->list_ilist, ilist, (size_t)inum * 4, cudaMemcpyHostToDevice, 0));
    EDM_CUDA(cudaMemcpyAsync(b->list_first, first, (size_t)(inum + 1) * 8, cudaMemcpyHostToDevice, 0));
    if (nlisted) EDM_CUDA(cudaMemcpyAsync(b->list_jlist, jlist, (size_t)nlisted * 4, cudaMemcpyHostToDevice, 0));
    list_rows_kernel<<<(unsigned)((inum + 7) / 8), dim3(32, 8)>>>(inum, b->list_first, b->list_row);
    count_launches(1);
    EDM_CUDA(cudaGetLastError());
  }
  b->list_valid = 1;
  return EDM_OK;
}

int edm_pair_step_listed(edm_bias_t* b, long nall, long nlocal, const double* x, double* f, const int* type, int itype,
                         int jtype, int do_hills, long long est_hill_count, const double* runiform, uint64_t seed,
                         uint64_t step, edm_pair_result_t* result) {
  EDM_REQUIRE(b && x && f, "NULL argument");
  EDM_REQUIRE(b->list_valid, "edm_pair_list_set has not been called");
  EDM_REQUIRE(b->prm.dim == 1, "Pairwise distance must be 1 dimension in EDM input file");
  EDM_TRY(ensure_device(b->device));
  est_hill_count = edm_job_est(b, est_hill_count);
  const long nlisted = b->list_nlisted;
  const size_t bx = (size_t)nall * 3 * sizeof(double);
  EDM_TRY(b->io.reserve(bx));
  EDM_TRY(b->io2.reserve(bx));
  // per-step scratch: type[nall] | ncalls | fmax bits | cell records of the bias grid | uniforms
  const int npts = b->bias->d.n[0];
  const size_t o_nc = ((size_t)nall * 4 + 7) / 8 * 8, o_fm = o_nc + 8, o_crec = (o_fm + 16 + 31) / 32 * 32;
  const size_t o_u = o_crec + (size_t)npts * 4 * sizeof(double);
  EDM_TRY(b->io4.reserve(o_u + (runiform ? (size_t)nlisted * 2 * sizeof(double) : 0)));
  char* base = b->io4.as<char>();
  // Two streams, as in edm_pair_step_cells: the pair forces accumulate into a zeroed buffer while the old
  // forces are still on their way up; one pass adds the two, and the result travels down beside the hill round.
  //   main: x up | pair kernel -> dF          | f += dF | hill round
  //   copy:      | f up (after x, full rate)  |         | f down
  EDM_TRY(b->io3.reserve(bx));
  EDM_TRY(pair_host_streams(b));
  cudaStream_t sm = b->st_main, sc = b->st_copy;
  double* dF = b->io3.as<double>();
  if (type) EDM_CUDA(cudaMemcpyAsync(base, type, (size_t)nall * 4, cudaMemcpyHostToDevice, sm));
  if (runiform)
    EDM_CUDA(cudaMemcpyAsync(base + o_u, runiform, (size_t)nlisted * 2 * sizeof(double), cudaMemcpyHostToDevice, sm));
  EDM_CUDA(cudaMemcpyAsync(b->io.p, x, bx, cudaMemcpyHostToDevice, sm));
  EDM_CUDA(cudaEventRecord(b->ev_f_final, sm));  // reused: "x is up"
  EDM_CUDA(cudaStreamWaitEvent(sc, b->ev_f_final, 0));
  EDM_CUDA(cudaMemcpyAsync(b->io2.p, f, bx, cudaMemcpyHostToDevice, sc));
  EDM_CUDA(cudaEventRecord(b->ev_f_up, sc));
  EDM_CUDA(cudaMemsetAsync(dF, 0, bx, sm));
  unsigned long long* ncalls = reinterpret_cast<unsigned long long*>(base + o_nc);
  unsigned long long* fmax_bits = reinterpret_cast<unsigned long long*>(base + o_fm);
  double* cellrec = reinterpret_cast<double*>(base + o_crec);
  if (do_hills) {
    EDM_TRY(edm_bias_size_accepted(b, 2.0 * (double)nlisted, est_hill_count));
    EDM_TRY(edm_bias_reset_accepted(b, sm));
  }
  PairParams pp = pair_params(b, type, itype, jtype, do_hills, est_hill_count, seed, step, 0.0, nall);
  reset_pairs_kernel<<<1, 1, 0, sm>>>(b->d_state, ncalls);
  pair_prep_kernel<<<(npts + 255) / 256, 256, 0, sm>>>(b->bias->d, cellrec, fmax_bits);
  long long blocks = (nlisted + 255) / 256;
  if (blocks > b->n_partial) blocks = b->n_partial;
  if (blocks < 1) blocks = 1;
  pair_list_flat_kernel<<<(int)blocks, 256, 0, sm>>>(b->bias->d, cellrec, pp, nlocal, nlisted, b->list_ilist, b->list_row,
                                                     b->list_jlist, b->io.as<double>(),
                                                     type ? reinterpret_cast<int*>(base) : nullptr,
                                                     runiform ? reinterpret_cast<double*>(base + o_u) : nullptr, dF,
                                                     b->d_energy_partial, b->d_state, b->d_accepted, ncalls);
  sum_partials2_kernel<<<1, 256, 0, sm>>>((int)blocks, b->d_energy_partial, b->d_scalar, nullptr, nullptr);
  EDM_CUDA(cudaStreamWaitEvent(sm, b->ev_f_up, 0));
  add_forces_kernel<<<(unsigned)((nall * 3 + 255) / 256), 256, 0, sm>>>(nall * 3, dF, b->io2.as<double>());
  count_launches(5);
  EDM_CUDA(cudaGetLastError());
  EDM_CUDA(cudaEventRecord(b->ev_f_final, sm));
  EDM_CUDA(cudaStreamWaitEvent(sc, b->ev_f_final, 0));
  EDM_CUDA(cudaMemcpyAsync(f, b->io2.p, bx, cudaMemcpyDeviceToHost, sc));  // the round does not touch the forces
  if (do_hills) EDM_TRY(edm_bias_launch_round(b, est_hill_count, sm));
  step_report_kernel<<<1, 1, 0, sm>>>(b->d_state, b->d_scalar, ncalls, reinterpret_cast<HostReport*>(b->d_pair_flags));
  count_launches(1);
  EDM_CUDA(cudaStreamSynchronize(sm));
  const int rc = pair_host_result(b, result, do_hills);
  EDM_CUDA(cudaStreamSynchronize(sc));
  return rc;
}

int edm_pair_step_list(edm_bias_t* b, long nall, long nlocal, const double* x, double* f, const int* type, int itype,
                       int jtype, long inum, const int* ilist, const long* first, const int* jlist, int do_hills,
                       long long est_hill_count, const double* runiform, uint64_t seed, uint64_t step,
                       edm_pair_result_t* result) {
  EDM_REQUIRE(b && x && f && ilist && first && (jlist || first[inum] == 0), "NULL argument");
  EDM_TRY(edm_pair_list_set(b, inum, ilist, first, jlist));
  return edm_pair_step_listed(b, nall, nlocal, x, f, type, itype, jtype, do_hills, est_hill_count, runiform, seed, step,
                              result);
}

}  // extern "C"

