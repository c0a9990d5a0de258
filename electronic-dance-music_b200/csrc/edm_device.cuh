// Device-side building blocks shared by every EDM kernel (sm_100a, fp64 throughout).
//
// Layout in HBM (DESIGN.md "Data layout"):
//   * one RECORD per grid point, dim 0 fastest: {V, dV/dx0[, dV/dx1, dV/dx2]} padded to 2 doubles
//     in 1-D and 4 doubles in 2-D/3-D, so a corner of the interpolation stencil is one aligned
//     16 B / 32 B access and the two dim-0 neighbours share a sector pair;
//   * per-dimension POINT TABLES ptab[d][index][8] = {xx, inside_boundary, uL, uU, t6, t7, Z, Z'}:
//     everything in the McGovern-De Pablo deposit that depends on the grid point only
//     (lib/gaussian_grid.h:270,308-323), built once on the host with the reference's exact
//     expressions so grid-point coordinates and table look-ups are bit-identical.
//
// Where a rounding decides something discontinuous (cell index, wrap count, support test
// dp^2 < 8, boundary membership) the reference's operation order is kept with explicit _rn
// intrinsics, which nvcc never contracts into FMAs.  Elsewhere FMAs and
// reciprocals are used freely; those change results by a few ulp, far inside the 1e-10 bar.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace edm {

constexpr double kGaussSupport = 8.0;      // GAUSS_SUPPORT, lib/gaussian_grid.h:10
constexpr int kBcTableSize = 65536;        // BC_TABLE_SIZE, lib/gaussian_grid.h:11
constexpr double kBcMar = 2.0;             // BC_MAR, lib/gaussian_grid.h:12
constexpr double kInterpZero = 0.0000001;  // lib/grid.h:113 (T6)
constexpr int kPtabW = 8;

template <int DIM> struct RecW { static constexpr int value = (DIM == 1) ? 2 : 4; };

struct GridDesc {
  int dim;
  int rec_w;
  int n[3];
  int periodic[3];  // grid periodicity (b_periodic_)
  int b_interp, b_deriv, is_gauss;
  int dup_possible;  // a periodic window wider than the grid revisits points (lib/gaussian_grid.h:251-262)
  double min[3], max[3], dx[3];
  double len[3];     // max - min as the reference evaluates it
  double upper[3];   // max - dx: exclusive non-periodic bound (T4), evaluated on the host
  double inv_dx[3];
  double* rec;
  long long size;
  // GaussGrid part
  double sigma[3];        // sigma * sqrt(2) (T8)
  double sqrtpi_sigma[3]; // sqrt(M_PI) * sigma_  (lib/gaussian_grid.h:340)
  double bmin[3], bmax[3], blen[3];
  int bper[3];
  int minisize[3];
  int supp[3];  // cells from the centre cell beyond which the support test dp^2 < 8 always fails (<= minisize)
  const double* ptab[3];
  const long long* dup_pairs;  // (outer, bound) linear indices of duplicate_boundary, lib/gaussian_grid.h:571-630
  int n_dup;
  double vol_element;
};

// lib/grid.h:17-20 (T1): floor for every finite in-range argument
__device__ __forceinline__ int d_int_floor(double v) { return (int)floor(v); }
// lib/grid.h:22-26 (T2): round half away from zero
__device__ __forceinline__ double d_round(double v) { return v < 0.0 ? ceil(v - 0.5) : floor(v + 0.5); }

// x -= L * int_floor((x - min) / L), lib/grid.h:270.  The quotient floors to 0 exactly when
// 0 <= x - min < L, so the division is skipped for points already inside.
__device__ __forceinline__ double d_wrap(double x, double mn, double len) {
  double t = __dsub_rn(x, mn);
  if (t >= 0.0 && t < len) return x;
  double k = (double)d_int_floor(__ddiv_rn(t, len));
  return __dsub_rn(x, __dmul_rn(len, k));
}

// lib/gaussian_grid.h:490-499 (inclusive)
template <int DIM> __device__ __forceinline__ bool d_in_bounds(const GridDesc& g, const double* x) {
#pragma unroll
  for (int d = 0; d < DIM; d++)
    if (x[d] < g.bmin[d] || x[d] > g.bmax[d]) return false;
  return true;
}

// lib/gaussian_grid.h:504-541: nearest-image remap
template <int DIM> __device__ __forceinline__ void d_remap(const GridDesc& g, double* x) {
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    if (x[d] < g.min[d] || x[d] > g.max[d]) {
      if (g.periodic[d]) {
        double k = (double)d_int_floor(__ddiv_rn(__dsub_rn(x[d], g.min[d]), g.len[d]));
        x[d] = __dsub_rn(x[d], __dmul_rn(g.len[d], k));
      } else if (g.bper[d]) {
        double bl = g.blen[d];
        double dp0 = __dmul_rn(d_round(__ddiv_rn(__dsub_rn(g.min[d], x[d]), bl)), bl);
        double dp1 = __dmul_rn(d_round(__ddiv_rn(__dsub_rn(g.max[d], x[d]), bl)), bl);
        double a = fabs(__dsub_rn(__dsub_rn(g.min[d], x[d]), dp0));
        double b = fabs(__dsub_rn(__dsub_rn(g.max[d], x[d]), dp1));
        x[d] = __dadd_rn(x[d], (a < b) ? dp0 : dp1);
      }
    }
  }
}

template <int W> struct RecLoad;
template <> struct RecLoad<2> {
  __device__ __forceinline__ static void ld(const double* p, double* out) {
    double2 v = *reinterpret_cast<const double2*>(p);
    out[0] = v.x;
    out[1] = v.y;
  }
};
template <> struct RecLoad<4> {
  __device__ __forceinline__ static void ld(const double* p, double* out) {
    double2 a = *reinterpret_cast<const double2*>(p);
    double2 b = *reinterpret_cast<const double2*>(p + 2);
    out[0] = a.x;
    out[1] = a.y;
    out[2] = b.x;
    out[3] = b.y;
  }
};

// Grid::get_value_deriv (lib/grid.h:390-446) behind GaussGrid::get_value_deriv
// (lib/gaussian_grid.h:118-138), in two pieces so the hill round can interpolate from corner
// records it has patched itself: d_locate finds the cell, d_interp_cell blends its 2^DIM corners.

template <int DIM> struct CellLoc {
  long long base;         // linear index of the cell's low corner
  long long stride[DIM];  // linear step to the upper neighbour in each dim (wraps at a periodic edge)
  int idx[DIM];           // the low corner's grid index
  double X0[DIM];         // fractional position inside the cell
};

// in_bounds / remap / in_grid, then get_index + multi2one (lib/gaussian_grid.h:128-135,
// lib/grid.h:264-273, 315-325, 428-433).  False: the reference returns 0 for this point.
template <int DIM>
__device__ __forceinline__ bool d_locate(const GridDesc& g, const double* xin, CellLoc<DIM>& L) {
  double x[DIM];
#pragma unroll
  for (int d = 0; d < DIM; d++) x[d] = xin[d];
  if (g.is_gauss) {
    if (!d_in_bounds<DIM>(g, x)) {
      d_remap<DIM>(g, x);
      if (!d_in_bounds<DIM>(g, x)) return false;
    }
  }
#pragma unroll
  for (int d = 0; d < DIM; d++)
    if (!g.periodic[d] && (x[d] < g.min[d] || x[d] >= g.upper[d])) return false;  // in_grid (T4)

  long long base = 0, pstride = 1;
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    double xi = x[d];
    if (g.periodic[d]) xi = d_wrap(xi, g.min[d], g.len[d]);
    double t = __dsub_rn(xi, g.min[d]);
    // T5: the reference floors a true quotient.  t * (1/dx) is within 2 ulp of it, so its floor is the same unless
    // it lies that close to an integer; only then is the exactly rounded division paid for.
    // (1-D keeps the plain division: its callers are not bound by it, and the generic 1-D routine is also the
    // out-of-line slow path of the pair kernels, whose register budget it must not disturb.)
    double q;
    if constexpr (DIM == 1) {
      q = __ddiv_rn(t, g.dx[d]);
    } else {
      q = t * g.inv_dx[d];
      if (fabs(q - rint(q)) <= 8e-16 * fabs(q)) q = __ddiv_rn(t, g.dx[d]);
    }
    long long idx = (long long)floor(q);
    int nd = g.n[d];
    long long hi = g.periodic[d] ? nd - 1 : nd - 2;
    idx = idx < 0 ? 0 : (idx > hi ? hi : idx);  // rounding can land one past the last cell; clamp
    double where = __dsub_rn(t, __dmul_rn((double)idx, g.dx[d]));
    L.X0[d] = where * g.inv_dx[d];
    L.idx[d] = (int)idx;
    L.stride[d] = (g.periodic[d] && idx == nd - 1) ? pstride * (1 - nd) : pstride;  // lib/grid.h:432-433
    base += idx * pstride;
    pstride *= nd;
  }
  L.base = base;
  return true;
}

template <int DIM> __device__ __forceinline__ long long d_corner_shift(const CellLoc<DIM>& L, int c) {
  long long shift = 0;
#pragma unroll
  for (int d = 0; d < DIM; d++)
    if ((c >> d) & 1) shift += L.stride[d];
  return shift;
}

// interp<DIM> (lib/grid.h:52-139) restructured:
//   tabf*C_d = tabf*a(X) + s*tabder_d*b(X)*dx   and   tabf*D_d = tabf*a'(X)*s/dx + tabder_d*c(X)
// in 1-D (no division), one reciprocal of tabf per corner in 2-D/3-D.  load(c, r) fills r with the
// record {V, dV/dx...} of corner c (bit d of c set = upper neighbour in dim d).
template <int DIM, typename Load>
__device__ __forceinline__ double d_interp_cell(const GridDesc& g, const double* X0, double* der, Load load) {
  constexpr int W = RecW<DIM>::value;
  constexpr int NC = 1 << DIM;
  double f = 0.0;
  if constexpr (DIM == 1) {
#pragma unroll
    for (int c = 0; c < NC; c++) {
      double r[W];
      load(c, r);
      double tabf = r[0];
      bool nz = !(fabs(tabf) < kInterpZero);  // T6
      int b = c & 1;
      double X = fabs(X0[0] - (double)b);
      double s = b ? -1.0 : 1.0;
      double X2 = X * X, X3 = X2 * X;
      double a = 1.0 - 3.0 * X2 + 2.0 * X3;
      double bp = X - 2.0 * X2 + X3;
      double ap = -6.0 * X + 6.0 * X2;
      double cc = 1.0 - 4.0 * X + 3.0 * X2;
      double td = nz ? r[1] : 0.0;
      f += tabf * a + s * td * bp * g.dx[0];
      der[0] += tabf * ap * s * g.inv_dx[0] + td * cc;
    }
    return f;
  } else {
  // 2-D / 3-D: every corner record first (the loads are in flight together), then ONE reciprocal per four corners:
  // 1/t_c = (product of the other three) / (product of all four).  Corners the reference treats as zero (T6) enter
  // as 1.  |t| >= 1e-7 for the others, so the product stays a normal number for any bias below 1e70; outside that
  // range each corner gets its own division.
  double r[NC][W], rinv[NC];
#pragma unroll
  for (int c = 0; c < NC; c++) load(c, r[c]);
#pragma unroll
  for (int q = 0; q < NC; q += 4) {
    bool nz[4];
    double t[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      nz[c] = !(fabs(r[q + c][0]) < kInterpZero);  // T6
      t[c] = nz[c] ? r[q + c][0] : 1.0;
    }
    const double p01 = t[0] * t[1], p23 = t[2] * t[3];
    const double all = p01 * p23, mag = fabs(all);
    if (mag > 1e-280 && mag < 1e280) {
      const double inv = 1.0 / all;
      const double i01 = inv * p23, i23 = inv * p01;  // 1/(t0 t1), 1/(t2 t3)
      rinv[q + 0] = nz[0] ? i01 * t[1] : 0.0;
      rinv[q + 1] = nz[1] ? i01 * t[0] : 0.0;
      rinv[q + 2] = nz[2] ? i23 * t[3] : 0.0;
      rinv[q + 3] = nz[3] ? i23 * t[2] : 0.0;
    } else {
#pragma unroll
      for (int c = 0; c < 4; c++) rinv[q + c] = nz[c] ? 1.0 / t[c] : 0.0;
    }
  }
  // the Hermite factors depend on (dimension, side) only
  double A[DIM][2], B[DIM][2], AP[DIM][2], CP[DIM][2];
#pragma unroll
  for (int d = 0; d < DIM; d++) {
#pragma unroll
    for (int b = 0; b < 2; b++) {
      double X = fabs(X0[d] - (double)b);
      double s = b ? -1.0 : 1.0;
      double X2 = X * X, X3 = X2 * X;
      A[d][b] = 1.0 - 3.0 * X2 + 2.0 * X3;
      B[d][b] = s * (X - 2.0 * X2 + X3) * g.dx[d];
      AP[d][b] = (-6.0 * X + 6.0 * X2) * s * g.inv_dx[d];
      CP[d][b] = (1.0 - 4.0 * X + 3.0 * X2);  // * s * dx * s / dx
    }
  }
#pragma unroll
  for (int c = 0; c < NC; c++) {
    const double tabf = r[c][0];
    double C[DIM], D[DIM];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      const int b = (c >> d) & 1;
      const double qq = -r[c][1 + d] * rinv[c];
      C[d] = A[d][b] - qq * B[d][b];
      D[d] = AP[d][b] - qq * CP[d][b];
    }
    double ff = 1.0;
#pragma unroll
    for (int d = 0; d < DIM; d++) ff *= C[d];
    f += tabf * ff;
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      double fd = D[d];
#pragma unroll
      for (int e = 0; e < DIM; e++)
        if (e != d) fd *= C[e];
      der[d] += tabf * fd;
    }
  }
  return f;
  }
}

template <int DIM>
__device__ __forceinline__ double d_eval_point(const GridDesc& g, const double* xin, double* der, bool interp) {
  constexpr int W = RecW<DIM>::value;
#pragma unroll
  for (int d = 0; d < DIM; d++) der[d] = 0.0;
  CellLoc<DIM> L;
  if (!d_locate<DIM>(g, xin, L)) return 0.0;
  if (!interp) {  // lib/grid.h:438-443
    double r[W];
    RecLoad<W>::ld(g.rec + L.base * W, r);
#pragma unroll
    for (int d = 0; d < DIM; d++) der[d] = r[1 + d];
    return r[0];
  }
  return d_interp_cell<DIM>(g, L.X0, der, [&](int c, double* r) {
    RecLoad<W>::ld(g.rec + (L.base + d_corner_shift<DIM>(L, c)) * W, r);
  });
}

// Grid::get_value (lib/grid.h:343-365) behind GaussGrid::get_value (lib/gaussian_grid.h:99-116)
template <int DIM> __device__ __forceinline__ double d_get_value(const GridDesc& g, const double* xin) {
  double der[DIM];
  return d_eval_point<DIM>(g, xin, der, g.b_interp && g.b_deriv);
}

// ---------------------------------------------------------------- hill deposit pieces

// exp(x) for -9 < x <= 0, the only range the Gaussian of a hill is evaluated on (sum dp^2 < 8).
// Cody-Waite reduction x = k ln2 + r, |r| <= ln2/2, degree-13 Taylor polynomial (truncation 4e-18
// relative), scaling through the exponent bits; error below 1 ulp like libdevice's exp, a handful of
// ulp from glibc's at worst -- six orders inside the 1e-10 bar.  The point is the instruction count of
// the deposit loops, which are issue-bound: the coefficients sit in the constant bank, where DFMA
// reads them as operands, instead of being rebuilt in uniform registers (two UMOV per coefficient)
// on every evaluation as the inlined libdevice code does, and the range checks are gone.
__constant__ double kExpTaylor[14] = {1.0,
                                      1.0,
                                      1.0 / 2,
                                      1.0 / 6,
                                      1.0 / 24,
                                      1.0 / 120,
                                      1.0 / 720,
                                      1.0 / 5040,
                                      1.0 / 40320,
                                      1.0 / 362880,
                                      1.0 / 3628800,
                                      1.0 / 39916800,
                                      1.0 / 479001600,
                                      1.0 / 6227020800.0};
__constant__ double kExpRed[4] = {1.4426950408889634074, 6755399441055744.0, -6.93147180369123816490e-01,
                                  -1.90821492927058770002e-10};

__device__ __forceinline__ double d_exp_support(double x) {
  const double t = fma(x, kExpRed[0], kExpRed[1]);  // the low word of t is rint(x log2 e)
  const int k = __double2loint(t);
  const double kd = t - kExpRed[1];
  double r = fma(kd, kExpRed[2], x);
  r = fma(kd, kExpRed[3], r);
  double p = kExpTaylor[13];
#pragma unroll
  for (int i = 12; i >= 0; i--) p = fma(p, r, kExpTaylor[i]);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

template <int DIM> struct HillGeom {
  double x[DIM];   // centre after remap
  int xi[DIM];     // centre cell, may be negative (lib/gaussian_grid.h:222-224)
  double t1[DIM];  // exp(-(x-L)^2/sigma^2), lib/gaussian_grid.h:310
  double t3[DIM];  // exp(-(x-U)^2/sigma^2), lib/gaussian_grid.h:312
};

// lib/gaussian_grid.h:205-224 (T10): false when the centre lies outside a non-periodic boundary
template <int DIM>
__device__ __forceinline__ bool d_hill_prepare(const GridDesc& g, const double* x0, HillGeom<DIM>& hg) {
#pragma unroll
  for (int d = 0; d < DIM; d++) hg.x[d] = x0[d];
  d_remap<DIM>(g, hg.x);
#pragma unroll
  for (int d = 0; d < DIM; d++)
    if (!g.bper[d] && (hg.x[d] < g.bmin[d] || hg.x[d] > g.bmax[d])) return false;
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    hg.xi[d] = d_int_floor(__ddiv_rn(__dsub_rn(hg.x[d], g.min[d]), g.dx[d]));
    hg.t1[d] = 0.0;
    hg.t3[d] = 0.0;
    if (!g.bper[d]) {
      double s2 = __dmul_rn(g.sigma[d], g.sigma[d]);
      double a = __dsub_rn(hg.x[d], g.bmin[d]);
      double b = __dsub_rn(hg.x[d], g.bmax[d]);
      hg.t1[d] = exp(__ddiv_rn(-__dmul_rn(a, a), s2));
      hg.t3[d] = exp(__ddiv_rn(-__dmul_rn(b, b), s2));
    }
  }
  return true;
}

// One (hill, grid point) term of lib/gaussian_grid.h:283-355 incl. the multi-D McGDP quirks
// (T11).  idx[] must already be a valid wrapped grid index.  On success: etot = expo + corr (per
// unit height), force[d] = per-unit-height addend of grid_deriv_[.][d].
template <int DIM>
__device__ __forceinline__ bool d_hill_term(const GridDesc& g, const HillGeom<DIM>& hg, const int* idx,
                                            double& etot, double* force, bool& corr_nonzero) {
  double dp[DIM];
  const double* row[DIM];
  double dp2 = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    row[d] = g.ptab[d] + (long long)idx[d] * kPtabW;
    if (row[d][1] == 0.0) return false;  // outside a non-periodic boundary (T10)
    double v = __dsub_rn(row[d][0], hg.x[d]);
    if (g.periodic[d]) v = __dsub_rn(v, __dmul_rn(d_round(__ddiv_rn(v, g.len[d])), g.len[d]));
    v = __ddiv_rn(v, g.sigma[d]);
    dp[d] = v;
    dp2 = __dadd_rn(dp2, __dmul_rn(v, v));
  }
  if (!(dp2 < kGaussSupport)) return false;
  double expo = d_exp_support(-dp2);
  double bc_denom = 1.0, corr = 0.0;
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    if (!g.bper[d]) {
      const double uL = row[d][2], uU = row[d][3], t6 = row[d][4], t7 = row[d][5];
      corr = (hg.t1[d] - expo) * uL + (hg.t3[d] - expo) * uU;
      bc_denom *= row[d][6];
      double t5 = -2.0 * dp[d] / g.sigma[d];
      double F = t5 * expo;
      F += (hg.t1[d] - expo) * t6 - t5 * expo * uL + (hg.t3[d] - expo) * t7 - t5 * expo * uU;
      F = F * bc_denom - row[d][7] * (expo + corr);
      F /= bc_denom * bc_denom;
      corr /= bc_denom;
      force[d] = F;
    } else {
      bc_denom *= g.sqrtpi_sigma[d];
    }
  }
  expo /= bc_denom;
  etot = expo + corr;
#pragma unroll
  for (int d = 0; d < DIM; d++)
    if (g.bper[d]) force[d] = -(2.0 * dp[d] / g.sigma[d] * expo);
  corr_nonzero = (corr * corr > 0.0);
  return true;
}

// Maps window offset w (dim 0 fastest over 2*supp+1 per dim: the part of the reference's
// 2*minisize+1 window that can pass the support test) to a wrapped grid index,
// lib/gaussian_grid.h:229-268.  False when the point falls off a non-periodic grid.
template <int DIM>
__device__ __forceinline__ bool d_window_index(const GridDesc& g, const HillGeom<DIM>& hg, long long w, int* idx,
                                               long long& linear) {
  linear = 0;
  long long pstride = 1;
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    int span = 2 * g.supp[d] + 1;
    int off;
    if (d < DIM - 1) {
      off = (int)(w % span);
      w = (w - off) / span;
    } else {
      off = (int)w;
    }
    int i = off - g.supp[d] + hg.xi[d];
    if (i >= g.n[d]) {
      if (!g.periodic[d]) return false;
      i %= g.n[d];
    }
    if (i < 0) {
      if (!g.periodic[d]) return false;
      i += g.n[d];
      if (i < 0) return false;  // window wider than two periods: outside anything the reference can index
    }
    idx[d] = i;
    linear += (long long)i * pstride;
    pstride *= g.n[d];
  }
  return true;
}

// programmatic dependent launch (see launch_pdl in edm_host.h)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- reductions / RNG

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the CTA in a fixed order (lane tree, then warps in order): run-to-run deterministic.
// red must hold 33 doubles of shared memory.  Every thread receives the total.
__device__ __forceinline__ double block_sum(double v, double* red) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; i++) t += red[i];
    red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// ---------------------------------------------------------------- one hill's window over a CTA

// One hill's window spread over the CTA.  kPassStore: plain read-modify-writes (one hill at a time,
// so no two threads touch the same point unless a periodic window revisits it; then fp64 atomics,
// whose operands are equal, keep the result order-independent).  kPassAtomic: fp64 REDs, for hills
// deposited concurrently.  kPassIntegrate: no writes, only add_value's integral.  The window-point ->
// thread mapping and the reduction tree are the same in every mode, so the integral comes out
// bit-identical whichever pass computed it.  Every thread receives the integral.
//
// When every dimension has a periodic boundary the hill is a plain product Gaussian
// (lib/gaussian_grid.h:340-352): the CTA first tabulates, per dimension and window offset, the
// wrapped grid index, dp^2, exp(-dp^2) and 2 dp / sigma~ (a few dozen exp instead of one per window
// point), then every window point is three table look-ups, the support test on the same sum of
// squares the reference forms, two multiplications and the adds.
// kPassOrdered: plain vector read-modify-writes through L2 (ld.cg / st.cg of the whole record) for the round's
// ticket-ordered deposit, where hills that run concurrently never overlap and a later overlapping hill has acquired
// its predecessor's release: no atomics needed, a third of the memory operations of kPassAtomic.
enum { kPassIntegrate = 0, kPassStore = 1, kPassAtomic = 2, kPassOrdered = 3 };
constexpr int kAxisMax = 96;  // window offsets per dimension the tables hold; wider windows take the general path

struct AxisEntry {
  double dpsq, e, t5;
  int idx;  // wrapped grid index, -1: the reference skips this offset
  int pad;
};

template <int DIM> __device__ __forceinline__ bool d_separable(const GridDesc& g) {
  if (DIM == 1 || g.dup_possible) return false;
#pragma unroll
  for (int d = 0; d < DIM; d++)
    if (!g.bper[d] || 2 * g.supp[d] + 1 > kAxisMax) return false;
  return true;
}

template <int DIM, int MODE>
__device__ __forceinline__ void d_point_add(const GridDesc& g, long long lin, double add, double h, const double* force) {
  constexpr int W = RecW<DIM>::value;
  if (MODE == kPassIntegrate) return;
  double* r = g.rec + lin * W;
  if (MODE == kPassOrdered && !g.dup_possible) {
    if (W == 2) {
      double2 v = __ldcg(reinterpret_cast<const double2*>(r));
      v.x += add;
      v.y += h * force[0];
      __stcg(reinterpret_cast<double2*>(r), v);
    } else {
      double2 a = __ldcg(reinterpret_cast<const double2*>(r));
      double2 b = __ldcg(reinterpret_cast<const double2*>(r) + 1);
      a.x += add;
      a.y += h * force[0];
      b.x += h * force[DIM > 1 ? 1 : 0];
      if (DIM > 2) b.y += h * force[DIM > 2 ? 2 : 0];
      __stcg(reinterpret_cast<double2*>(r), a);
      __stcg(reinterpret_cast<double2*>(r) + 1, b);
    }
    return;
  }
  if (MODE == kPassAtomic || MODE == kPassOrdered || g.dup_possible) {
    atomicAdd(r, add);
#pragma unroll
    for (int d = 0; d < DIM; d++) atomicAdd(r + 1 + d, h * force[d]);
  } else {
    r[0] += add;
#pragma unroll
    for (int d = 0; d < DIM; d++) r[1 + d] += h * force[d];
  }
}

// axis: shared memory, DIM * kAxisMax entries (only touched on the separable path)
template <int DIM, int MODE>
__device__ double cta_window_pass(const GridDesc& g, const double* x0, double h, double* red, bool& dirty,
                                  AxisEntry* axis) {
  HillGeom<DIM> hg;
  dirty = false;
  double ba = 0.0;
  const bool ok = d_hill_prepare<DIM>(g, x0, hg);
  long long total = 1;
#pragma unroll
  for (int d = 0; d < DIM; d++) total *= (2 * g.supp[d] + 1);
  if (ok && d_separable<DIM>(g)) {
    __syncthreads();  // the tables may still be read by the previous hill's pass
    for (int t = threadIdx.x; t < DIM * kAxisMax; t += blockDim.x) {
      const int d = t / kAxisMax, o = t - d * kAxisMax;
      if (o > 2 * g.supp[d]) continue;
      AxisEntry a;
      a.idx = -1;
      a.dpsq = a.e = a.t5 = 0.0;
      a.pad = 0;
      int i = o - g.supp[d] + hg.xi[d];  // lib/gaussian_grid.h:229-268
      bool in = true;
      if (i >= g.n[d]) {
        if (!g.periodic[d]) in = false;
        i %= g.n[d];
      }
      if (i < 0) {
        if (!g.periodic[d]) in = false;
        i += g.n[d];
        if (i < 0) in = false;
      }
      if (in) {
        const double* row = g.ptab[d] + (long long)i * kPtabW;
        double v = __dsub_rn(row[0], hg.x[d]);
        if (g.periodic[d]) v = __dsub_rn(v, __dmul_rn(d_round(__ddiv_rn(v, g.len[d])), g.len[d]));
        v = __ddiv_rn(v, g.sigma[d]);
        const double sq = __dmul_rn(v, v);
        if (row[1] != 0.0 && sq < kGaussSupport) {  // a larger square alone already fails the support test
          a.idx = i;
          a.dpsq = sq;
          a.e = d_exp_support(-sq);
          a.t5 = 2.0 * v / g.sigma[d];
        }
      }
      axis[t] = a;
    }
    __syncthreads();
    double inv_denom = 1.0;
#pragma unroll
    for (int d = 0; d < DIM; d++) inv_denom *= g.sqrtpi_sigma[d];
    inv_denom = 1.0 / inv_denom;
    // every span is at most kAxisMax here, so the window has fewer than 2^20 points: 32-bit index arithmetic
    // (a 64-bit division by a run-time span costs more than the rest of the loop body)
    const unsigned span0 = 2 * g.supp[0] + 1, span1 = 2 * g.supp[DIM > 1 ? 1 : 0] + 1;
    const unsigned total32 = (unsigned)total;
    // one window point: table look-ups, the support test on the sum of squares the reference forms
    auto point = [&](unsigned w, long long& lin, double& add, double* force) -> bool {
      int o[3];
      unsigned q = w / span0;
      o[0] = (int)(w - q * span0);
      if (DIM > 2) {
        const unsigned q2 = q / span1;
        o[1] = (int)(q - q2 * span1);
        o[2] = (int)q2;
      } else {
        o[1] = (int)q;
        o[2] = 0;
      }
      double dp2 = 0.0, expo = inv_denom, t5[DIM];
      long long pstride = 1;
      lin = 0;
      bool in = true;
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        const AxisEntry a = axis[d * kAxisMax + o[d]];
        in = in && a.idx >= 0;
        dp2 = __dadd_rn(dp2, a.dpsq);
        expo *= a.e;
        t5[d] = a.t5;
        lin += (long long)a.idx * pstride;
        pstride *= g.n[d];
      }
      if (!in || !(dp2 < kGaussSupport)) return false;
#pragma unroll
      for (int d = 0; d < DIM; d++) force[d] = -(t5[d] * expo);
      add = h * expo;
      return true;
    };
    // (Two window points per trip with both record loads in flight was measured for the ordered deposit: no gain in
    // 2-D, 20 % slower in 3-D, where the pass moves 64 B per point through L2 and is bound by that, not by latency.)
    for (unsigned w = threadIdx.x; w < total32; w += blockDim.x) {
      long long lin;
      double add, force[DIM];
      if (!point(w, lin, add, force)) continue;
      d_point_add<DIM, MODE>(g, lin, add, h, force);
      ba += add * g.vol_element;
    }
  } else if (ok) {
    for (long long w = threadIdx.x; w < total; w += blockDim.x) {
      int idx[DIM];
      long long lin;
      if (!d_window_index<DIM>(g, hg, w, idx, lin)) continue;
      double etot, force[DIM];
      bool cnz;
      if (!d_hill_term<DIM>(g, hg, idx, etot, force, cnz)) continue;
      const double add = h * etot;
      d_point_add<DIM, MODE>(g, lin, add, h, force);
      ba += add * g.vol_element;
      dirty |= cnz;
    }
  }
  if (MODE == kPassOrdered) {  // the integral was taken by the integrals pass; the caller fences and signals itself
    __syncthreads();
    return 0.0;
  }
  return block_sum(ba, red);  // contains __syncthreads: record writes are visible CTA-wide after it
}

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
// stream key for (seed, step); computed once per launch on the host
__host__ __device__ __forceinline__ uint64_t uniform_key(uint64_t seed, uint64_t step) {
  return mix64(seed ^ mix64(step + 0x9E3779B97F4A7C15ULL));
}
__host__ __device__ __forceinline__ double uniform_from_key(uint64_t key, uint64_t counter) {
  uint64_t bits = mix64(key + counter * 0x9E3779B97F4A7C15ULL);
  return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
}

// pair proposals: one hash per pair, one 32-bit uniform per proposal (LAMMPS' RanMars has 24 bits)
__host__ __device__ __forceinline__ uint64_t pair_bits(uint64_t key, uint64_t pairkey) {
  return mix64(key + pairkey * 0x9E3779B97F4A7C15ULL);
}
__host__ __device__ __forceinline__ double pair_uniform_from_bits(uint64_t bits, int which) {
  uint64_t half = which == 0 ? (bits >> 32) : (bits & 0xffffffffULL);
  return (double)half * (1.0 / 4294967296.0);
}

}  // namespace edm
