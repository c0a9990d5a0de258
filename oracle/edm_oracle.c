/* CPU restatement of the EDM per-timestep bias engine.  TEST INFRASTRUCTURE ONLY.
 * See edm_oracle.h for scope, parity status ("pinned") and build flags.
 * Citations are file:line into the reference tree (whitead/electronic-dance-music).
 */
#include "edm_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ scalars */

/* lib/grid.h:17-20 (T1): the cast binds before the comparison; still floor() for in-range x. */
static int orc_int_floor(double number) {
  return (int)((int)number < 0.0 ? -ceil(fabs(number)) : floor(number));
}

/* lib/grid.h:22-26 (T2): round half away from zero. */
static double orc_round(double number) {
  return number < 0.0 ? ceil(number - 0.5) : floor(number + 0.5);
}

/* lib/gaussian_grid.h:16-23 */
static double orc_sigmoid(double x) {
  if (x < 0) return 1;
  if (x > 1) return 0;
  return 2 * x * x * x - 3 * x * x + 1;
}

/* lib/gaussian_grid.h:25-32 */
static double orc_sigmoid_dx(double x) {
  if (x < 0) return 0;
  if (x > 1) return 0;
  return 6 * x * x - 6 * x;
}

/* ------------------------------------------------------------------ plain grid */

/* lib/grid.h:892-904 (T7: the reference over-allocates; sizes here are the used ones) */
static void grid_alloc(orc_grid* g) {
  g->size = 1;
  for (int i = 0; i < g->dim; i++) g->size *= (size_t)g->n[i];
  g->grid = (double*)calloc(g->size, sizeof(double));
  g->deriv = g->b_derivatives ? (double*)calloc(g->size * (size_t)g->dim, sizeof(double)) : NULL;
}

/* lib/grid.h:190-213 (T3): n = ceil(L/spacing), dx = L/n; non-periodic adds a point and a dx. */
static void grid_init(orc_grid* g, int dim, const double* mn, const double* mx, const double* spacing,
                      const int* periodic, int b_deriv, int b_interp) {
  memset(g, 0, sizeof(*g));
  g->dim = dim;
  g->b_derivatives = b_deriv;
  g->b_interpolate = b_interp;
  for (int i = 0; i < dim; i++) {
    g->min[i] = mn[i];
    g->max[i] = mx[i];
    g->periodic[i] = periodic[i];
    g->n[i] = (int)ceil((g->max[i] - g->min[i]) / spacing[i]);
    g->dx[i] = (g->max[i] - g->min[i]) / g->n[i];
    g->n[i] = g->periodic[i] ? g->n[i] : g->n[i] + 1;
    if (!g->periodic[i]) g->max[i] += g->dx[i];
  }
  grid_alloc(g);
}

static void grid_free(orc_grid* g) {
  free(g->grid);
  free(g->deriv);
  g->grid = g->deriv = NULL;
}

/* lib/grid.h:865-874 (T4) */
static int grid_in_grid(const orc_grid* g, const double* x) {
  for (int i = 0; i < g->dim; i++)
    if (!g->periodic[i] && (x[i] < g->min[i] || x[i] >= g->max[i] - g->dx[i])) return 0;
  return 1;
}

/* lib/grid.h:264-273 (T5: a true division) */
static void grid_get_index(const orc_grid* g, const double* x, size_t* result) {
  for (int i = 0; i < g->dim; i++) {
    double xi = x[i];
    if (g->periodic[i])
      xi -= (g->max[i] - g->min[i]) * orc_int_floor((xi - g->min[i]) / (g->max[i] - g->min[i]));
    result[i] = (size_t)floor((xi - g->min[i]) / g->dx[i]);
  }
}

/* lib/grid.h:315-325: dim 0 runs fastest */
static size_t grid_multi2one(const orc_grid* g, const size_t* index) {
  size_t result = index[g->dim - 1];
  for (int i = g->dim - 1; i > 0; i--) result = result * (size_t)g->n[i - 1] + index[i - 1];
  return result;
}

/* lib/grid.h:52-139: PLUMED-1.3 style Hermite blend over the 2^D corners (a2, T6). */
static double orc_interp(int dim, const double* dx, const double* where, const double* tabf,
                         const double* tabder, const int* stride, double* der) {
  int npoints = 1 << dim;
  double f = 0;
  for (int d = 0; d < dim; d++) der[d] = 0;
  for (int ipoint = 0; ipoint < npoints; ipoint++) {
    int x0[3];
    double C[3], D[3], fd[3];
    int tmp = ipoint, shift = 0;
    for (int d = 0; d < dim; d++) {
      x0[d] = tmp % 2;
      tmp /= 2;
      shift += stride[d] * x0[d];
    }
    double ff = 1.0;
    for (int d = 0; d < dim; d++) {
      double X = fabs(where[d] / dx[d] - x0[d]);
      double X2 = X * X;
      double X3 = X2 * X;
      double qq;
      if (fabs(tabf[shift]) < 0.0000001)
        qq = 0.0;
      else
        qq = -tabder[shift * dim + d] / tabf[shift];
      C[d] = (1 - 3 * X2 + 2 * X3) - (x0[d] ? -1 : 1) * qq * (X - 2 * X2 + X3) * dx[d];
      D[d] = (-6 * X + 6 * X2) - (x0[d] ? -1 : 1) * qq * (1 - 4 * X + 3 * X2) * dx[d];
      D[d] *= (x0[d] ? -1 : 1) / dx[d];
      ff *= C[d];
    }
    for (int d = 0; d < dim; d++) {
      fd[d] = D[d];
      for (int e = 0; e < dim; e++)
        if (e != d) fd[d] *= C[e];
    }
    f += tabf[shift] * ff;
    for (int d = 0; d < dim; d++) der[d] += tabf[shift] * fd[d];
  }
  return f;
}

/* lib/grid.h:390-446 */
static double grid_get_value_deriv(const orc_grid* g, const double* x, double* der) {
  size_t index[3];
  if (!grid_in_grid(g, x)) {
    for (int i = 0; i < g->dim; i++) der[i] = 0;
    return 0;
  }
  grid_get_index(g, x, index);
  size_t index1 = grid_multi2one(g, index);
  if (g->b_interpolate) {
    double where[3];
    int stride[3];
    stride[0] = 1;
    for (int i = 1; i < g->dim; i++) stride[i] = stride[i - 1] * g->n[i - 1];
    for (int i = 0; i < g->dim; i++) {
      double wrapped_x = x[i];
      if (g->periodic[i])
        wrapped_x -= (g->max[i] - g->min[i]) * orc_int_floor((wrapped_x - g->min[i]) / (g->max[i] - g->min[i]));
      where[i] = wrapped_x - g->min[i] - index[i] * g->dx[i];
      if (g->periodic[i] && index[i] == (size_t)(g->n[i] - 1)) stride[i] *= (1 - g->n[i]);
    }
    return orc_interp(g->dim, g->dx, where, &g->grid[index1], &g->deriv[index1 * (size_t)g->dim], stride, der);
  }
  for (int i = 0; i < g->dim; i++) der[i] = g->deriv[index1 * (size_t)g->dim + i];
  return g->grid[index1];
}

/* lib/grid.h:343-365 */
static double grid_get_value(const orc_grid* g, const double* x) {
  if (!grid_in_grid(g, x)) return 0;
  if (g->b_interpolate && g->b_derivatives) {
    double temp[3];
    return grid_get_value_deriv(g, x, temp);
  }
  size_t index[3];
  grid_get_index(g, x, index);
  return g->grid[grid_multi2one(g, index)];
}

/* lib/grid.h:370-385: histogram bump (a5); the reference aborts when interpolation is on */
static double grid_add_value(orc_grid* g, const double* x0, double value) {
  if (g->b_interpolate) {
    fprintf(stderr, "[oracle] Cannot add_value when using derivatives\n");
    abort();
  }
  if (!grid_in_grid(g, x0)) return 0;
  size_t index[3];
  grid_get_index(g, x0, index);
  g->grid[grid_multi2one(g, index)] += value;
  return value;
}

/* lib/grid.h:692-710 */
static double grid_expected_bias(const orc_grid* g) {
  double Z = 0, offset = 0, avg = 0;
  for (size_t i = 0; i < g->size; i++) offset = fmax(offset, g->grid[i]);
  for (size_t i = 0; i < g->size; i++) Z += exp(-g->grid[i] - offset);
  for (size_t i = 0; i < g->size; i++) avg += g->grid[i] * exp(-g->grid[i] - offset);
  return avg / Z;
}

void* orc_grid_create(int dim, const double* mn, const double* mx, const double* spacing,
                      const int* periodic, int b_deriv, int b_interp) {
  orc_grid* g = (orc_grid*)malloc(sizeof(orc_grid));
  grid_init(g, dim, mn, mx, spacing, periodic, b_deriv, b_interp);
  return g;
}
void orc_grid_destroy(void* p) {
  grid_free((orc_grid*)p);
  free(p);
}
void orc_grid_info(void* p, int* n, double* dx, double* mn, double* mx, int* flags) {
  orc_grid* g = (orc_grid*)p;
  for (int i = 0; i < g->dim; i++) {
    n[i] = g->n[i];
    dx[i] = g->dx[i];
    mn[i] = g->min[i];
    mx[i] = g->max[i];
    flags[2 + i] = g->periodic[i];
  }
  flags[0] = g->b_derivatives;
  flags[1] = g->b_interpolate;
}
size_t orc_grid_size(void* p) { return ((orc_grid*)p)->size; }
void orc_grid_get_arrays(void* p, double* v, double* d) {
  orc_grid* g = (orc_grid*)p;
  memcpy(v, g->grid, g->size * sizeof(double));
  if (g->b_derivatives && d) memcpy(d, g->deriv, g->size * (size_t)g->dim * sizeof(double));
}
void orc_grid_set_arrays(void* p, double* v, double* d) {
  orc_grid* g = (orc_grid*)p;
  memcpy(g->grid, v, g->size * sizeof(double));
  if (g->b_derivatives && d) memcpy(g->deriv, d, g->size * (size_t)g->dim * sizeof(double));
}
void orc_grid_set_interpolation(void* p, int b) { ((orc_grid*)p)->b_interpolate = b; }
void orc_grid_eval(void* p, long n, const double* x, double* val, double* der) {
  orc_grid* g = (orc_grid*)p;
  for (long i = 0; i < n; i++) val[i] = grid_get_value_deriv(g, x + i * g->dim, der + i * g->dim);
}
void orc_grid_get_value(void* p, long n, const double* x, double* val) {
  orc_grid* g = (orc_grid*)p;
  for (long i = 0; i < n; i++) val[i] = grid_get_value(g, x + i * g->dim);
}
void orc_grid_hist_add(void* p, long n, const double* x, const double* v) {
  orc_grid* g = (orc_grid*)p;
  for (long i = 0; i < n; i++) grid_add_value(g, x + i * g->dim, v[i]);
}
double orc_grid_expected_bias(void* p) { return grid_expected_bias((orc_grid*)p); }

/* ------------------------------------------------------------------ gaussian grid */

/* lib/gaussian_grid.h:378-435 (a11): McGovern-De Pablo + zero-force denominator tables */
static void gauss_set_boundary(orc_gauss* gg, const double* mn, const double* mx, const int* periodic) {
  int dim = gg->g.dim;
  gg->dirty = 0;
  for (int i = 0; i < dim; i++) {
    gg->bmin[i] = mn[i];
    gg->bmax[i] = mx[i];
    gg->bper[i] = periodic[i];
  }
  for (int i = 0; i < dim; i++) {
    if (gg->bper[i]) continue;
    if (!gg->bc_denom[i]) {
      gg->bc_denom[i] = (double*)malloc(sizeof(double) * ORC_BC_TABLE_SIZE);
      gg->bc_deriv[i] = (double*)malloc(sizeof(double) * ORC_BC_TABLE_SIZE);
    }
    double sg = gg->sigma[i], lo = gg->bmin[i], hi = gg->bmax[i];
    for (size_t j = 0; j < ORC_BC_TABLE_SIZE; j++) {
      double s = j * (hi - lo) / (ORC_BC_TABLE_SIZE - 1) + lo;
      double tmp1 = sqrt(M_PI) * sg / 2. * (erf((s - lo) / sg) + erf((hi - s) / sg));
      double den = tmp1;
      double tmp2 = sqrt(M_PI) * sg / 2. * erf((hi - lo) / sg);
      den += (tmp2 - tmp1) * orc_sigmoid((s - lo) / (ORC_BC_MAR * sg));
      den += (tmp2 - tmp1) * orc_sigmoid((hi - s) / (ORC_BC_MAR * sg));
      gg->bc_denom[i][j] = den;
      double tmp3 = 1. * (exp(-pow(s - lo, 2) / pow(sg, 2)) - exp(-pow(hi - s, 2) / pow(sg, 2)));
      double dd = tmp3;
      dd += (tmp2 - tmp1) * orc_sigmoid_dx((s - lo) / (ORC_BC_MAR * sg)) / (ORC_BC_MAR * sg) -
            tmp3 * orc_sigmoid((s - lo) / (ORC_BC_MAR * sg));
      dd += -(tmp2 - tmp1) * orc_sigmoid_dx((hi - s) / (ORC_BC_MAR * sg)) / (ORC_BC_MAR * sg) -
            tmp3 * orc_sigmoid((hi - s) / (ORC_BC_MAR * sg));
      gg->bc_deriv[i][j] = dd;
    }
  }
}

/* lib/gaussian_grid.h:559-569 (a12, T8) */
static void gauss_update_minigrid(orc_gauss* gg) {
  gg->minisize_total = 1;
  for (int i = 0; i < gg->g.dim; i++) {
    double dist = sqrt(2 * ORC_GAUSS_SUPPORT) * gg->sigma[i];
    gg->minisize[i] = (size_t)orc_int_floor(dist / gg->g.dx[i]);
    gg->minisize_total *= (2 * gg->minisize[i] + 1);
  }
}

/* lib/gaussian_grid.h:490-499 (inclusive on both sides, T4) */
static int gauss_in_bounds(const orc_gauss* gg, const double* x) {
  for (int i = 0; i < gg->g.dim; i++)
    if (x[i] < gg->bmin[i] || x[i] > gg->bmax[i]) return 0;
  return 1;
}

/* lib/gaussian_grid.h:504-541: nearest-image (not minimum-image) remap */
static void gauss_remap(const orc_gauss* gg, double* x) {
  const orc_grid* g = &gg->g;
  double dp[2];
  for (int i = 0; i < g->dim; i++) {
    if (x[i] < g->min[i] || x[i] > g->max[i]) {
      if (g->periodic[i]) {
        x[i] -= (g->max[i] - g->min[i]) * orc_int_floor((x[i] - g->min[i]) / (g->max[i] - g->min[i]));
      } else if (gg->bper[i]) {
        dp[0] = orc_round((g->min[i] - x[i]) / (gg->bmax[i] - gg->bmin[i])) * (gg->bmax[i] - gg->bmin[i]);
        dp[1] = orc_round((g->max[i] - x[i]) / (gg->bmax[i] - gg->bmin[i])) * (gg->bmax[i] - gg->bmin[i]);
        if (fabsl(g->min[i] - x[i] - dp[0]) < fabsl(g->max[i] - x[i] - dp[1]))
          x[i] += dp[0];
        else
          x[i] += dp[1];
      }
    }
  }
}

/* lib/gaussian_grid.h:571-630 (a13, T13): values only, corner combinations only */
static void gauss_duplicate_boundary(orc_gauss* gg) {
  orc_grid* g = &gg->g;
  int dim = g->dim;
  size_t min_i[3], max_i[3], index_outter[3], index_bound[3];
  grid_get_index(g, gg->bmin, min_i);
  grid_get_index(g, gg->bmax, max_i);
  for (int i = 0; i < dim; i++) {
    while (min_i[i] * g->dx[i] + g->min[i] < gg->bmin[i]) min_i[i] += 1;
    while (max_i[i] * g->dx[i] + g->min[i] > gg->bmax[i] || max_i[i] == (size_t)g->n[i]) max_i[i] -= 1;
  }
  size_t offset_size = 1;
  for (int i = 0; i < dim; i++) offset_size *= 4;
  for (size_t i = 0; i < offset_size; i++) {
    int b_flag = 0;
    size_t temp = i;
    for (int j = 0; j < dim; j++) {
      int off = (int)(temp % 4);
      temp = (temp - (size_t)off) / 4;
      switch (off) {
        case 0:
          b_flag |= gg->bper[j];
          b_flag |= (min_i[j] == 0);
          index_outter[j] = min_i[j] - 1;
          index_bound[j] = min_i[j];
          break;
        case 1:
          index_outter[j] = min_i[j];
          index_bound[j] = min_i[j];
          break;
        case 2:
          index_outter[j] = max_i[j];
          index_bound[j] = max_i[j];
          break;
        default:
          b_flag |= gg->bper[j];
          b_flag |= (max_i[j] == (size_t)(g->n[j] - 1));
          index_outter[j] = max_i[j] + 1;
          index_bound[j] = max_i[j];
          break;
      }
    }
    /* When the boundary reaches past the grid, max_i lands beyond grid_number_ and the reference's
     * copy falls into the unused tail of its DIM-fold over-allocation (lib/grid.h:897, T7): no
     * visible effect there, so such pairs are skipped here instead of written out of bounds. */
    for (int j = 0; j < dim; j++)
      if (index_outter[j] >= (size_t)g->n[j] || index_bound[j] >= (size_t)g->n[j]) b_flag = 1;
    if (!b_flag) g->grid[grid_multi2one(g, index_outter)] = g->grid[grid_multi2one(g, index_bound)];
  }
}

/* lib/gaussian_grid.h:176-372 (a10): hill deposit incl. McGDP / zero-force hills (T10-T14) */
static double gauss_add_value(orc_gauss* gg, const double* x0, double height) {
  orc_grid* g = &gg->g;
  int dim = g->dim;
  int index[3], x_index[3];
  long index1;
  double xx[3], x[3], dp[3], bc_force[3];
  size_t xx_index[3];
  double bias_added = 0, vol_element = 1;

  for (int i = 0; i < dim; i++) vol_element *= g->dx[i];
  for (int i = 0; i < dim; i++) x[i] = x0[i];
  gauss_remap(gg, x);
  for (int i = 0; i < dim; i++)
    if (!gg->bper[i] && (x[i] < gg->bmin[i] || x[i] > gg->bmax[i])) return 0;
  for (int i = 0; i < dim; i++) x_index[i] = orc_int_floor((x[i] - g->min[i]) / g->dx[i]);

  for (size_t i = 0; i < gg->minisize_total; i++) {
    int j;
    index1 = (long)i;
    for (j = 0; j < dim - 1; j++) {
      index[j] = (int)(index1 % (long)(2 * gg->minisize[j] + 1));
      index1 = (index1 - index[j]) / (long)(2 * gg->minisize[j] + 1);
    }
    index[j] = (int)index1;
    for (j = 0; j < dim; j++) index[j] -= (int)gg->minisize[j];

    int b_flag = 0;
    for (j = 0; j < dim; j++) {
      index[j] += x_index[j];
      if (index[j] >= g->n[j]) {
        if (g->periodic[j]) {
          index[j] %= g->n[j];
        } else {
          b_flag = 1;
          break;
        }
      }
      if (index[j] < 0) {
        if (g->periodic[j]) {
          index[j] += g->n[j];
        } else {
          b_flag = 1;
          break;
        }
      }
      xx_index[j] = (size_t)index[j];
      xx[j] = g->min[j] + g->dx[j] * xx_index[j];
      if (!gg->bper[j] && (xx[j] < gg->bmin[j] || xx[j] > gg->bmax[j])) {
        b_flag = 1;
        break;
      }
    }
    if (b_flag) continue;

    double dp2 = 0;
    for (j = 0; j < dim; j++) {
      dp[j] = xx[j] - x[j];
      if (g->periodic[j]) dp[j] -= orc_round(dp[j] / (g->max[j] - g->min[j])) * (g->max[j] - g->min[j]);
      dp[j] /= gg->sigma[j];
      dp2 += dp[j] * dp[j];
    }
    if (dp2 < ORC_GAUSS_SUPPORT) {
      double expo = exp(-dp2);
      double bc_denom = 1.0, bc_correction = 0;
      for (j = 0; j < dim; j++) {
        if (!gg->bper[j]) {
          /* T12: truncating index into the 65 536-entry tables */
          size_t bc_index =
              (size_t)((ORC_BC_TABLE_SIZE - 1) * (xx[j] - gg->bmin[j]) / (gg->bmax[j] - gg->bmin[j]));
          double sg = gg->sigma[j];
          double temp1 = exp(-pow(x[j] - gg->bmin[j], 2) / (pow(sg, 2)));
          double temp2 = orc_sigmoid((xx[j] - gg->bmin[j]) / (sg * ORC_BC_MAR));
          double temp3 = exp(-pow(x[j] - gg->bmax[j], 2) / (pow(sg, 2)));
          double temp4 = orc_sigmoid((gg->bmax[j] - xx[j]) / (sg * ORC_BC_MAR));
          bc_correction = (temp1 - expo) * temp2 + (temp3 - expo) * temp4; /* T11: overwritten per dim */
          bc_denom *= gg->bc_denom[j][bc_index];
          double temp5 = -2 * dp[j] / sg;
          double temp6 = orc_sigmoid_dx((xx[j] - gg->bmin[j]) / (sg * ORC_BC_MAR)) / (ORC_BC_MAR * sg);
          double temp7 = -orc_sigmoid_dx((gg->bmax[j] - xx[j]) / (sg * ORC_BC_MAR)) / (ORC_BC_MAR * sg);
          bc_force[j] = temp5 * expo;
          bc_force[j] += (temp1 - expo) * temp6 - temp5 * expo * temp2 + (temp3 - expo) * temp7 - temp5 * expo * temp4;
          bc_force[j] = bc_force[j] * bc_denom - gg->bc_deriv[j][bc_index] * (expo + bc_correction);
          bc_force[j] /= bc_denom * bc_denom;
          bc_correction /= bc_denom;
        } else {
          bc_denom *= sqrt(M_PI) * gg->sigma[j];
        }
      }
      expo /= bc_denom;
      size_t xx_index1 = grid_multi2one(g, xx_index);
      g->grid[xx_index1] += height * (expo + bc_correction);
      bias_added += height * (expo + bc_correction) * vol_element; /* T14 */
      for (j = 0; j < dim; j++) {
        if (gg->bper[j])
          g->deriv[xx_index1 * (size_t)dim + j] -= height * (2 * dp[j] / gg->sigma[j] * expo);
        else
          g->deriv[xx_index1 * (size_t)dim + j] += height * bc_force[j];
      }
      if (!gg->dirty && bc_correction * bc_correction > 0) gg->dirty = 1;
    }
  }
  if (gg->dirty) {
    gauss_duplicate_boundary(gg);
    gg->dirty = 0;
  }
  return bias_added;
}

/* lib/gaussian_grid.h:118-138 */
static double gauss_get_value_deriv(const orc_gauss* gg, const double* x, double* der) {
  double xx[3];
  for (int i = 0; i < gg->g.dim; i++) xx[i] = x[i];
  if (!gauss_in_bounds(gg, xx)) {
    gauss_remap(gg, xx);
    if (!gauss_in_bounds(gg, xx)) {
      for (int i = 0; i < gg->g.dim; i++) der[i] = 0;
      return 0;
    }
  }
  return grid_get_value_deriv(&gg->g, xx, der);
}

/* lib/gaussian_grid.h:99-116 */
static double gauss_get_value(const orc_gauss* gg, const double* x) {
  double xx[3];
  for (int i = 0; i < gg->g.dim; i++) xx[i] = x[i];
  if (!gauss_in_bounds(gg, xx)) {
    gauss_remap(gg, xx);
    if (!gauss_in_bounds(gg, xx)) return 0;
  }
  return grid_get_value(&gg->g, xx);
}

/* lib/gaussian_grid.h:65-80 (T8, T9) */
static void gauss_init(orc_gauss* gg, int dim, const double* mn, const double* mx, const double* spacing,
                       const int* periodic, int interp, const double* sigma) {
  memset(gg, 0, sizeof(*gg));
  grid_init(&gg->g, dim, mn, mx, spacing, periodic, 1, interp);
  for (int i = 0; i < dim; i++) gg->sigma[i] = sigma[i] * sqrt(2.);
  gauss_set_boundary(gg, mn, mx, periodic);
  gauss_update_minigrid(gg);
}

static void gauss_free(orc_gauss* gg) {
  grid_free(&gg->g);
  for (int i = 0; i < 3; i++) {
    free(gg->bc_denom[i]);
    free(gg->bc_deriv[i]);
  }
}

void* orc_gauss_create(int dim, const double* mn, const double* mx, const double* spacing,
                       const int* periodic, int interp, const double* sigma) {
  orc_gauss* gg = (orc_gauss*)malloc(sizeof(orc_gauss));
  gauss_init(gg, dim, mn, mx, spacing, periodic, interp, sigma);
  return gg;
}
void orc_gauss_destroy(void* p) {
  gauss_free((orc_gauss*)p);
  free(p);
}
void orc_gauss_set_boundary(void* p, const double* mn, const double* mx, const int* periodic) {
  gauss_set_boundary((orc_gauss*)p, mn, mx, periodic);
}
void orc_gauss_info(void* p, int* n, double* dx, double* mn, double* mx, int* mini) {
  orc_gauss* gg = (orc_gauss*)p;
  for (int i = 0; i < gg->g.dim; i++) {
    n[i] = gg->g.n[i];
    dx[i] = gg->g.dx[i];
    mn[i] = gg->g.min[i];
    mx[i] = gg->g.max[i];
    mini[i] = (int)gg->minisize[i];
  }
}
size_t orc_gauss_size(void* p) { return ((orc_gauss*)p)->g.size; }
void orc_gauss_get_arrays(void* p, double* v, double* d) { orc_grid_get_arrays(&((orc_gauss*)p)->g, v, d); }
void orc_gauss_set_arrays(void* p, double* v, double* d) { orc_grid_set_arrays(&((orc_gauss*)p)->g, v, d); }
void orc_gauss_tables(void* p, int dimi, double* denom, double* deriv) {
  orc_gauss* gg = (orc_gauss*)p;
  memcpy(denom, gg->bc_denom[dimi], sizeof(double) * ORC_BC_TABLE_SIZE);
  memcpy(deriv, gg->bc_deriv[dimi], sizeof(double) * ORC_BC_TABLE_SIZE);
}
double orc_gauss_add_value(void* p, const double* x, double height) {
  return gauss_add_value((orc_gauss*)p, x, height);
}
void orc_gauss_add_values(void* p, long n, const double* x, const double* height, double* bias_added) {
  orc_gauss* gg = (orc_gauss*)p;
  for (long i = 0; i < n; i++) {
    double ba = gauss_add_value(gg, x + i * gg->g.dim, height[i]);
    if (bias_added) bias_added[i] = ba;
  }
}
void orc_gauss_eval(void* p, long n, const double* x, double* val, double* der) {
  orc_gauss* gg = (orc_gauss*)p;
  for (long i = 0; i < n; i++) val[i] = gauss_get_value_deriv(gg, x + i * gg->g.dim, der + i * gg->g.dim);
}
void orc_gauss_get_value(void* p, long n, const double* x, double* val) {
  orc_gauss* gg = (orc_gauss*)p;
  for (long i = 0; i < n; i++) val[i] = gauss_get_value(gg, x + i * gg->g.dim);
}
void orc_gauss_remap(void* p, double* x) { gauss_remap((orc_gauss*)p, x); }
void orc_gauss_set_interpolation(void* p, int b) { ((orc_gauss*)p)->g.b_interpolate = b; }

/* ------------------------------------------------------------------ EDMBias */

struct orc_bias { /* lib/edm_bias.h:118-178 */
  int dim;
  int b_tempering, b_targeting;
  double global_tempering, bias_factor, boltzmann_factor, temperature;
  double hill_prefactor, bias_per_step, hill_density;
  double cum_bias, total_volume, expected_target;
  int b_outofbounds;
  double bias_dx[3], bias_sigma[3], min[3], max[3];
  int bper[3];
  orc_grid* target; /* borrowed */
  orc_gauss* bias;
  orc_grid* cv_hist;
  const int* mask;
  double temp_hill_cum, temp_hill_prefactor;
  int est_hill_count;
  int hills_added;
  long long steps;
  double* overflow; /* T19: zero-filled, with slack for the D=3 write past BIAS_BUFFER_DBLS */
  size_t overflow_left, overflow_right;
  int b_skip_hill_add;
  orc_hill_event* log;
  long log_n, log_cap;
  int log_on;
};

void* orc_bias_create(int dim, int b_tempering, double global_tempering, double bias_factor,
                      double hill_prefactor, double bias_per_step, double hill_density,
                      const double* dx, const double* sigma, const double* mn, const double* mx) {
  /* field defaults: lib/edm_bias.cpp:34-60; values: read_input, lib/edm_bias.cpp:1009-1051 */
  orc_bias* b = (orc_bias*)calloc(1, sizeof(orc_bias));
  b->dim = dim;
  b->b_tempering = b_tempering;
  b->global_tempering = global_tempering;
  b->bias_factor = bias_factor;
  b->hill_prefactor = hill_prefactor;
  b->bias_per_step = bias_per_step;
  b->hill_density = hill_density;
  b->temperature = -1.0;
  b->temp_hill_cum = -1;
  b->temp_hill_prefactor = -1;
  for (int i = 0; i < dim; i++) {
    b->bias_dx[i] = dx[i];
    b->bias_sigma[i] = sigma[i];
    b->min[i] = mn[i];
    b->max[i] = mx[i];
  }
  b->overflow = (double*)calloc(ORC_BUFFER_DBLS + 8, sizeof(double));
  b->log_on = 1;
  return b;
}

void orc_bias_destroy(void* p) {
  orc_bias* b = (orc_bias*)p;
  if (b->bias) orc_gauss_destroy(b->bias);
  if (b->cv_hist) orc_grid_destroy(b->cv_hist);
  free(b->overflow);
  free(b->log);
  free(b);
}

void orc_bias_set_target(void* p, void* target_grid, double expected_target) {
  orc_bias* b = (orc_bias*)p;
  b->target = (orc_grid*)target_grid;
  b->b_targeting = target_grid != NULL;
  b->expected_target = expected_target;
}

/* lib/edm_bias.cpp:264-269 */
void orc_bias_setup(void* p, double temperature, double boltz) {
  orc_bias* b = (orc_bias*)p;
  b->temperature = temperature;
  b->boltzmann_factor = boltz * temperature;
}

/* lib/edm_bias.cpp:98-222, EDM_SERIAL build (geometry only; T9) */
void orc_bias_subdivide(void* p, const double* sublo, const double* subhi, const double* boxlo,
                        const double* boxhi, const int* periodic, const double* skin) {
  orc_bias* b = (orc_bias*)p;
  if (b->bias != NULL) return;
  if (b->temperature < 0) {
    fprintf(stderr, "[oracle] Must call setup before subdivide\n");
    abort();
  }
  int grid_period[3] = {0, 0, 0};
  double mn[3], mx[3];
  int bounds_flag = 1;
  for (int i = 0; i < b->dim; i++) {
    b->bper[i] = 0;
    if (fabs(boxlo[i] - b->min[i]) < 0.000001 && fabs(boxhi[i] - b->max[i]) < 0.000001) b->bper[i] = periodic[i];
  }
  for (int i = 0; i < b->dim; i++) {
    mn[i] = sublo[i];
    mx[i] = subhi[i];
    if (fabs(sublo[i] - b->min[i]) < 0.000001 && fabs(subhi[i] - b->max[i]) < 0.000001) {
      grid_period[i] = periodic[i];
      bounds_flag = 0;
    } else {
      mn[i] -= skin[i];
      mx[i] += skin[i];
    }
    bounds_flag &= (mn[i] >= b->max[i] || mx[i] <= b->min[i]);
  }
  b->bias = (orc_gauss*)orc_gauss_create(b->dim, mn, mx, b->bias_dx, grid_period, 1, b->bias_sigma);
  b->cv_hist = (orc_grid*)orc_grid_create(b->dim, mn, mx, b->bias_sigma, grid_period, 0, 0);
  gauss_set_boundary(b->bias, b->min, b->max, b->bper);
  if (bounds_flag) {
    b->b_outofbounds = 1;
    return;
  }
  double vol = 1;
  for (int i = 0; i < b->dim; i++) vol *= b->bias->bmax[i] - b->bias->bmin[i]; /* gaussian_grid.h:437-444 */
  b->total_volume = 0;
  b->total_volume += vol;
}

void* orc_bias_gauss(void* p) { return ((orc_bias*)p)->bias; }
void* orc_bias_hist(void* p) { return ((orc_bias*)p)->cv_hist; }

void orc_bias_params(void* p, double* out) {
  orc_bias* b = (orc_bias*)p;
  out[0] = b->dim;
  out[1] = b->b_tempering;
  out[2] = b->b_targeting;
  out[3] = b->global_tempering;
  out[4] = b->bias_factor;
  out[5] = b->boltzmann_factor;
  out[6] = b->hill_prefactor;
  out[7] = b->bias_per_step;
  out[8] = b->hill_density;
  out[9] = b->cum_bias;
  out[10] = b->total_volume;
  out[11] = b->expected_target;
  out[12] = b->b_outofbounds;
  out[13] = (double)b->steps;
}
void orc_bias_set_cum_bias(void* p, double v) { ((orc_bias*)p)->cum_bias = v; }
void orc_bias_backlog(void* p, long* left, long* right, double* buffer) {
  orc_bias* b = (orc_bias*)p;
  *left = (long)b->overflow_left;
  *right = (long)b->overflow_right;
  if (buffer) memcpy(buffer, b->overflow, sizeof(double) * ORC_BUFFER_DBLS);
}
void orc_bias_set_mask(void* p, const int* mask) { ((orc_bias*)p)->mask = mask; }

/* lib/edm_bias.cpp:586-612 (a19, T21) */
static void bias_output_hill(orc_bias* b, const double* position, double height, double bias_added, int type) {
  if (b->log_on) {
    if (b->log_n == b->log_cap) {
      b->log_cap = b->log_cap ? 2 * b->log_cap : 1024;
      b->log = (orc_hill_event*)realloc(b->log, sizeof(orc_hill_event) * (size_t)b->log_cap);
    }
    orc_hill_event* e = &b->log[b->log_n++];
    memset(e, 0, sizeof(*e));
    e->steps = b->steps;
    e->type = type;
    e->hills_added = b->hills_added;
    for (int i = 0; i < b->dim; i++) e->pos[i] = position[i];
    e->height = height;
    e->bias_added = bias_added;
    e->cum_over_vol = b->cum_bias / b->total_volume;
  }
  if (type == 'n' || type == 'b' || type == 'h')
    grid_add_value(b->cv_hist, position, 1);
  else if (type == 'u' || type == 'v')
    grid_add_value(b->cv_hist, position, -1);
}

/* lib/edm_bias.cpp:276-295 (a7, T23) */
double orc_bias_update_forces(void* p, long n, const double* x, long xstride, double* f, long fstride,
                              int apply_mask) {
  orc_bias* b = (orc_bias*)p;
  if (b->b_outofbounds) return 0.0;
  double der[3] = {0, 0, 0};
  double energy = 0;
  for (long i = 0; i < n; i++) {
    if (apply_mask < 0 || (b->mask[i] & apply_mask)) {
      energy += gauss_get_value_deriv(b->bias, x + i * xstride, der);
      for (int j = 0; j < b->dim; j++) f[i * fstride + j] -= der[j];
    }
  }
  return energy;
}

/* lib/edm_bias.cpp:313-380 (a16, T20): drain the backlog up to max_bias */
static double bias_flush_bias_buffer(orc_bias* b, double max_bias) {
  int w = b->dim + 1;
  double bias_added = 0;
  for (; b->overflow_left < b->overflow_right; b->overflow_left++) {
    double* slot = &b->overflow[b->overflow_left * (size_t)w];
    double temp = gauss_add_value(b->bias, slot, slot[b->dim]);
    b->hills_added++;
    bias_added += temp;
    bias_output_hill(b, slot, slot[b->dim], temp, 'b');
    if (bias_added > max_bias) {
      double h = fmax(max_bias - bias_added, -slot[b->dim]);
      slot[b->dim] = -h;
      temp = gauss_add_value(b->bias, slot, h);
      bias_output_hill(b, slot, h, temp, 'v');
      b->hills_added++;
      bias_added += temp;
      break;
    }
  }
  if (b->overflow_left == b->overflow_right) b->overflow_left = b->overflow_right = 0;
  return bias_added;
}

/* lib/edm_bias.cpp:413-442 (T15 threshold tempering, T18 whole-round skip) */
void orc_bias_pre_add_hill(void* p, int est) {
  orc_bias* b = (orc_bias*)p;
  if (b->b_outofbounds) return;
  b->est_hill_count = est;
  b->temp_hill_prefactor = b->hill_prefactor;
  if (b->global_tempering > 0)
    if (b->cum_bias / b->total_volume >= b->global_tempering)
      b->temp_hill_prefactor *= exp(-(b->cum_bias / b->total_volume - b->global_tempering) /
                                    (b->global_tempering * (b->bias_factor - 1) * b->boltzmann_factor));
  b->temp_hill_cum = 0;
  b->hills_added = 0;
  b->temp_hill_cum += bias_flush_bias_buffer(b, b->bias_per_step);
  if (b->overflow_left == 0 && b->overflow_right == 0)
    b->b_skip_hill_add = 0;
  else
    b->b_skip_hill_add = 1;
}

/* lib/edm_bias.cpp:444-526 (a15, T19 off-by-one push, T20 undo height); serial build: no send buffer */
static double bias_do_add_hill(orc_bias* b, const double* position, double this_h) {
  int buffer_flag = 0;
  double bias_added = 0, temp_h;
  int w = b->dim + 1;
  if (b->temp_hill_cum < b->bias_per_step) {
    bias_added = gauss_add_value(b->bias, position, this_h);
    b->temp_hill_cum += bias_added;
    b->hills_added++;
    bias_output_hill(b, position, this_h, bias_added, 'h');
    if (b->temp_hill_cum > b->bias_per_step) {
      temp_h = fmax(b->bias_per_step - b->temp_hill_cum, -this_h);
      bias_added = gauss_add_value(b->bias, position, temp_h);
      b->hills_added++;
      bias_output_hill(b, position, temp_h, bias_added, 'u');
      b->temp_hill_cum += bias_added;
      buffer_flag = 1;
      this_h = -temp_h;
    }
  } else {
    bias_output_hill(b, position, 0, 0, 'h');
    buffer_flag = 1;
  }
  if (buffer_flag) {
    if (b->overflow_right == ORC_BUFFER_SLOTS) {
      if (b->overflow_left == 0) {
        fprintf(stderr, "[oracle] The bias overflow buffer is full\n");
        abort();
      } else {
        b->overflow_left--;
        for (int i = 0; i < b->dim; i++) b->overflow[b->overflow_left * (size_t)w + (size_t)i] = position[i];
        b->overflow[b->overflow_left * (size_t)w + (size_t)b->dim] = this_h;
      }
    } else {
      b->overflow_right++; /* incremented BEFORE the write: slot `right` is written, [left,right) is read */
      for (int i = 0; i < b->dim; i++) b->overflow[b->overflow_right * (size_t)w + (size_t)i] = position[i];
      b->overflow[b->overflow_right * (size_t)w + (size_t)b->dim] = this_h;
    }
  }
  return bias_added;
}

/* lib/edm_bias.cpp:528-563 (T17 acceptance, T15 local tempering) */
static void bias_add_hill(orc_bias* b, const double* position, double runiform) {
  if (b->temp_hill_prefactor < 0) {
    fprintf(stderr, "[oracle] Must call pre_add_hill before add_hill\n");
    abort();
  }
  if (b->b_skip_hill_add) return;
  double this_h = b->temp_hill_prefactor;
  if (!b->b_outofbounds) {
    if (b->hill_density < 0 || runiform < b->hill_density / b->est_hill_count) {
      if (b->b_targeting) this_h *= exp(grid_get_value(b->target, position) - b->expected_target);
      if (b->b_tempering && b->global_tempering < 0)
        this_h *= exp(-gauss_get_value(b->bias, position) / ((b->bias_factor - 1) * b->boltzmann_factor));
      if (b->hill_density < 0)
        this_h /= b->est_hill_count;
      else
        this_h /= b->hill_density;
      this_h = fmin(this_h, ORC_BIAS_CLAMP * b->bias_per_step);
      bias_do_add_hill(b, position, this_h);
    }
  }
}

void orc_bias_add_hill_many(void* p, long n, const double* x, const double* runiform) {
  orc_bias* b = (orc_bias*)p;
  for (long i = 0; i < n; i++) bias_add_hill(b, x + i * b->dim, runiform[i]);
}

/* lib/edm_bias.cpp:565-583, 922-931 (T22); serial build: flush_buffers is a no-op */
void orc_bias_post_add_hill(void* p) {
  orc_bias* b = (orc_bias*)p;
  b->cum_bias += b->temp_hill_cum;
  b->temp_hill_cum = -1;
  b->temp_hill_prefactor = -1;
  b->steps++;
}

/* lib/edm_bias.cpp:401-411 */
void orc_bias_add_hills(void* p, long n, const double* x, long xstride, const double* runiform, int apply_mask) {
  orc_bias* b = (orc_bias*)p;
  orc_bias_pre_add_hill(b, (int)n);
  for (long i = 0; i < n; i++)
    if (apply_mask < 0 || (apply_mask & b->mask[i])) bias_add_hill(b, x + i * xstride, runiform[i]);
  orc_bias_post_add_hill(b);
}

long orc_bias_log_size(void* p) { return ((orc_bias*)p)->log_n; }
void orc_bias_log_copy(void* p, orc_hill_event* out) {
  orc_bias* b = (orc_bias*)p;
  memcpy(out, b->log, sizeof(orc_hill_event) * (size_t)b->log_n);
}
void orc_bias_log_clear(void* p) { ((orc_bias*)p)->log_n = 0; }
void orc_bias_log_enable(void* p, int on) { ((orc_bias*)p)->log_on = on; }

/* ------------------------------------------------------------------ pair-distance CV */

/* lammps/fix_edm_pair.cpp:177-240 restated at the lib level: all pairs are
 * evaluated against the start-of-step bias, then the hills are proposed in pair order, two
 * per pair because both atoms are local (fix_edm_pair.cpp:230-236). */
double orc_pair_step(void* p, long npairs, const int* pi, const int* pj, const double* x, double* f,
                     const double* shift, int do_hills, int est, const double* uniforms, double* r_out) {
  orc_bias* b = (orc_bias*)p;
  double energy = 0;
  double* rr = (double*)malloc(sizeof(double) * (size_t)(npairs > 0 ? npairs : 1));
  for (long k = 0; k < npairs; k++) {
    int i = pi[k], j = pj[k];
    double delx = x[3 * i + 0] - x[3 * j + 0];
    double dely = x[3 * i + 1] - x[3 * j + 1];
    double delz = x[3 * i + 2] - x[3 * j + 2];
    if (shift) {
      delx -= shift[3 * k + 0];
      dely -= shift[3 * k + 1];
      delz -= shift[3 * k + 2];
    }
    double r = sqrt(delx * delx + dely * dely + delz * delz);
    double rinv = 1.0 / r;
    delx *= rinv;
    dely *= rinv;
    delz *= rinv;
    double der[3] = {0, 0, 0};
    double edm_force = 0;
    if (!b->b_outofbounds) { /* update_force, lib/edm_bias.cpp:297-311 */
      energy += gauss_get_value_deriv(b->bias, &r, der);
      edm_force -= der[0];
    }
    f[3 * i + 0] += delx * edm_force;
    f[3 * i + 1] += dely * edm_force;
    f[3 * i + 2] += delz * edm_force;
    f[3 * j + 0] -= delx * edm_force;
    f[3 * j + 1] -= dely * edm_force;
    f[3 * j + 2] -= delz * edm_force;
    rr[k] = r;
    if (r_out) r_out[k] = r;
  }
  if (do_hills) {
    orc_bias_pre_add_hill(b, est);
    for (long k = 0; k < npairs; k++) {
      bias_add_hill(b, &rr[k], uniforms[2 * k]);
      bias_add_hill(b, &rr[k], uniforms[2 * k + 1]);
    }
    orc_bias_post_add_hill(b);
  }
  free(rr);
  return energy;
}

/* The same with ghost atoms, as one rank of a decomposed LAMMPS run sees it (lammps/fix_edm_pair.cpp:177-236, newton
 * off): every listed pair has a LOCAL first atom i (< nlocal); j may be a ghost (>= nlocal): then no force on j (:223)
 * and one hill proposal instead of two (:233).  uniforms still holds two per pair; the second is unused for a ghost.
 * Lib-level order as above.  ncalls_out = hill proposals made (the caller's next last_calls, :245). */
double orc_pair_step_ghost(void* p, long npairs, const int* pi, const int* pj, const double* x, double* f,
                           const double* shift, long nlocal, int do_hills, int est, const double* uniforms,
                           double* r_out, long* ncalls_out) {
  orc_bias* b = (orc_bias*)p;
  double energy = 0;
  long ncalls = 0;
  double* rr = (double*)malloc(sizeof(double) * (size_t)(npairs > 0 ? npairs : 1));
  for (long k = 0; k < npairs; k++) {
    int i = pi[k], j = pj[k];
    double delx = x[3 * i + 0] - x[3 * j + 0];
    double dely = x[3 * i + 1] - x[3 * j + 1];
    double delz = x[3 * i + 2] - x[3 * j + 2];
    if (shift) {
      delx -= shift[3 * k + 0];
      dely -= shift[3 * k + 1];
      delz -= shift[3 * k + 2];
    }
    double r = sqrt(delx * delx + dely * dely + delz * delz);
    double rinv = 1.0 / r;
    delx *= rinv;
    dely *= rinv;
    delz *= rinv;
    double der[3] = {0, 0, 0};
    double edm_force = 0;
    if (!b->b_outofbounds) {
      energy += gauss_get_value_deriv(b->bias, &r, der);
      edm_force -= der[0];
    }
    f[3 * i + 0] += delx * edm_force;
    f[3 * i + 1] += dely * edm_force;
    f[3 * i + 2] += delz * edm_force;
    if (j < nlocal) {
      f[3 * j + 0] -= delx * edm_force;
      f[3 * j + 1] -= dely * edm_force;
      f[3 * j + 2] -= delz * edm_force;
    }
    rr[k] = r;
    if (r_out) r_out[k] = r;
  }
  if (do_hills) {
    orc_bias_pre_add_hill(b, est);
    for (long k = 0; k < npairs; k++) {
      bias_add_hill(b, &rr[k], uniforms[2 * k]);
      ncalls++;
      if (pj[k] < nlocal) {
        bias_add_hill(b, &rr[k], uniforms[2 * k + 1]);
        ncalls++;
      }
    }
    orc_bias_post_add_hill(b);
  }
  free(rr);
  if (ncalls_out) *ncalls_out = ncalls;
  return energy;
}

/* Stand-in for the LAMMPS half neighbour list (not part of the reference tree): all i<j with
 * minimum-image distance < cutoff, ordered by (i, j).  shift[k] is the image vector s such that
 * the pair separation is x[i] - x[j] - s (a ghost atom in LAMMPS terms). */
typedef struct { int i, j; double s[3]; } orc_pair_rec;
static int pair_cmp(const void* a, const void* b) {
  const orc_pair_rec* pa = (const orc_pair_rec*)a;
  const orc_pair_rec* pb = (const orc_pair_rec*)b;
  if (pa->i != pb->i) return pa->i < pb->i ? -1 : 1;
  if (pa->j != pb->j) return pa->j < pb->j ? -1 : 1;
  return 0;
}
long orc_build_half_list(long natoms, const double* x, const double* box, double cutoff, long cap, int* pi,
                         int* pj, double* shift) {
  int nc[3];
  double cs[3];
  for (int d = 0; d < 3; d++) {
    nc[d] = (int)floor(box[d] / cutoff);
    if (nc[d] < 1) nc[d] = 1;
    cs[d] = box[d] / nc[d];
  }
  long ncell = (long)nc[0] * nc[1] * nc[2];
  long* head = (long*)malloc(sizeof(long) * (size_t)(ncell + 1));
  long* next = (long*)malloc(sizeof(long) * (size_t)(natoms > 0 ? natoms : 1));
  for (long c = 0; c < ncell; c++) head[c] = -1;
  for (long a = natoms - 1; a >= 0; a--) {
    int c[3];
    for (int d = 0; d < 3; d++) {
      c[d] = (int)floor(x[3 * a + d] / cs[d]);
      if (c[d] < 0) c[d] = 0;
      if (c[d] >= nc[d]) c[d] = nc[d] - 1;
    }
    long ci = ((long)c[2] * nc[1] + c[1]) * nc[0] + c[0];
    next[a] = head[ci];
    head[ci] = a;
  }
  long cnt = 0, reccap = 1024;
  orc_pair_rec* rec = (orc_pair_rec*)malloc(sizeof(orc_pair_rec) * (size_t)reccap);
  double rc2 = cutoff * cutoff;
  for (long a = 0; a < natoms; a++) {
    int c[3];
    for (int d = 0; d < 3; d++) {
      c[d] = (int)floor(x[3 * a + d] / cs[d]);
      if (c[d] < 0) c[d] = 0;
      if (c[d] >= nc[d]) c[d] = nc[d] - 1;
    }
    /* visit each distinct neighbour cell once, even when the box has < 3 cells per side */
    int lo[3], hi[3];
    for (int d = 0; d < 3; d++) {
      lo[d] = nc[d] >= 3 ? -1 : 0;
      hi[d] = nc[d] >= 3 ? 1 : nc[d] - 1;
    }
    for (int dz = lo[2]; dz <= hi[2]; dz++)
      for (int dy = lo[1]; dy <= hi[1]; dy++)
        for (int dxx = lo[0]; dxx <= hi[0]; dxx++) {
          int q[3];
          if (nc[0] >= 3) q[0] = (c[0] + dxx + nc[0]) % nc[0]; else q[0] = dxx;
          if (nc[1] >= 3) q[1] = (c[1] + dy + nc[1]) % nc[1]; else q[1] = dy;
          if (nc[2] >= 3) q[2] = (c[2] + dz + nc[2]) % nc[2]; else q[2] = dz;
          long ci = ((long)q[2] * nc[1] + q[1]) * nc[0] + q[0];
          for (long bb = head[ci]; bb >= 0; bb = next[bb]) {
            if (bb <= a) continue;
            double s[3], d2 = 0;
            for (int d = 0; d < 3; d++) {
              double del = x[3 * a + d] - x[3 * bb + d];
              s[d] = box[d] * orc_round(del / box[d]);
              del -= s[d];
              d2 += del * del;
            }
            if (d2 < rc2) {
              if (cnt == reccap) {
                reccap *= 2;
                rec = (orc_pair_rec*)realloc(rec, sizeof(orc_pair_rec) * (size_t)reccap);
              }
              rec[cnt].i = (int)a;
              rec[cnt].j = (int)bb;
              rec[cnt].s[0] = s[0];
              rec[cnt].s[1] = s[1];
              rec[cnt].s[2] = s[2];
              cnt++;
            }
          }
        }
  }
  qsort(rec, (size_t)cnt, sizeof(orc_pair_rec), pair_cmp);
  for (long k = 0; k < cnt && k < cap; k++) {
    pi[k] = rec[k].i;
    pj[k] = rec[k].j;
    if (shift) {
      shift[3 * k + 0] = rec[k].s[0];
      shift[3 * k + 1] = rec[k].s[1];
      shift[3 * k + 2] = rec[k].s[2];
    }
  }
  free(rec);
  free(head);
  free(next);
  return cnt;
}

/* ------------------------------------------------------------------ counter-based uniforms */

static uint64_t orc_mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
double orc_uniform(unsigned long long seed, unsigned long long step, unsigned long long counter) {
  uint64_t key = orc_mix64(seed ^ orc_mix64(step + 0x9E3779B97F4A7C15ULL));
  uint64_t bits = orc_mix64(key + counter * 0x9E3779B97F4A7C15ULL);
  return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
}
void orc_uniform_fill(unsigned long long seed, unsigned long long step, unsigned long long first, long n,
                      double* out) {
  for (long i = 0; i < n; i++) out[i] = orc_uniform(seed, step, first + (unsigned long long)i);
}

/* pair proposals: one hash per pair, 32-bit resolution per proposal (edm_uniform_pair) */
double orc_uniform_pair(unsigned long long seed, unsigned long long step, unsigned long long pairkey, int which) {
  uint64_t key = orc_mix64(seed ^ orc_mix64(step + 0x9E3779B97F4A7C15ULL));
  uint64_t bits = orc_mix64(key + pairkey * 0x9E3779B97F4A7C15ULL);
  uint64_t half = which == 0 ? (bits >> 32) : (bits & 0xffffffffULL);
  return (double)half * (1.0 / 4294967296.0);
}
void orc_uniform_pair_fill(unsigned long long seed, unsigned long long step, long n,
                           const unsigned long long* pairkeys, double* out) {
  for (long i = 0; i < n; i++) {
    out[2 * i] = orc_uniform_pair(seed, step, pairkeys[i], 0);
    out[2 * i + 1] = orc_uniform_pair(seed, step, pairkeys[i], 1);
  }
}

/* ------------------------------------------------------------------ bench-size neighbour list

 * The same half list as orc_build_half_list (all i<j with minimum-image distance < cutoff, rows ascending in
 * i), built fast enough for the 10^6-atom benchmark configuration: atoms counting-sorted into cells, each
 * row atom walks its 27 neighbour cells over contiguous cell-sorted coordinates.  The image shift is
 * returned as three int8 codes per pair (shift = code * box) instead of three doubles.  Stand-in for the
 * LAMMPS NeighList (half, newton off, every atom local), which is outside the reference tree. */
typedef struct { long n, cap; int* pi; int* pj; signed char* img; } orc_half_list;

void* orc_half_list_build(long natoms, const double* x, const double* box, double cutoff) {
  int nc[3];
  double cs[3];
  for (int d = 0; d < 3; d++) {
    nc[d] = (int)floor(box[d] / cutoff);
    if (nc[d] < 3) return NULL; /* small boxes: use orc_build_half_list */
    cs[d] = box[d] / nc[d];
  }
  const long ncell = (long)nc[0] * nc[1] * nc[2];
  long* start = (long*)calloc((size_t)ncell + 1, sizeof(long));
  int* cell_of = (int*)malloc(sizeof(int) * (size_t)natoms);
  for (long a = 0; a < natoms; a++) {
    int c[3];
    for (int d = 0; d < 3; d++) {
      c[d] = (int)floor(x[3 * a + d] / cs[d]);
      if (c[d] < 0) c[d] = 0;
      if (c[d] >= nc[d]) c[d] = nc[d] - 1;
    }
    cell_of[a] = (int)(((long)c[2] * nc[1] + c[1]) * nc[0] + c[0]);
    start[cell_of[a] + 1]++;
  }
  for (long c = 0; c < ncell; c++) start[c + 1] += start[c];
  long* fill = (long*)malloc(sizeof(long) * (size_t)ncell);
  memcpy(fill, start, sizeof(long) * (size_t)ncell);
  double* xs = (double*)malloc(sizeof(double) * 3 * (size_t)natoms);
  int* idx = (int*)malloc(sizeof(int) * (size_t)natoms);
  for (long a = 0; a < natoms; a++) { /* ascending a: slots inside a cell are in ascending atom index */
    const long s = fill[cell_of[a]]++;
    idx[s] = (int)a;
    xs[3 * s + 0] = x[3 * a + 0];
    xs[3 * s + 1] = x[3 * a + 1];
    xs[3 * s + 2] = x[3 * a + 2];
  }
  orc_half_list* hl = (orc_half_list*)calloc(1, sizeof(orc_half_list));
  hl->cap = 32 * natoms + 1024;
  hl->pi = (int*)malloc(sizeof(int) * (size_t)hl->cap);
  hl->pj = (int*)malloc(sizeof(int) * (size_t)hl->cap);
  hl->img = (signed char*)malloc(3 * (size_t)hl->cap);
  const double rc2 = cutoff * cutoff;
  for (long a = 0; a < natoms; a++) {
    const int ca = cell_of[a];
    const int c0 = ca % nc[0], c1 = (ca / nc[0]) % nc[1], c2 = ca / (nc[0] * nc[1]);
    const double ax = x[3 * a + 0], ay = x[3 * a + 1], az = x[3 * a + 2];
    for (int dz = -1; dz <= 1; dz++)
      for (int dy = -1; dy <= 1; dy++)
        for (int dxx = -1; dxx <= 1; dxx++) {
          int q0 = c0 + dxx, q1 = c1 + dy, q2 = c2 + dz;
          signed char i0 = 0, i1 = 0, i2 = 0; /* x[a] - x[b] - img*box is the minimum image */
          if (q0 < 0) { q0 += nc[0]; i0 = -1; } else if (q0 >= nc[0]) { q0 -= nc[0]; i0 = 1; }
          if (q1 < 0) { q1 += nc[1]; i1 = -1; } else if (q1 >= nc[1]) { q1 -= nc[1]; i1 = 1; }
          if (q2 < 0) { q2 += nc[2]; i2 = -1; } else if (q2 >= nc[2]) { q2 -= nc[2]; i2 = 1; }
          /* a partner reached across the upper face sits at x[b] + box: x[a] - x[b] - (+box) is the separation */
          const double s0 = i0 * box[0], s1 = i1 * box[1], s2 = i2 * box[2];
          const long q = ((long)q2 * nc[1] + q1) * nc[0] + q0;
          for (long sidx = start[q]; sidx < start[q + 1]; sidx++) {
            const int bb = idx[sidx];
            if (bb <= a) continue;
            const double d0 = (ax - xs[3 * sidx + 0]) - s0, d1 = (ay - xs[3 * sidx + 1]) - s1;
            const double d2v = (az - xs[3 * sidx + 2]) - s2;
            const double d2 = d0 * d0 + d1 * d1 + d2v * d2v;
            if (d2 < rc2) {
              if (hl->n == hl->cap) {
                hl->cap *= 2;
                hl->pi = (int*)realloc(hl->pi, sizeof(int) * (size_t)hl->cap);
                hl->pj = (int*)realloc(hl->pj, sizeof(int) * (size_t)hl->cap);
                hl->img = (signed char*)realloc(hl->img, 3 * (size_t)hl->cap);
              }
              hl->pi[hl->n] = (int)a;
              hl->pj[hl->n] = bb;
              hl->img[3 * hl->n + 0] = i0;
              hl->img[3 * hl->n + 1] = i1;
              hl->img[3 * hl->n + 2] = i2;
              hl->n++;
            }
          }
        }
  }
  free(start); free(cell_of); free(fill); free(xs); free(idx);
  return hl;
}
long orc_half_list_size(void* p) { return p ? ((orc_half_list*)p)->n : -1; }
void orc_half_list_copy(void* p, int* pi, int* pj, signed char* img) {
  orc_half_list* hl = (orc_half_list*)p;
  memcpy(pi, hl->pi, sizeof(int) * (size_t)hl->n);
  memcpy(pj, hl->pj, sizeof(int) * (size_t)hl->n);
  memcpy(img, hl->img, 3 * (size_t)hl->n);
}
void orc_half_list_free(void* p) {
  orc_half_list* hl = (orc_half_list*)p;
  if (!hl) return;
  free(hl->pi); free(hl->pj); free(hl->img); free(hl);
}

/* ------------------------------------------------------------------ CPU-baseline timers */

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
/* FixEDMPair::post_force's pair loop, lammps/fix_edm_pair.cpp:173-247, in the REFERENCE'S OWN order — per pair:
 * r = sqrt(del^2), rinv, update_force at r, +-force scatter into f[i] and f[j] (both local), and on a hill step
 * two add_hill calls (interleaved with the evaluations, T24) between pre_add_hill(last_calls) and post_add_hill.
 * Timed here so that marshalling stays outside the measured region; the uniforms come from a counter hash in
 * place of LAMMPS' RanMars (one draw per add_hill, as there).  Returns seconds. */
double orc_time_fix_pair(void* p, long npairs, const int* pi, const int* pj, const signed char* img,
                         const double* box, const double* x, double* f, int do_hills, int est,
                         unsigned long long seed, unsigned long long step, double* energy_out, long* ncalls_out) {
  orc_bias* b = (orc_bias*)p;
  const uint64_t key = orc_mix64(seed ^ orc_mix64(step + 0x9E3779B97F4A7C15ULL));
  uint64_t ctr = 0;
  double energy = 0;
  long ncalls = 0;
  const double t0 = now_s();
  if (do_hills) orc_bias_pre_add_hill(b, est);
  for (long k = 0; k < npairs; k++) {
    const int i = pi[k], j = pj[k];
    double delx = x[3 * i + 0] - x[3 * j + 0] - img[3 * k + 0] * box[0];
    double dely = x[3 * i + 1] - x[3 * j + 1] - img[3 * k + 1] * box[1];
    double delz = x[3 * i + 2] - x[3 * j + 2] - img[3 * k + 2] * box[2];
    double r = sqrt(delx * delx + dely * dely + delz * delz);
    const double rinv = 1.0 / r;
    delx *= rinv;
    dely *= rinv;
    delz *= rinv;
    double der[3] = {0, 0, 0};
    double edm_force = 0;
    if (!b->b_outofbounds) {
      energy += gauss_get_value_deriv(b->bias, &r, der);
      edm_force -= der[0];
    }
    f[3 * i + 0] += delx * edm_force;
    f[3 * i + 1] += dely * edm_force;
    f[3 * i + 2] += delz * edm_force;
    f[3 * j + 0] -= delx * edm_force;
    f[3 * j + 1] -= dely * edm_force;
    f[3 * j + 2] -= delz * edm_force;
    if (do_hills) {
      for (int w = 0; w < 2; w++) {
        const uint64_t bits = orc_mix64(key + (ctr++) * 0x9E3779B97F4A7C15ULL);
        bias_add_hill(b, &r, (double)(bits >> 11) * (1.0 / 9007199254740992.0));
        ncalls++;
      }
    }
  }
  if (do_hills) orc_bias_post_add_hill(b);
  const double t1 = now_s();
  if (energy_out) *energy_out = energy;
  if (ncalls_out) *ncalls_out = ncalls;
  return t1 - t0;
}
double orc_time_pair_eval(void* p, long npairs, const double* r, int repeats) {
  orc_bias* b = (orc_bias*)p;
  volatile double sink = 0;
  double t0 = now_s();
  for (int it = 0; it < repeats; it++)
    for (long k = 0; k < npairs; k++) {
      double der[3] = {0, 0, 0};
      sink += gauss_get_value_deriv(b->bias, &r[k], der) - der[0];
    }
  return now_s() - t0;
}
double orc_time_add_values(void* p, long n, const double* x, const double* h) {
  orc_gauss* gg = (orc_gauss*)p;
  volatile double sink = 0;
  double t0 = now_s();
  for (long i = 0; i < n; i++) sink += gauss_add_value(gg, x + i * gg->g.dim, h[i]);
  return now_s() - t0;
}
