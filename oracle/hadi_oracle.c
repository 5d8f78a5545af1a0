/* TEST INFRASTRUCTURE — see hadi_oracle.h.  Plain-C restatement of the reference's Heston ADI hot
 * path (Douglas device path, Craig-Sneyd host path, FD Jacobian, LM update and loop).
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (no FMA contraction, no -march=native) — SURVEY §8(c).
 * All expressions keep the reference's left-to-right evaluation order; comments give file:line
 * relative to /root/reference.
 */
#include "hadi_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ grids */

static int cmp_double(const void *a, const void *b) {
  const double x = *(const double *)a, y = *(const double *)b;
  return (x > y) - (x < y);
}

/* src/grid.cpp:26-42: s_i = K + c sinh(xi_i); append S0, sort, drop the largest node. */
void ho_grid_s(int m1, double S, double S0, double K, double c, double *s, double *ds) {
  double *tmp = (double *)malloc(sizeof(double) * (size_t)(m1 + 2));
  const double lo = asinh(-K / c);
  const double dxi = (1.0 / m1) * (asinh((S - K) / c) - asinh(-K / c));
  for (int i = 0; i <= m1; ++i) {
    const double xi = lo + i * dxi;
    tmp[i] = K + c * sinh(xi);
  }
  tmp[m1 + 1] = S0;
  qsort(tmp, (size_t)(m1 + 2), sizeof(double), cmp_double);
  for (int i = 0; i <= m1; ++i) s[i] = tmp[i];
  for (int i = 0; i < m1; ++i) ds[i] = s[i + 1] - s[i];
  free(tmp);
}

/* src/grid.cpp:44-61 and src/grid_pod.hpp:25-73: v_j = d sinh(j * d_eta); append V0, sort, drop last. */
void ho_grid_v(int m2, double V, double V0, double d, double *v, double *dv) {
  double *tmp = (double *)malloc(sizeof(double) * (size_t)(m2 + 2));
  const double deta = (1.0 / m2) * asinh(V / d);
  for (int j = 0; j <= m2; ++j) {
    const double xi = j * deta;
    tmp[j] = d * sinh(xi);
  }
  tmp[m2 + 1] = V0;
  qsort(tmp, (size_t)(m2 + 2), sizeof(double), cmp_double);
  for (int j = 0; j <= m2; ++j) v[j] = tmp[j];
  for (int j = 0; j < m2; ++j) dv[j] = v[j + 1] - v[j];
  free(tmp);
}

/* first index with |x_i - x0| < 1e-10, else -1 (src/jacobian_computation.cpp:275-281). */
int ho_find_index(const double *x, int n, double x0) {
  for (int i = 0; i < n; ++i)
    if (fabs(x[i] - x0) < 1e-10) return i;
  return -1;
}

/* ------------------------------------------------------------------ FD weights (src/coeff.hpp:25-127) */

static double w_delta(const double *D, int i, int pos) { /* second derivative, central */
  if (pos == -1) return 2 / (D[i] * (D[i] + D[i + 1]));
  if (pos == 0) return -2 / (D[i] * D[i + 1]);
  if (pos == 1) return 2 / (D[i + 1] * (D[i] + D[i + 1]));
  return 0.0;
}
static double w_beta(const double *D, int i, int pos) { /* first derivative, central */
  if (pos == -1) return -D[i + 1] / (D[i] * (D[i] + D[i + 1]));
  if (pos == 0) return (D[i + 1] - D[i]) / (D[i] * D[i + 1]);
  if (pos == 1) return D[i] / (D[i + 1] * (D[i] + D[i + 1]));
  return 0.0;
}
static double w_alpha(const double *D, int i, int pos) { /* first derivative, backward one-sided */
  if (pos == -2) return D[i] / (D[i - 1] * (D[i - 1] + D[i]));
  if (pos == -1) return (-D[i - 1] - D[i]) / (D[i - 1] * D[i]);
  if (pos == 0) return (D[i - 1] + 2 * D[i]) / (D[i] * (D[i - 1] + D[i]));
  return 0.0;
}
static double w_gamma(const double *D, int i, int pos) { /* first derivative, forward one-sided */
  if (pos == 0) return (-2 * D[i + 1] - D[i + 2]) / (D[i + 1] * (D[i + 1] + D[i + 2]));
  if (pos == 1) return (D[i + 1] + D[i + 2]) / (D[i + 1] * D[i + 2]);
  if (pos == 2) return -D[i + 1] / (D[i + 2] * (D[i + 1] + D[i + 2]));
  return 0.0;
}

/* ------------------------------------------------------------------ operators */

typedef struct {
  int m1, m2, P;
  double *s, *ds, *v, *dv;
  /* A0: src/hes_a0_kernels.hpp:30-55 — values[j][9 i + 3(l+1) + (k+1)] */
  double *a0;
  /* A1: src/hes_a1_kernels.hpp:51-107 — per v-row tridiagonals, explicit and (I - theta dt A) */
  double *a1l, *a1m, *a1u, *i1l, *i1m, *i1u, *piv;
  /* A2: src/hes_a2_shuffled_kernels.hpp:103-176 — pentadiagonal in v; identical for every s-column,
   * so ONE copy is kept (the reference stores m1+1 identical copies). */
  double *l2, *l1, *d0, *u1, *u2, *il2, *il1, *id0, *iu1, *iu2, *cp, *c2p, *dp;
  /* boundary vectors: src/hes_boundary_kernels.hpp:41-75 */
  double *b, *b1, *b2;
  /* work */
  double *R0, *R1, *R2, *Y0, *Y1, *Ut, *lam, *U0;
} ho_ws;

static double *dalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

static ho_ws *ws_new(int m1, int m2) {
  ho_ws *w = (ho_ws *)calloc(1, sizeof(ho_ws));
  const size_t P = (size_t)(m1 + 1) * (size_t)(m2 + 1);
  w->m1 = m1;
  w->m2 = m2;
  w->P = (int)P;
  w->s = dalloc(m1 + 1); w->ds = dalloc(m1); w->v = dalloc(m2 + 1); w->dv = dalloc(m2);
  w->a0 = dalloc((size_t)(m2 - 1) * (size_t)(m1 - 1) * 9);
  w->a1l = dalloc(P); w->a1m = dalloc(P); w->a1u = dalloc(P);
  w->i1l = dalloc(P); w->i1m = dalloc(P); w->i1u = dalloc(P); w->piv = dalloc(P);
  const size_t n2 = (size_t)m2 + 1;
  w->l2 = dalloc(n2); w->l1 = dalloc(n2); w->d0 = dalloc(n2); w->u1 = dalloc(n2); w->u2 = dalloc(n2);
  w->il2 = dalloc(n2); w->il1 = dalloc(n2); w->id0 = dalloc(n2); w->iu1 = dalloc(n2); w->iu2 = dalloc(n2);
  w->cp = dalloc(n2); w->c2p = dalloc(n2); w->dp = dalloc(n2);
  w->b = dalloc(P); w->b1 = dalloc(P); w->b2 = dalloc(P);
  w->R0 = dalloc(P); w->R1 = dalloc(P); w->R2 = dalloc(P); w->Y0 = dalloc(P); w->Y1 = dalloc(P);
  w->Ut = dalloc(P); w->lam = dalloc(P); w->U0 = dalloc(P);
  return w;
}

static void ws_free(ho_ws *w) {
  double **all[] = {&w->s, &w->ds, &w->v, &w->dv, &w->a0, &w->a1l, &w->a1m, &w->a1u, &w->i1l, &w->i1m,
                    &w->i1u, &w->piv, &w->l2, &w->l1, &w->d0, &w->u1, &w->u2, &w->il2, &w->il1, &w->id0,
                    &w->iu1, &w->iu2, &w->cp, &w->c2p, &w->dp, &w->b, &w->b1, &w->b2, &w->R0, &w->R1,
                    &w->R2, &w->Y0, &w->Y1, &w->Ut, &w->lam, &w->U0};
  for (size_t k = 0; k < sizeof(all) / sizeof(all[0]); ++k) free(*all[k]);
  free(w);
}

/* src/hes_a0_kernels.hpp:30-55 */
static void build_a0(ho_ws *w, double rho, double sigma) {
  const int m1 = w->m1, m2 = w->m2;
  for (int j = 0; j < m2 - 1; ++j)
    for (int i = 0; i < m1 - 1; ++i) {
      const double c = rho * sigma * w->s[i + 1] * w->v[j + 1];
      for (int l = -1; l <= 1; ++l)
        for (int k = -1; k <= 1; ++k)
          w->a0[(size_t)j * (size_t)(m1 - 1) * 9 + (size_t)i * 9 + (size_t)((l + 1) * 3 + (k + 1))] =
              c * w_beta(w->ds, i, k) * w_beta(w->dv, j, l);
    }
}

/* src/hes_a0_kernels.hpp:59-94 */
static void mul_a0(const ho_ws *w, const double *x, double *r) {
  const int m1 = w->m1, m2 = w->m2, ld = m1 + 1;
  for (int p = 0; p < w->P; ++p) r[p] = 0.0;
  for (int j = 0; j < m2 - 1; ++j)
    for (int i = 0; i < m1 - 1; ++i) {
      const double *val = w->a0 + (size_t)j * (size_t)(m1 - 1) * 9 + (size_t)i * 9;
      double sum = 0.0;
      for (int l = -1; l <= 1; ++l)
        for (int k = -1; k <= 1; ++k) sum += val[(l + 1) * 3 + (k + 1)] * x[(i + 1 + k) + (j + 1 + l) * ld];
      r[(j + 1) * ld + (i + 1)] = sum;
    }
}

/* src/hes_a1_kernels.hpp:51-107 */
static void build_a1(ho_ws *w, double r_d, double r_f, double theta, double dt) {
  const int m1 = w->m1, m2 = w->m2, ld = m1 + 1;
  for (int j = 0; j <= m2; ++j) {
    double *L = w->a1l + j * ld, *M = w->a1m + j * ld, *U = w->a1u + j * ld;
    double *iL = w->i1l + j * ld, *iM = w->i1m + j * ld, *iU = w->i1u + j * ld;
    M[0] = 0.0;
    iM[0] = 1.0;
    U[0] = 0.0;
    iU[0] = 0.0;
    for (int i = 1; i < m1; ++i) {
      const double s = w->s[i], v = w->v[j];
      const double a = 0.5 * s * s * v;
      const double b = (r_d - r_f) * s;
      L[i - 1] = a * w_delta(w->ds, i - 1, -1) + b * w_beta(w->ds, i - 1, -1);
      M[i] = a * w_delta(w->ds, i - 1, 0) + b * w_beta(w->ds, i - 1, 0) - 0.5 * r_d;
      U[i] = a * w_delta(w->ds, i - 1, 1) + b * w_beta(w->ds, i - 1, 1);
      iL[i - 1] = -theta * dt * L[i - 1];
      iM[i] = 1.0 - theta * dt * M[i];
      iU[i] = -theta * dt * U[i];
    }
    M[m1] = -0.5 * r_d;
    iM[m1] = 1.0 - theta * dt * M[m1];
    L[m1 - 1] = 0.0;
    iL[m1 - 1] = 0.0;
  }
}

/* src/hes_a1_kernels.hpp:111-135 (device order: lower, main, upper) */
static void mul_a1(const ho_ws *w, const double *x, double *r) {
  const int m1 = w->m1, m2 = w->m2, ld = m1 + 1;
  for (int j = 0; j <= m2; ++j) {
    const double *L = w->a1l + j * ld, *M = w->a1m + j * ld, *U = w->a1u + j * ld;
    const double *xx = x + j * ld;
    double *rr = r + j * ld;
    double sum = M[0] * xx[0];
    sum += U[0] * xx[1];
    rr[0] = sum;
    for (int i = 1; i < m1; ++i) rr[i] = L[i - 1] * xx[i - 1] + M[i] * xx[i] + U[i] * xx[i + 1];
    rr[m1] = L[m1 - 1] * xx[m1 - 1] + M[m1] * xx[m1];
  }
}

/* host-class order (main, lower, upper): src/hes_mat_fac.hpp:222-245 — used by Craig-Sneyd only */
static void mul_a1_host(const ho_ws *w, const double *x, double *r) {
  const int m1 = w->m1, m2 = w->m2, ld = m1 + 1;
  for (int j = 0; j <= m2; ++j) {
    const double *L = w->a1l + j * ld, *M = w->a1m + j * ld, *U = w->a1u + j * ld;
    const double *xx = x + j * ld;
    double *rr = r + j * ld;
    for (int i = 0; i <= m1; ++i) {
      double sum = M[i] * xx[i];
      if (i > 0) sum += L[i - 1] * xx[i - 1];
      if (i < m1) sum += U[i] * xx[i + 1];
      rr[i] = sum;
    }
  }
}

/* Thomas, pivots recomputed on every call: src/hes_a1_kernels.hpp:139-161.  x may alias b. */
static void solve_a1(ho_ws *w, double *x, const double *b) {
  const int m1 = w->m1, m2 = w->m2, ld = m1 + 1;
  for (int j = 0; j <= m2; ++j) {
    const double *iL = w->i1l + j * ld, *iM = w->i1m + j * ld, *iU = w->i1u + j * ld;
    double *t = w->piv + j * ld, *xx = x + j * ld;
    const double *bb = b + j * ld;
    t[0] = iM[0];
    xx[0] = bb[0];
    for (int i = 1; i <= m1; ++i) {
      const double m = iL[i - 1] / t[i - 1];
      t[i] = iM[i] - m * iU[i - 1];
      xx[i] = bb[i] - m * xx[i - 1];
    }
    xx[m1] /= t[m1];
    for (int i = m1 - 1; i >= 0; --i) xx[i] = (xx[i] - iU[i] * xx[i + 1]) / t[i];
  }
}

/* src/hes_a2_shuffled_kernels.hpp:103-176 (the upwind branch writes to j+1-shifted slots and the
 * central stencil is still added: SURVEY quirk Q2). */
static void build_a2(ho_ws *w, double r_d, double kappa, double eta, double sigma, double theta, double dt) {
  const int m2 = w->m2;
  const double *v = w->v, *D = w->dv;
  for (int j = 0; j <= m2; ++j) w->l2[j] = w->l1[j] = w->d0[j] = w->u1[j] = w->u2[j] = 0.0;
  for (int j = 0; j < m2 - 1; ++j) {
    const double temp = kappa * (eta - v[j]);
    const double temp2 = 0.5 * sigma * sigma * v[j];
    w->d0[j] += -0.5 * r_d;
    if (v[j] > 1.0) {
      w->l2[j + 1 - 2] += temp * w_alpha(D, j, -2);
      w->l1[j + 1 - 1] += temp * w_alpha(D, j, -1);
      w->d0[j + 1 - 0] += temp * w_alpha(D, j, 0);
      w->l1[j + 1 - 1] += temp2 * w_delta(D, j - 1, -1);
      w->d0[j + 1 + 0] += temp2 * w_delta(D, j - 1, 0);
      w->u1[j + 1] += temp2 * w_delta(D, j - 1, 1);
    }
    if (j == 0) {
      w->d0[j] += temp * w_gamma(D, j, 0);
      w->u1[j] += temp * w_gamma(D, j, 1);
      w->u2[j] += temp * w_gamma(D, j, 2);
    } else {
      w->l1[j - 1] += temp * w_beta(D, j - 1, -1) + temp2 * w_delta(D, j - 1, -1);
      w->d0[j] += temp * w_beta(D, j - 1, 0) + temp2 * w_delta(D, j - 1, 0);
      w->u1[j] += temp * w_beta(D, j - 1, 1) + temp2 * w_delta(D, j - 1, 1);
    }
  }
  for (int j = 0; j <= m2; ++j) w->id0[j] = 1.0 - theta * dt * w->d0[j];
  for (int j = 0; j < m2; ++j) {
    w->il1[j] = -theta * dt * w->l1[j];
    w->iu1[j] = -theta * dt * w->u1[j];
  }
  for (int j = 0; j < m2 - 1; ++j) {
    w->il2[j] = -theta * dt * w->l2[j];
    w->iu2[j] = -theta * dt * w->u2[j];
  }
}

/* src/hes_a2_shuffled_kernels.hpp:180-239, applied in natural layout (column i, stride ld). */
static void mul_a2(const ho_ws *w, const double *x, double *r) {
  const int m1 = w->m1, m2 = w->m2, ld = m1 + 1;
  const double *l2 = w->l2, *l1 = w->l1, *d0 = w->d0, *u1 = w->u1, *u2 = w->u2;
  for (int i = 0; i <= m1; ++i) {
#define X(j) x[(j) * ld + i]
#define R(j) r[(j) * ld + i]
    R(0) = d0[0] * X(0);
    if (0 < m2) R(0) += u1[0] * X(1);
    if (1 < m2) R(0) += u2[0] * X(2);
    if (0 < m2) {
      R(1) = l1[0] * X(0) + d0[1] * X(1);
      if (1 < m2) R(1) += u1[1] * X(2);
      if (2 < m2) R(1) += u2[1] * X(3);
    }
    for (int j = 2; j < m2 - 1; ++j) {
      R(j) = l2[j - 2] * X(j - 2) + l1[j - 1] * X(j - 1) + d0[j] * X(j) + u1[j] * X(j + 1);
      if (j < m2 - 2) R(j) += u2[j] * X(j + 2);
    }
    if (m2 > 2) {
      const int j = m2 - 1;
      R(j) = l2[j - 2] * X(j - 2) + l1[j - 1] * X(j - 1) + d0[j] * X(j);
      if (j < m2) R(j) += u1[j] * X(j + 1);
    }
    if (m2 > 1) {
      const int j = m2;
      R(j) = l2[j - 2] * X(j - 2) + l1[j - 1] * X(j - 1) + d0[j] * X(j);
    }
#undef X
#undef R
  }
}

/* src/hes_a2_shuffled_kernels.hpp:243-299; c2'(j) is left untouched (0) for j >= m2-1 (quirk Q5).
 * x may alias b. */
static void solve_a2(ho_ws *w, double *x, const double *b) {
  const int m1 = w->m1, m2 = w->m2, ld = m1 + 1, n = m2 + 1;
  const double *il2 = w->il2, *il1 = w->il1, *id0 = w->id0, *iu1 = w->iu1, *iu2 = w->iu2;
  double *c = w->cp, *c2 = w->c2p, *d = w->dp;
  for (int i = 0; i <= m1; ++i) {
    for (int j = 0; j < n; ++j) c2[j] = 0.0; /* zero-initialised view, never written for j >= n-2 */
#define B(j) b[(j) * ld + i]
#define X(j) x[(j) * ld + i]
    c[0] = iu1[0] / id0[0];
    c2[0] = iu2[0] / id0[0];
    d[0] = B(0) / id0[0];
    if (n > 1) {
      const double mm = 1.0 / (id0[1] - il1[0] * c[0]);
      c[1] = (iu1[1] - il1[0] * c2[0]) * mm;
      c2[1] = iu2[1] * mm;
      d[1] = (B(1) - il1[0] * d[0]) * mm;
    }
    for (int j = 2; j < n; ++j) {
      const double den = id0[j] - (il1[j - 1] - il2[j - 2] * c[j - 2]) * c[j - 1] - il2[j - 2] * c2[j - 2];
      const double m = 1.0 / den;
      /* The reference also evaluates c'(m2) from impl_upper(i, m2), one past that view's [m1+1][m2]
       * extent (src/hes_a2_shuffled_kernels.hpp:272); c'(m2) is never used by the back-substitution, so
       * it is simply not formed here. */
      if (j < n - 1) c[j] = (iu1[j] - (il1[j - 1] - il2[j - 2] * c[j - 2]) * c2[j - 1]) * m;
      if (j < n - 2) c2[j] = iu2[j] * m;
      d[j] = (B(j) - (il1[j - 1] - il2[j - 2] * c[j - 2]) * d[j - 1] - il2[j - 2] * d[j - 2]) * m;
    }
    X(n - 1) = d[n - 1];
    if (n > 1) X(n - 2) = d[n - 2] - c[n - 2] * X(n - 1);
    for (int j = n - 3; j >= 0; --j) X(j) = d[j] - c[j] * X(j + 1) - c2[j] * X(j + 2);
#undef B
#undef X
  }
}

/* ------------------------------------------------------------------ boundary vectors */

/* device path: src/hes_boundary_kernels.hpp:41-75 (b1 at index m1*(j+1): quirk Q3; b2 from i=0).
 * host path (Craig-Sneyd): src/BoundaryConditions.hpp:52-92 (b2 from i=1: quirk Q4). */
static void build_bounds(ho_ws *w, double r_d, double r_f, int N, double dt, int host_variant, int bc) {
  const int m1 = w->m1, m2 = w->m2, P = w->P;
  for (int p = 0; p < P; ++p) w->b[p] = w->b1[p] = w->b2[p] = 0.0;
  if (bc == 1) return;   /* put-correct set (extension): no inflow terms at s_max and v_max */
  const double ef = exp(-r_f * dt * (N - 1));
  for (int j = 0; j <= m2; ++j) w->b1[m1 * (j + 1)] = (r_d - r_f) * w->s[m1] * ef;
  for (int i = host_variant ? 1 : 0; i <= m1; ++i) w->b2[P - m1 - 1 + i] = -0.5 * r_d * w->s[i] * ef;
  for (int p = 0; p < P; ++p) w->b[p] = 0.0 + w->b1[p] + w->b2[p];
}

/* ------------------------------------------------------------------ Douglas stepping (device path) */

/* dividend jump: src/device_solver.hpp:448-504 */
static void dividend_jump(ho_ws *w, double *U, double amount, double pct) {
  const int m1 = w->m1, m2 = w->m2, ld = m1 + 1;
  memcpy(w->Ut, U, sizeof(double) * (size_t)w->P);
  for (int j = 0; j <= m2; ++j) {
    const int off = j * ld;
    for (int i = 0; i <= m1; ++i) {
      const double old_s = w->s[i];
      const double new_s = old_s * (1.0 - pct) - amount;
      if (new_s > 0) {
        int idx = 0;
        for (int k = 0; k <= m1; ++k)
          if (w->s[k] > new_s) {
            idx = k;
            break;
          }
        if (idx > 0 && idx < m1 + 1) {
          const double s_low = w->s[idx - 1], s_high = w->s[idx];
          const double weight = (new_s - s_low) / (s_high - s_low);
          U[off + i] = (1.0 - weight) * w->Ut[off + idx - 1] + weight * w->Ut[off + idx];
        } else if (idx == 0) {
          U[off + i] = w->Ut[off];
        } else {
          U[off + i] = w->Ut[off + m1];
        }
      } else {
        U[off + i] = 0.0;
      }
    }
  }
}

/* src/device_solver.hpp:194-266 (European), :276-374 (American), :384-641 (dividends), :652-942 (both).
 * U holds the payoff on entry and the solution on exit. */
static void douglas(ho_ws *w, const ho_numerics *num, int N, double dt, double r_f, double r_d, double K, double *U) {
  const int P = w->P, m1 = w->m1;
  const double theta = num->theta;
  const int am = num->style == 1;
  int div_idx = 0;
  if (am)
    for (int p = 0; p < P; ++p) w->lam[p] = 0;
  for (int n = 1; n <= N; ++n) {
    if (num->nd > 0 && num->div_all) {
      /* extension: the host solver's schedule (src/solver.hpp:363), every dividend dated inside the step */
      const double t = n * dt;
      while (div_idx < num->nd && t <= num->div_dates[div_idx] && num->div_dates[div_idx] < (n + 1) * dt) {
        dividend_jump(w, U, num->div_amounts[div_idx], num->div_pcts[div_idx]);
        div_idx++;
      }
    } else if (num->nd > 0) {
      /* one dividend per step at most, rank-0 index logic: src/device_solver.hpp:432-516 (quirk Q7) */
      const double t = n * dt;
      const int hit = (div_idx < num->nd && t <= num->div_dates[div_idx] && num->div_dates[div_idx] < (n + 1) * dt);
      if (hit) dividend_jump(w, U, num->div_amounts[div_idx], num->div_pcts[div_idx]);
      if (div_idx < num->nd && t > num->div_dates[div_idx]) div_idx++;
    }
    mul_a0(w, U, w->R0);
    mul_a1(w, U, w->R1);
    mul_a2(w, U, w->R2);
    {
      const double e0 = exp(r_f * dt * (n - 1));
      if (am)
        for (int p = 0; p < P; ++p)
          w->Y0[p] = U[p] + dt * (w->R0[p] + w->R1[p] + w->R2[p] + w->b[p] * e0 + w->lam[p]);
      else
        for (int p = 0; p < P; ++p) w->Y0[p] = U[p] + dt * (w->R0[p] + w->R1[p] + w->R2[p] + w->b[p] * e0);
    }
    {
      const double e1 = exp(r_f * dt * n), e0 = exp(r_f * dt * (n - 1));
      for (int p = 0; p < P; ++p) w->Y0[p] = w->Y0[p] + theta * dt * (w->b1[p] * e1 - (w->R1[p] + w->b1[p] * e0));
      solve_a1(w, w->Y1, w->Y0);
      for (int p = 0; p < P; ++p) w->Y1[p] = w->Y1[p] + theta * dt * (w->b2[p] * e1 - (w->R2[p] + w->b2[p] * e0));
      solve_a2(w, U, w->Y1);
    }
    if (num->bc == 1) {
      /* extension: Dirichlet value of a put at s_0 (row i = 0 of A1 is the identity row, so nothing else moves it) */
      const double g = K * exp(-r_d * dt * n);
      for (int j = 0; j <= w->m2; ++j) U[j * (m1 + 1)] = g;
    }
    if (am) {
      /* Ikonen-Toivanen projection: src/device_solver.hpp:358-372 */
      for (int p = 0; p < P; ++p) {
        const double U_bar = U[p];
        U[p] = fmax(U_bar - dt * w->lam[p], w->U0[p]);
        w->lam[p] = fmax(0.0, w->lam[p] + (w->U0[p] - U_bar) / dt);
        if (p % (m1 + 1) == m1) w->lam[p] = 0.0;
      }
    }
  }
}

/* ------------------------------------------------------------------ Craig-Sneyd (host path) */

/* src/solver.hpp:781-907, on the host classes: A1 multiply in (main, lower, upper) order
 * (src/hes_mat_fac.hpp:222-245), host boundary vectors, European only. */
static void craig_sneyd(ho_ws *w, const ho_numerics *num, int N, double dt, double r_f, double *U) {
  const int P = w->P;
  const double theta = num->theta;
  double *Y2 = dalloc(P), *A0Y2 = dalloc(P), *Y0t = dalloc(P), *Y1t = dalloc(P);
  for (int n = 1; n <= N; ++n) {
    const double e1 = exp(r_f * dt * n), e0 = exp(r_f * dt * (n - 1));
    mul_a0(w, U, w->R0);
    mul_a1_host(w, U, w->R1);
    mul_a2(w, U, w->R2);
    for (int p = 0; p < P; ++p) w->Y0[p] = U[p] + dt * (w->R0[p] + w->R1[p] + w->R2[p] + w->b[p] * e0);
    for (int p = 0; p < P; ++p) w->Y1[p] = w->Y0[p] + theta * dt * (w->b1[p] * e1 - (w->R1[p] + w->b1[p] * e0));
    solve_a1(w, w->Y1, w->Y1);
    for (int p = 0; p < P; ++p) Y2[p] = w->Y1[p] + theta * dt * (w->b2[p] * e1 - (w->R2[p] + w->b2[p] * e0));
    solve_a2(w, Y2, Y2);
    mul_a0(w, Y2, A0Y2);
    /* b0 == 0: (A0Y2 + b0 e1) - (A0U + b0 e0) */
    for (int p = 0; p < P; ++p) Y0t[p] = w->Y0[p] + 0.5 * dt * ((A0Y2[p] + 0.0 * e1) - (w->R0[p] + 0.0 * e0));
    for (int p = 0; p < P; ++p) Y1t[p] = Y0t[p] + theta * dt * (w->b1[p] * e1 - (w->R1[p] + w->b1[p] * e0));
    solve_a1(w, Y1t, Y1t);
    for (int p = 0; p < P; ++p) U[p] = Y1t[p] + theta * dt * (w->b2[p] * e1 - (w->R2[p] + w->b2[p] * e0));
    solve_a2(w, U, U);
  }
  free(Y2); free(A0Y2); free(Y0t); free(Y1t);
}


/* src/solver.hpp:917-1075, MCS_scheme_shuffled as shipped: the corrector starts from Y_0 AFTER it was overwritten with
 * the right-hand side of the first A1 solve (":968 Reuse Y_0 to store RHS"), and takes A1 U, A2 U in its two implicit
 * stages.  European only, host matrix classes and boundary vectors. */
static void modified_craig_sneyd(ho_ws *w, const ho_numerics *num, int N, double dt, double r_f, double *U) {
  const int P = w->P;
  const double theta = num->theta;
  double *Y2 = dalloc(P), *A0Y2 = dalloc(P), *A1Y2 = dalloc(P), *A2Y2 = dalloc(P), *Yt = dalloc(P), *prev = dalloc(P);
  for (int n = 1; n <= N; ++n) {
    const double e1 = exp(r_f * dt * n), e0 = exp(r_f * dt * (n - 1));
    mul_a0(w, U, w->R0);
    mul_a1_host(w, U, w->R1);
    mul_a2(w, U, w->R2);
    for (int p = 0; p < P; ++p) {
      prev[p] = w->R0[p] + w->R1[p] + w->R2[p] + w->b[p] * e0;
      w->Y0[p] = U[p] + dt * prev[p];
    }
    for (int p = 0; p < P; ++p) w->Y0[p] = w->Y0[p] + theta * dt * (w->b1[p] * e1 - (w->R1[p] + w->b1[p] * e0));
    solve_a1(w, w->Y1, w->Y0);
    for (int p = 0; p < P; ++p) w->Y1[p] = w->Y1[p] + theta * dt * (w->b2[p] * e1 - (w->R2[p] + w->b2[p] * e0));
    solve_a2(w, Y2, w->Y1);
    mul_a0(w, Y2, A0Y2);
    mul_a1_host(w, Y2, A1Y2);
    mul_a2(w, Y2, A2Y2);
    for (int p = 0; p < P; ++p) {
      const double F0_n = A0Y2[p] + 0.0 * e1, F0_nm1 = w->R0[p] + 0.0 * e0;   /* b0 == 0 */
      const double y0_hat = w->Y0[p] + theta * dt * (F0_n - F0_nm1);
      const double curr = A0Y2[p] + A1Y2[p] + A2Y2[p] + w->b[p] * e1;
      Yt[p] = y0_hat + (0.5 - theta) * dt * (curr - prev[p]);
    }
    for (int p = 0; p < P; ++p) Yt[p] = Yt[p] + theta * dt * (w->b1[p] * e1 - (w->R1[p] + w->b1[p] * e0));
    solve_a1(w, w->Y1, Yt);
    for (int p = 0; p < P; ++p) w->Y1[p] = w->Y1[p] + theta * dt * (w->b2[p] * e1 - (w->R2[p] + w->b2[p] * e0));
    solve_a2(w, U, w->Y1);
  }
  free(Y2); free(A0Y2); free(A1Y2); free(A2Y2); free(Yt); free(prev);
}

/* Hundsdorfer-Verwer (extension; in 't Hout & Foulon 2010, scheme (2.6)), on the same operators and boundary vectors:
 *   Y0 = U + dt F(U);  Yj = Y(j-1) + theta dt (Fj(Yj) - Fj(U)), j = 1, 2;
 *   Y0~ = Y0 + 1/2 dt (F(Y2) - F(U));  Yj~ = Y(j-1)~ + theta dt (Fj(Yj~) - Fj(Y2)), j = 1, 2;  U <- Y2~.
 * F(X) = A0 X + A1 X + A2 X + b e,  Fj(X) = Aj X + bj e; the explicit parts are summed left to right. */
static void hundsdorfer_verwer(ho_ws *w, const ho_numerics *num, int N, double dt, double r_f, double *U) {
  const int P = w->P;
  const double theta = num->theta;
  double *Y2 = dalloc(P), *A0Y2 = dalloc(P), *A1Y2 = dalloc(P), *A2Y2 = dalloc(P), *Yt = dalloc(P);
  for (int n = 1; n <= N; ++n) {
    const double e1 = exp(r_f * dt * n), e0 = exp(r_f * dt * (n - 1));
    mul_a0(w, U, w->R0);
    mul_a1_host(w, U, w->R1);
    mul_a2(w, U, w->R2);
    for (int p = 0; p < P; ++p) w->Y0[p] = U[p] + dt * (w->R0[p] + w->R1[p] + w->R2[p] + w->b[p] * e0);
    for (int p = 0; p < P; ++p) w->Y1[p] = w->Y0[p] + theta * dt * (w->b1[p] * e1 - (w->R1[p] + w->b1[p] * e0));
    solve_a1(w, w->Y1, w->Y1);
    for (int p = 0; p < P; ++p) Y2[p] = w->Y1[p] + theta * dt * (w->b2[p] * e1 - (w->R2[p] + w->b2[p] * e0));
    solve_a2(w, Y2, Y2);
    mul_a0(w, Y2, A0Y2);
    mul_a1_host(w, Y2, A1Y2);
    mul_a2(w, Y2, A2Y2);
    for (int p = 0; p < P; ++p) {
      const double curr = A0Y2[p] + A1Y2[p] + A2Y2[p] + w->b[p] * e1;
      const double prev = w->R0[p] + w->R1[p] + w->R2[p] + w->b[p] * e0;
      Yt[p] = w->Y0[p] + 0.5 * dt * (curr - prev);
    }
    /* the implicit stages of the corrector are centred on Y2: (I - theta dt A1) Y1~ = Y0~ - theta dt A1 Y2 (b1 terms cancel) */
    for (int p = 0; p < P; ++p) Yt[p] = Yt[p] + theta * dt * (w->b1[p] * e1 - (A1Y2[p] + w->b1[p] * e1));
    solve_a1(w, Yt, Yt);
    for (int p = 0; p < P; ++p) U[p] = Yt[p] + theta * dt * (w->b2[p] * e1 - (A2Y2[p] + w->b2[p] * e1));
    solve_a2(w, U, U);
  }
  free(Y2); free(A0Y2); free(A1Y2); free(A2Y2); free(Yt);
}

/* ------------------------------------------------------------------ one solve */

static void setup_option(ho_ws *w, const ho_model *mdl, const ho_numerics *num, double K, double V0_grid) {
  /* callers always use S = 8K, c = K/5, V = 5, d = V/500: src/heston_calibration.cpp:2614,
   * src/jacobian_computation.cpp:253 */
  ho_grid_s(w->m1, 8 * K, mdl->S0, K, K / 5, w->s, w->ds);
  ho_grid_v(w->m2, 5.0, V0_grid, 5.0 / 500, w->v, w->dv);
  for (int j = 0; j <= w->m2; ++j)
    for (int i = 0; i <= w->m1; ++i)
      w->U0[i + j * (w->m1 + 1)] = num->payoff_put ? fmax(K - w->s[i], 0.0) : fmax(w->s[i] - K, 0.0);
}

static void build_all(ho_ws *w, const ho_model *mdl, const ho_numerics *num, double dt) {
  build_a0(w, mdl->rho, mdl->sigma);
  build_a1(w, mdl->r_d, mdl->r_f, num->theta, dt);
  build_a2(w, mdl->r_d, mdl->kappa, mdl->eta, mdl->sigma, num->theta, dt);
}

static int solve_ws(ho_ws *w, const ho_model *mdl, const ho_numerics *num, int N, double dt, double K, double *U) {
  memcpy(U, w->U0, sizeof(double) * (size_t)w->P);
  if (num->scheme == 1)
    craig_sneyd(w, num, N, dt, mdl->r_f, U);
  else if (num->scheme == 2)
    modified_craig_sneyd(w, num, N, dt, mdl->r_f, U);
  else if (num->scheme == 3)
    hundsdorfer_verwer(w, num, N, dt, mdl->r_f, U);
  else
    douglas(w, num, N, dt, mdl->r_f, mdl->r_d, K, U);
  return 0;
}

/* price pick: src/jacobian_computation.cpp:433-445; find_v0_index returns 0 when no node matches
 * (src/grid_pod.hpp:76-87, quirk Q8). */
static int pick(const ho_ws *w, double S0, double V0, const double *U, double *price) {
  const int is = ho_find_index(w->s, w->m1 + 1, S0);
  int iv = ho_find_index(w->v, w->m2 + 1, V0);
  if (iv < 0) iv = 0;
  if (is < 0) return -1;
  *price = U[is + iv * (w->m1 + 1)];
  return 0;
}

int ho_solve(const ho_model *mdl, const ho_numerics *num, double K, int N, double dt, double V0_for_grid,
             double *price, double *U_out, double *lambda_out) {
  ho_ws *w = ws_new(num->m1, num->m2);
  double *U = dalloc(w->P);
  setup_option(w, mdl, num, K, V0_for_grid);
  build_bounds(w, mdl->r_d, mdl->r_f, N, dt, num->scheme >= 1, num->bc);
  build_all(w, mdl, num, dt);
  solve_ws(w, mdl, num, N, dt, K, U);
  const int rc = pick(w, mdl->S0, V0_for_grid, U, price);
  if (U_out) memcpy(U_out, U, sizeof(double) * (size_t)w->P);
  if (lambda_out) memcpy(lambda_out, w->lam, sizeof(double) * (size_t)w->P);
  free(U);
  ws_free(w);
  return rc;
}

/* ------------------------------------------------------------------ batches */

int ho_price_batch(const ho_model *mdl, const ho_numerics *num, int n, const double *strikes, const int *Ns,
                   const double *dts, double *prices) {
  int rc = 0;
  for (int k = 0; k < n; ++k)
    if (ho_solve(mdl, num, strikes[k], Ns[k], dts[k], mdl->V0, &prices[k], 0, 0) != 0) rc = -1;
  return rc;
}

/* src/jacobian_computation.cpp:204-364: base, then kappa/eta/sigma/rho + eps on the same grid, then a
 * solve on the grid rebuilt for V0 + eps read at the new v-index; J = (pert - base) / eps. */
int ho_jacobian_batch(const ho_model *mdl, const ho_numerics *num, int n, const double *strikes,
                      const int *Ns, const double *dts, double eps, double *J, double *base) {
  int rc = 0;
  ho_ws *w = ws_new(num->m1, num->m2);
  double *U = dalloc(w->P);
  for (int k = 0; k < n; ++k) {
    const int N = Ns[k];
    const double dt = dts[k];
    setup_option(w, mdl, num, strikes[k], mdl->V0);
    build_bounds(w, mdl->r_d, mdl->r_f, N, dt, num->scheme >= 1, num->bc);
    build_all(w, mdl, num, dt);
    solve_ws(w, mdl, num, N, dt, strikes[k], U);
    double base_price = 0.0;
    if (pick(w, mdl->S0, mdl->V0, U, &base_price) != 0) rc = -1;
    base[k] = base_price;
    for (int param = 0; param < 4; ++param) {
      ho_model m = *mdl;
      switch (param) {
        case 0: m.kappa += eps; break;
        case 1: m.eta += eps; break;
        case 2: m.sigma += eps; break;
        case 3: m.rho += eps; break;
      }
      build_all(w, &m, num, dt);
      solve_ws(w, &m, num, N, dt, strikes[k], U);
      double pert = 0.0;
      pick(w, mdl->S0, mdl->V0, U, &pert);
      J[k * 5 + param] = (pert - base_price) / eps;
    }
    {
      const double V0p = mdl->V0 + eps;
      ho_grid_v(w->m2, 5.0, V0p, 5.0 / 500, w->v, w->dv);
      build_all(w, mdl, num, dt);
      solve_ws(w, mdl, num, N, dt, strikes[k], U);
      double pert = 0.0;
      pick(w, mdl->S0, V0p, U, &pert);
      J[k * 5 + 4] = (pert - base_price) / eps;
    }
  }
  free(U);
  ws_free(w);
  return rc;
}

/* ------------------------------------------------------------------ LM */

/* src/jacobian_computation.cpp:20-104: partial-pivot Gaussian elimination with row normalisation. */
void ho_solve5(const double *Ain, const double *bin, double *x) {
  enum { N = 5 };
  double A[N * N], b[N];
  for (int i = 0; i < N; ++i) {
    b[i] = bin[i];
    for (int j = 0; j < N; ++j) A[i * N + j] = Ain[i * N + j];
  }
  for (int k = 0; k < N; ++k) {
    double maxA = fabs(A[k * N + k]);
    int piv = k;
    for (int p = k + 1; p < N; ++p) {
      const double val = fabs(A[p * N + k]);
      if (val > maxA) {
        maxA = val;
        piv = p;
      }
    }
    if (piv != k) {
      for (int c = 0; c < N; ++c) {
        const double t = A[k * N + c];
        A[k * N + c] = A[piv * N + c];
        A[piv * N + c] = t;
      }
      const double tb = b[k];
      b[k] = b[piv];
      b[piv] = tb;
    }
    const double pivot = A[k * N + k];
    for (int c = k + 1; c < N; ++c) A[k * N + c] /= pivot;
    b[k] /= pivot;
    A[k * N + k] = 1.0;
    for (int i = k + 1; i < N; ++i) {
      const double f = A[i * N + k];
      for (int c = k + 1; c < N; ++c) A[i * N + c] -= f * A[k * N + c];
      b[i] -= f * b[k];
      A[i * N + k] = 0.0;
    }
  }
  for (int k = N - 1; k >= 0; --k) {
    double val = b[k];
    for (int c = k + 1; c < N; ++c) val -= A[k * N + c] * b[c];
    b[k] = val;
  }
  for (int i = 0; i < N; ++i) x[i] = b[i];
}

/* src/jacobian_computation.cpp:107-195: JTJ (ascending-k sums), diag *= (1+lambda), JTr, 5x5 solve. */
void ho_lm_update(int n, const double *J, const double *r, double lambda, double *delta) {
  double A[25], g[5];
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 5; ++j) {
      double acc = 0.0;
      for (int k = 0; k < n; ++k) acc += J[k * 5 + i] * J[k * 5 + j];
      A[i * 5 + j] = 1.0 * acc;
    }
  for (int i = 0; i < 5; ++i) A[i * 5 + i] *= (1.0 + lambda);
  for (int i = 0; i < 5; ++i) {
    double acc = 0.0;
    for (int k = 0; k < n; ++k) acc += J[k * 5 + i] * r[k];
    g[i] = 1.0 * acc;
  }
  ho_solve5(A, g, delta);
}

/* src/heston_calibration.cpp:2692-2831 (multi-maturity twin of :204-417). */
int ho_calibrate(const ho_model *mdl0, const ho_numerics *num, int n, const double *strikes, const int *Ns,
                 const double *dts, const double *market, const ho_lm_opts *opt, ho_lm_result *res) {
  ho_model cur = *mdl0;
  double lambda = opt->lambda0;
  double *J = dalloc((size_t)n * 5), *base = dalloc(n), *r = dalloc(n), *newp = dalloc(n);
  int converged = 0, iters = 0, solves = 0;
  double final_error = 100.0, delta_norm = 0.0;
  for (int iter = 0; iter < opt->max_iter && !converged; ++iter) {
    ho_jacobian_batch(&cur, num, n, strikes, Ns, dts, opt->eps, J, base);
    solves += 6 * n;
    for (int i = 0; i < n; ++i) r[i] = market[i] - base[i];
    double delta[5];
    ho_lm_update(n, J, r, lambda, delta);
    ho_model nw = cur;
    nw.kappa = fmax(1e-3, cur.kappa + delta[0]);
    nw.eta = fmax(1e-2, cur.eta + delta[1]);
    nw.sigma = fmax(1e-2, cur.sigma + delta[2]);
    nw.rho = fmin(1.0, fmax(-1.0, cur.rho + delta[3]));
    nw.V0 = fmax(1e-2, cur.V0 + delta[4]);
    delta_norm = 0.0;
    for (int i = 0; i < 5; ++i) delta_norm += delta[i] * delta[i];
    delta_norm = sqrt(delta_norm);
    double cur_err = 0;
    for (int i = 0; i < n; ++i) cur_err += r[i] * r[i];
    if (delta_norm < opt->delta_tol || cur_err < opt->tol) {
      converged = 1;
      cur = nw;
      final_error = cur_err;
      iters = iter + 1;
      break;
    }
    ho_price_batch(&nw, num, n, strikes, Ns, dts, newp);
    solves += n;
    double new_err = 0.0;
    for (int i = 0; i < n; ++i) {
      const double rr = market[i] - newp[i];
      new_err += rr * rr;
    }
    if (new_err < cur_err) {
      cur = nw;
      lambda = fmax(lambda / 10.0, 1e-7);
    } else {
      lambda = fmin(lambda * 10.0, 1e7);
    }
    final_error = fmin(new_err, cur_err);
    iters = iter + 1;
  }
  res->params[0] = cur.kappa;
  res->params[1] = cur.eta;
  res->params[2] = cur.sigma;
  res->params[3] = cur.rho;
  res->params[4] = cur.V0;
  res->final_error = final_error;
  res->lambda = lambda;
  res->delta_norm = delta_norm;
  res->iterations = iters;
  res->converged = converged;
  res->pde_solves = solves;
  free(J); free(base); free(r); free(newp);
  return 0;
}

/* src/bs.hpp:44-55 */
double ho_bs_call(double S, double K, double r, double vol, double T) {
  const double sqrt_T = sqrt(T);
  const double log_SK = log(S / K);
  const double vol_sqrt_T = vol * sqrt_T;
  const double d1 = (log_SK + (r + 0.5 * vol * vol) * T) / vol_sqrt_T;
  const double d2 = d1 - vol_sqrt_T;
  return S * erfc(-d1 / sqrt(2.0)) / 2.0 - K * exp(-r * T) * erfc(-d2 / sqrt(2.0)) / 2.0;
}
