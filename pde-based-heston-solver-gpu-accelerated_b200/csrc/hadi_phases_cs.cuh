// hadi — Craig-Sneyd stages (global-state kernel only).
//
// The reference defines Craig-Sneyd only in its host-driven solver, CS_scheme_shuffled
// (src/solver.hpp:781-907), on the host matrix classes: A1 is multiplied in (main, lower, upper) order
// (src/hes_mat_fac.hpp:222-245) and the boundary vector b2 starts at i = 1
// (src/BoundaryConditions.hpp:77; quirk Q4).  Per step, with e0 = exp(r_f dt (n-1)), e1 = exp(r_f dt n):
//
//   R0 = A0 U, R1 = A1 U, R2 = A2 U
//   Y0 = U + dt (R0 + R1 + R2 + b e0)
//   Y1 = Y0 + theta dt (b1 e1 - (R1 + b1 e0));  (I - theta dt A1) Y1 = Y1
//   Y2 = Y1 + theta dt (b2 e1 - (R2 + b2 e0));  (I - theta dt A2) Y2 = Y2
//   Y0~ = Y0 + 0.5 dt ((A0 Y2 + b0 e1) - (R0 + b0 e0)),  b0 = 0
//   Y1~ = Y0~ + theta dt (b1 e1 - (R1 + b1 e0)); (I - theta dt A1) Y1~ = Y1~
//   U   = Y1~ + theta dt (b2 e1 - (R2 + b2 e0)); (I - theta dt A2) U = U
//
// R0, R1, R2 and Y0 are needed again by the corrector, so they are kept (per-CTA global scratch,
// L2 resident); Y2 overwrites U, which is dead once R0..R2 exist.  The line solves are the Douglas
// kernel's phases S1 / S2 (hadi_phases.cuh), unchanged.  Same phase discipline as hadi_phases.cuh:
// a thread only reads what other threads wrote in EARLIER phases.
//
// The same frame carries the two neighbouring splitting schemes, which share the
// predictor and differ in the corrector only:
//   scheme 2, Modified Craig-Sneyd AS THE REFERENCE SHIPS IT (src/solver.hpp:917-1075; its own comment says it does not
//     work, and it does not: the corrector starts from Y_0 after it was overwritten with the right-hand side of the first
//     A1 solve) — reproduced operation for operation and pinned against oracle/_ref:
//       Y0^ = Y0rhs + theta dt (A0 Y2 - A0 U);  Y0~ = Y0^ + (1/2 - theta) dt (F(Y2) - F(U));  then as Craig-Sneyd
//   scheme 3, Hundsdorfer-Verwer (not in the reference; extension, parity unpinned, defined by oracle/hadi_oracle.c):
//       Y0~ = Y0 + 1/2 dt (F(Y2) - F(U));  (I - theta dt Aj) Yj~ = Y(j-1)~ - theta dt Aj Y2
#pragma once
#include "hadi_phases.cuh"
#define HADI_SCHEME_CS 1
#define HADI_SCHEME_MCS 2
#define HADI_SCHEME_HV 3

struct HadiCsView {
  double *Y0, *R0, *R1, *R2;   // [m2+1][ld]
};

// ---- the explicit products from neighbour VALUES (the callers fetch them: several rows at once from global memory, so
// that more than one memory round trip is in flight per thread; the wide kernel with ld.global.cg).  Halo rows / zero
// frame coefficients as in phase E: no index clamping. -----------------------------------------------------------------
struct HadiNb {
  double mm, m0, mp, zm, z0, zp, pm, p0, pp, m2, p2;   // (j-1, i-1..i+1), (j, i-1..i+1), (j+1, i-1..i+1), (j-2, i), (j+2, i)
};
// 9-point A0 product, l outer, k inner: phase E, operation for operation
HADI_HD double hadi_nb_a0(const HadiView& w, const HadiNb& n, int i, int j) {
  const double rs = hadi_ti(w, TI_RS)[i];
  const double bsm = hadi_ti(w, TI_BSM)[i], bs0 = hadi_ti(w, TI_BS0)[i], bsp = hadi_ti(w, TI_BSP)[i];
  const double* tj = w.tj;
  const int n2 = w.n2;
  const double cij = rs * tj[TJ_V * n2 + j];
  const double csm = cij * bsm, cs0 = cij * bs0, csp = cij * bsp;
  const double bm = tj[TJ_BVM * n2 + j], b0 = tj[TJ_BV0 * n2 + j], bp = tj[TJ_BVP * n2 + j];
  double r0 = (csm * bm) * n.mm;
  r0 += (cs0 * bm) * n.m0;
  r0 += (csp * bm) * n.mp;
  r0 += (csm * b0) * n.zm;
  r0 += (cs0 * b0) * n.z0;
  r0 += (csp * b0) * n.zp;
  r0 += (csm * bp) * n.pm;
  r0 += (cs0 * bp) * n.p0;
  r0 += (csp * bp) * n.pp;
  return r0;
}
// A1 coefficients of node (j, i)
HADI_HD void hadi_nb_a1c(const HadiView& w, int i, int j, double& lo, double& ma, double& up) {
  const double a = hadi_ti(w, TI_HS2)[i] * w.tj[TJ_V * w.n2 + j];
  lo = a * hadi_ti(w, TI_DSM)[i] + hadi_ti(w, TI_BBM)[i];
  ma = a * hadi_ti(w, TI_DS0)[i] + hadi_ti(w, TI_BB0)[i] - hadi_ti(w, TI_HRD)[i];
  up = a * hadi_ti(w, TI_DSP)[i] + hadi_ti(w, TI_BBP)[i];
}
// A1 product in the host order of the Craig-Sneyd family: main, lower, upper
HADI_HD double hadi_nb_a1_host(const HadiView& w, const HadiNb& n, int i, int j) {
  double lo, ma, up;
  hadi_nb_a1c(w, i, j, lo, ma, up);
  double r1 = ma * n.z0;
  if (i > 0) r1 += lo * n.zm;
  if (i < w.m1) r1 += up * n.zp;
  return r1;
}
// A2 product (padded diagonals)
HADI_HD double hadi_nb_a2(const HadiView& w, const HadiNb& n, int j) {
  const double* tj = w.tj;
  const int n2 = w.n2;
  double r2 = tj[TJ_L2 * n2 + j] * n.m2 + tj[TJ_L1 * n2 + j] * n.m0 + tj[TJ_D0 * n2 + j] * n.z0 + tj[TJ_U1 * n2 + j] * n.p0;
  r2 += tj[TJ_U2 * n2 + j] * n.p2;
  return r2;
}
// Window of X over rows jb-2 .. jb+UNR+1 of column i and rows jb-1 .. jb+UNR of columns i-1, i+1: everything UNR
// consecutive nodes of a column need, requested before the first is used.  Rows past the lower halo are clamped (their
// values are never used: the nodes they would feed do not exist).
#define HADI_CS_UNR 4
template <int UNR>
struct HadiWin {
  double c0[UNR + 4], cm[UNR + 2], cp[UNR + 2];
};
template <int UNR>
HADI_HD void hadi_win_load(const double* X, int ld, int i, int jb, int m2, HadiWin<UNR>& W) {
#pragma unroll
  for (int r = 0; r < UNR + 4; ++r) {
    const int jr = (jb - 2 + r < m2 + 2) ? jb - 2 + r : m2 + 2;
    W.c0[r] = X[jr * ld + i];
  }
#pragma unroll
  for (int r = 0; r < UNR + 2; ++r) {
    const int jr = (jb - 1 + r < m2 + 1) ? jb - 1 + r : m2 + 1;
    W.cm[r] = X[jr * ld + i - 1];
    W.cp[r] = X[jr * ld + i + 1];
  }
}
template <int UNR>
HADI_HD HadiNb hadi_win_nb(const HadiWin<UNR>& W, int k) {
  HadiNb n;
  n.mm = W.cm[k]; n.m0 = W.c0[k + 1]; n.mp = W.cp[k];
  n.zm = W.cm[k + 1]; n.z0 = W.c0[k + 2]; n.zp = W.cp[k + 1];
  n.pm = W.cm[k + 2]; n.p0 = W.c0[k + 3]; n.pp = W.cp[k + 2];
  n.m2 = W.c0[k]; n.p2 = W.c0[k + 4];
  return n;
}

// Douglas explicit stage at node (j, i) from neighbour values and the current multiplier (hadi_phase_explicit, operation
// for operation: A1 in the device order lower, main, upper)
HADI_HD double hadi_nb_explicit(const HadiItem& it, const HadiView& w, const HadiNb& n, double lam_cur, double e0, double e1,
                                int i, int j) {
  const int m1 = w.m1, m2 = w.m2;
  const double dt = it.dt, c = w.c;
  const bool am = it.style == 1;
  const double x = n.z0;
  const double r0 = hadi_nb_a0(w, n, i, j);
  double lo, ma, upc;
  hadi_nb_a1c(w, i, j, lo, ma, upc);
  const double r1 = lo * n.zm + ma * x + upc * n.zp;
  const double r2 = hadi_nb_a2(w, n, j);
  const bool is_b1 = (i + j == m1);
  double y;
  if (is_b1 || j == m2) {
    const double b1v = it.bc ? 0.0 : (it.r_d - it.r_f) * hadi_ti(w, TI_S)[m1] * it.ef;
    const double b1p = is_b1 ? b1v : 0.0;
    const double b2p = (j == m2) ? hadi_ti(w, TI_B2V)[i] : 0.0;
    const double bp_ = 0.0 + b1p + b2p;
    double sum = r0 + r1 + r2 + bp_ * e0;
    if (am) sum = sum + lam_cur;
    y = x + dt * sum;
    y = y + c * (b1p * e1 - (r1 + b1p * e0));
  } else {
    double sum = r0 + r1 + r2;
    if (am) sum = sum + lam_cur;
    y = x + dt * sum;
    y = y - c * r1;
  }
  return y;
}
// host boundary vectors at node (j, i): b1 at index m1*(j+1) (= node (j, m1-j), quirk Q3), b2 on the
// last v-row from i = 1 (src/BoundaryConditions.hpp:72,77)
HADI_HD void hadi_cs_bounds(const HadiItem& it, const HadiView& w, int i, int j, double& b1p, double& b2p) {
  const int m1 = w.m1, m2 = w.m2;
  b1p = (i + j == m1) ? (it.r_d - it.r_f) * hadi_ti(w, TI_S)[m1] * it.ef : 0.0;
  b2p = (j == m2 && i >= 1) ? hadi_ti(w, TI_B2V)[i] : 0.0;
}

// Predictor, explicit part: R0, R1, R2, Y0 and the right-hand side of the first A1 solve (into Y).
// The point-wise stages of the global-state kernels stream P-sized arrays through L2 / HBM with one thread per column
// walking down the rows; each thread fetches what HADI_CS_UNR rows need before it computes the first, so that four
// memory round trips overlap where one was in flight (ncu, 148 solves of 401 x 201: these stages were 45 % of the
// kernel, all of it long-scoreboard stall).
HADI_HD void hadi_cs_predict(const HadiItem& it, const HadiView& w, const HadiCsView& cs, double e0, double e1,
                             int tid, int nt, int scheme = HADI_SCHEME_CS) {
  const int m1 = w.m1, m2 = w.m2, ld = w.ld;
  const HadiMap mp = hadi_map(m1, m2, tid, nt);
  if (!mp.active) return;
  const int i = mp.i;
  const double dt = it.dt, c = w.c;
  constexpr int UNR = HADI_CS_UNR;
  for (int jb = mp.j0; jb < mp.j1; jb += UNR) {
    HadiWin<UNR> W;
    hadi_win_load<UNR>(w.U, ld, i, jb, m2, W);
#pragma unroll
    for (int k = 0; k < UNR; ++k) {
      const int j = jb + k;
      if (j < mp.j1) {
        const HadiNb n = hadi_win_nb<UNR>(W, k);
        const double x = n.z0;
        const double r0 = hadi_nb_a0(w, n, i, j);
        const double r1 = hadi_nb_a1_host(w, n, i, j);   // A1, host order: main, lower, upper
        const double r2 = hadi_nb_a2(w, n, j);           // A2 (padded diagonals, as phase E)
        double b1p, b2p;
        hadi_cs_bounds(it, w, i, j, b1p, b2p);
        const double bb = 0.0 + b1p + b2p;
        const double y0 = x + dt * (r0 + r1 + r2 + bb * e0);
        const int q = j * ld + i;
        cs.R0[q] = r0;
        cs.R1[q] = r1;
        cs.R2[q] = r2;
        const double rhs = y0 + c * (b1p * e1 - (r1 + b1p * e0));
        cs.Y0[q] = (scheme == HADI_SCHEME_MCS) ? rhs : y0;   // the shipped MCS keeps the overwritten Y_0 (src/solver.hpp:968)
        w.Y[q] = rhs;
      }
    }
  }
}

// Right-hand side of an A2 solve from the stored R2: Y += theta dt (b2 e1 - (R2 + b2 e0)).
HADI_HD void hadi_cs_rhs2(const HadiItem& it, const HadiView& w, const HadiCsView& cs, double e0, double e1, int tid,
                          int nt) {
  const int m1 = w.m1, m2 = w.m2, ld = w.ld;
  const HadiMap mp = hadi_map(m1, m2, tid, nt);
  if (!mp.active) return;
  const int i = mp.i;
  const double c = w.c;
  constexpr int UNR = 2 * HADI_CS_UNR;
  for (int jb = mp.j0; jb < mp.j1; jb += UNR) {
    double yv[UNR], rv[UNR];
#pragma unroll
    for (int k = 0; k < UNR; ++k) {
      const int j = (jb + k < mp.j1) ? jb + k : mp.j1 - 1;
      yv[k] = w.Y[j * ld + i];
      rv[k] = cs.R2[j * ld + i];
    }
#pragma unroll
    for (int k = 0; k < UNR; ++k) {
      const int j = jb + k;
      if (j < mp.j1) {
        double b1p, b2p;
        hadi_cs_bounds(it, w, i, j, b1p, b2p);
        w.Y[j * ld + i] = yv[k] + c * (b2p * e1 - (rv[k] + b2p * e0));
      }
    }
  }
}

// Corrector, explicit part: Y2 sits in U.  Y0~ = Y0 + 0.5 dt ((A0 Y2 + b0 e1) - (R0 + b0 e0)) and the
// right-hand side of the second A1 solve (into Y).
HADI_HD void hadi_cs_correct(const HadiItem& it, const HadiView& w, const HadiCsView& cs, double e0, double e1,
                             int tid, int nt) {
  const int m1 = w.m1, m2 = w.m2, ld = w.ld;
  const HadiMap mp = hadi_map(m1, m2, tid, nt);
  if (!mp.active) return;
  const int i = mp.i;
  const double dt = it.dt, c = w.c;
  constexpr int UNR = HADI_CS_UNR;
  for (int jb = mp.j0; jb < mp.j1; jb += UNR) {
    HadiWin<UNR> W;
    hadi_win_load<UNR>(w.U, ld, i, jb, m2, W);
    double y0v[UNR], r0v[UNR], r1v[UNR];
#pragma unroll
    for (int k = 0; k < UNR; ++k) {
      const int q = ((jb + k < mp.j1) ? jb + k : mp.j1 - 1) * ld + i;
      y0v[k] = cs.Y0[q];
      r0v[k] = cs.R0[q];
      r1v[k] = cs.R1[q];
    }
#pragma unroll
    for (int k = 0; k < UNR; ++k) {
      const int j = jb + k;
      if (j < mp.j1) {
        const HadiNb n = hadi_win_nb<UNR>(W, k);
        const double a0y2 = hadi_nb_a0(w, n, i, j);
        double b1p, b2p;
        hadi_cs_bounds(it, w, i, j, b1p, b2p);
        const double y0t = y0v[k] + 0.5 * dt * ((a0y2 + 0.0 * e1) - (r0v[k] + 0.0 * e0));
        w.Y[j * ld + i] = y0t + c * (b1p * e1 - (r1v[k] + b1p * e0));
      }
    }
  }
}

// Correctors of the Modified Craig-Sneyd (as shipped) and Hundsdorfer-Verwer schemes; Y2 sits in U.  Both need
// F(Y2) = A0 Y2 + A1 Y2 + A2 Y2 + b e1 and F(U) = R0 + R1 + R2 + b e0 (summed left to right, as the reference does).
// Hundsdorfer-Verwer centres its two implicit stages on Y2: A1 Y2 enters the right-hand side formed here, A2 Y2
// replaces R2 (dead after this stage) for hadi_cs_rhs2, which is then called with e0 = e1.
HADI_HD void hadi_cs_correct2(const HadiItem& it, const HadiView& w, const HadiCsView& cs, double e0, double e1,
                              int tid, int nt, int scheme) {
  const int m1 = w.m1, m2 = w.m2, ld = w.ld;
  const HadiMap mp = hadi_map(m1, m2, tid, nt);
  if (!mp.active) return;
  const int i = mp.i;
  const double dt = it.dt, c = w.c, theta = it.theta;
  constexpr int UNR = HADI_CS_UNR;
  for (int jb = mp.j0; jb < mp.j1; jb += UNR) {
    HadiWin<UNR> W;
    hadi_win_load<UNR>(w.U, ld, i, jb, m2, W);
    double y0v[UNR], r0v[UNR], r1v[UNR], r2v[UNR];
#pragma unroll
    for (int k = 0; k < UNR; ++k) {
      const int q = ((jb + k < mp.j1) ? jb + k : mp.j1 - 1) * ld + i;
      y0v[k] = cs.Y0[q];
      r0v[k] = cs.R0[q];
      r1v[k] = cs.R1[q];
      r2v[k] = cs.R2[q];
    }
#pragma unroll
    for (int k = 0; k < UNR; ++k) {
      const int j = jb + k;
      if (j < mp.j1) {
        const HadiNb n = hadi_win_nb<UNR>(W, k);
        const double a0y2 = hadi_nb_a0(w, n, i, j);
        const double a1y2 = hadi_nb_a1_host(w, n, i, j);
        const double a2y2 = hadi_nb_a2(w, n, j);
        double b1p, b2p;
        hadi_cs_bounds(it, w, i, j, b1p, b2p);
        const double bb = 0.0 + b1p + b2p;
        const int q = j * ld + i;
        const double prev = r0v[k] + r1v[k] + r2v[k] + bb * e0;
        const double curr = a0y2 + a1y2 + a2y2 + bb * e1;
        if (scheme == HADI_SCHEME_MCS) {
          const double f0n = a0y2 + 0.0 * e1, f0m = r0v[k] + 0.0 * e0;
          const double y0h = y0v[k] + c * (f0n - f0m);
          const double y0t = y0h + (0.5 - theta) * dt * (curr - prev);
          w.Y[q] = y0t + c * (b1p * e1 - (r1v[k] + b1p * e0));
        } else {
          const double y0t = y0v[k] + 0.5 * dt * (curr - prev);
          w.Y[q] = y0t + c * (b1p * e1 - (a1y2 + b1p * e1));
          cs.R2[q] = a2y2;
        }
      }
    }
  }
}
