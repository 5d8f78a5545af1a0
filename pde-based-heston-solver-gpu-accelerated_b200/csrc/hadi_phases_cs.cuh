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

// 9-point A0 product at node (j, i) of array X (halo rows / zero frame coefficients as in phase E)
HADI_HD double hadi_cs_a0(const HadiView& w, const double* X, int i, int j) {
  const int ld = w.ld, n2 = w.n2;
  const double* p = X + j * ld + i;
  const double rs = hadi_ti(w, TI_RS)[i];
  const double bsm = hadi_ti(w, TI_BSM)[i], bs0 = hadi_ti(w, TI_BS0)[i], bsp = hadi_ti(w, TI_BSP)[i];
  const double* tj = w.tj;
  const double vj = tj[TJ_V * n2 + j];
  const double cij = rs * vj;
  const double csm = cij * bsm, cs0 = cij * bs0, csp = cij * bsp;
  const double bm = tj[TJ_BVM * n2 + j], b0 = tj[TJ_BV0 * n2 + j], bp = tj[TJ_BVP * n2 + j];
  double r0 = (csm * bm) * p[-ld - 1];
  r0 += (cs0 * bm) * p[-ld];
  r0 += (csp * bm) * p[-ld + 1];
  r0 += (csm * b0) * p[-1];
  r0 += (cs0 * b0) * p[0];
  r0 += (csp * b0) * p[1];
  r0 += (csm * bp) * p[ld - 1];
  r0 += (cs0 * bp) * p[ld];
  r0 += (csp * bp) * p[ld + 1];
  return r0;
}

// host boundary vectors at node (j, i): b1 at index m1*(j+1) (= node (j, m1-j), quirk Q3), b2 on the
// last v-row from i = 1 (src/BoundaryConditions.hpp:72,77)
HADI_HD void hadi_cs_bounds(const HadiItem& it, const HadiView& w, int i, int j, double& b1p, double& b2p) {
  const int m1 = w.m1, m2 = w.m2;
  b1p = (i + j == m1) ? (it.r_d - it.r_f) * hadi_ti(w, TI_S)[m1] * it.ef : 0.0;
  b2p = (j == m2 && i >= 1) ? hadi_ti(w, TI_B2V)[i] : 0.0;
}

// Predictor, explicit part: R0, R1, R2, Y0 and the right-hand side of the first A1 solve (into Y).
HADI_HD void hadi_cs_predict(const HadiItem& it, const HadiView& w, const HadiCsView& cs, double e0, double e1,
                             int tid, int nt, int scheme = HADI_SCHEME_CS) {
  const int m1 = w.m1, m2 = w.m2, ld = w.ld, n2 = w.n2;
  const HadiMap mp = hadi_map(m1, m2, tid, nt);
  if (!mp.active) return;
  const int i = mp.i;
  const double dt = it.dt, c = w.c;
  const double hs2 = hadi_ti(w, TI_HS2)[i];
  const double dsm = hadi_ti(w, TI_DSM)[i], ds0 = hadi_ti(w, TI_DS0)[i], dsp = hadi_ti(w, TI_DSP)[i];
  const double bbm = hadi_ti(w, TI_BBM)[i], bb0 = hadi_ti(w, TI_BB0)[i], bbp = hadi_ti(w, TI_BBP)[i];
  const double hrd = hadi_ti(w, TI_HRD)[i];
  const double* tj = w.tj;
  for (int j = mp.j0; j < mp.j1; ++j) {
    const double* p = w.U + j * ld + i;
    const double x = p[0];
    const double r0 = hadi_cs_a0(w, w.U, i, j);
    // A1, host order: main, lower, upper
    const double a = hs2 * tj[TJ_V * n2 + j];
    const double lo = a * dsm + bbm;
    const double ma = a * ds0 + bb0 - hrd;
    const double up = a * dsp + bbp;
    double r1 = ma * x;
    if (i > 0) r1 += lo * p[-1];
    if (i < m1) r1 += up * p[1];
    // A2 (padded diagonals, as phase E)
    double r2 = tj[TJ_L2 * n2 + j] * p[-2 * ld] + tj[TJ_L1 * n2 + j] * p[-ld] + tj[TJ_D0 * n2 + j] * x +
                tj[TJ_U1 * n2 + j] * p[ld];
    r2 += tj[TJ_U2 * n2 + j] * p[2 * ld];
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

// Right-hand side of an A2 solve from the stored R2: Y += theta dt (b2 e1 - (R2 + b2 e0)).
HADI_HD void hadi_cs_rhs2(const HadiItem& it, const HadiView& w, const HadiCsView& cs, double e0, double e1, int tid,
                          int nt) {
  const int m1 = w.m1, m2 = w.m2, ld = w.ld;
  const HadiMap mp = hadi_map(m1, m2, tid, nt);
  if (!mp.active) return;
  const int i = mp.i;
  const double c = w.c;
  for (int j = mp.j0; j < mp.j1; ++j) {
    double b1p, b2p;
    hadi_cs_bounds(it, w, i, j, b1p, b2p);
    const int q = j * ld + i;
    w.Y[q] = w.Y[q] + c * (b2p * e1 - (cs.R2[q] + b2p * e0));
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
  for (int j = mp.j0; j < mp.j1; ++j) {
    const double a0y2 = hadi_cs_a0(w, w.U, i, j);
    double b1p, b2p;
    hadi_cs_bounds(it, w, i, j, b1p, b2p);
    const int q = j * ld + i;
    const double y0t = cs.Y0[q] + 0.5 * dt * ((a0y2 + 0.0 * e1) - (cs.R0[q] + 0.0 * e0));
    w.Y[q] = y0t + c * (b1p * e1 - (cs.R1[q] + b1p * e0));
  }
}

// A1 (host order: main, lower, upper) and A2 products at node (j, i) of array X — the predictor's expressions
HADI_HD void hadi_cs_a1a2(const HadiItem& it, const HadiView& w, const double* X, int i, int j, double& r1, double& r2) {
  const int m1 = w.m1, ld = w.ld, n2 = w.n2;
  (void)it;
  const double* tj = w.tj;
  const double* p = X + j * ld + i;
  const double x = p[0];
  const double a = hadi_ti(w, TI_HS2)[i] * tj[TJ_V * n2 + j];
  const double lo = a * hadi_ti(w, TI_DSM)[i] + hadi_ti(w, TI_BBM)[i];
  const double ma = a * hadi_ti(w, TI_DS0)[i] + hadi_ti(w, TI_BB0)[i] - hadi_ti(w, TI_HRD)[i];
  const double up = a * hadi_ti(w, TI_DSP)[i] + hadi_ti(w, TI_BBP)[i];
  r1 = ma * x;
  if (i > 0) r1 += lo * p[-1];
  if (i < m1) r1 += up * p[1];
  r2 = tj[TJ_L2 * n2 + j] * p[-2 * ld] + tj[TJ_L1 * n2 + j] * p[-ld] + tj[TJ_D0 * n2 + j] * x + tj[TJ_U1 * n2 + j] * p[ld];
  r2 += tj[TJ_U2 * n2 + j] * p[2 * ld];
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
  for (int j = mp.j0; j < mp.j1; ++j) {
    const double a0y2 = hadi_cs_a0(w, w.U, i, j);
    double a1y2, a2y2;
    hadi_cs_a1a2(it, w, w.U, i, j, a1y2, a2y2);
    double b1p, b2p;
    hadi_cs_bounds(it, w, i, j, b1p, b2p);
    const double bb = 0.0 + b1p + b2p;
    const int q = j * ld + i;
    const double prev = cs.R0[q] + cs.R1[q] + cs.R2[q] + bb * e0;
    const double curr = a0y2 + a1y2 + a2y2 + bb * e1;
    if (scheme == HADI_SCHEME_MCS) {
      const double f0n = a0y2 + 0.0 * e1, f0m = cs.R0[q] + 0.0 * e0;
      const double y0h = cs.Y0[q] + c * (f0n - f0m);
      const double y0t = y0h + (0.5 - theta) * dt * (curr - prev);
      w.Y[q] = y0t + c * (b1p * e1 - (cs.R1[q] + b1p * e0));
    } else {
      const double y0t = cs.Y0[q] + 0.5 * dt * (curr - prev);
      w.Y[q] = y0t + c * (b1p * e1 - (a1y2 + b1p * e1));
      cs.R2[q] = a2y2;
    }
  }
}
