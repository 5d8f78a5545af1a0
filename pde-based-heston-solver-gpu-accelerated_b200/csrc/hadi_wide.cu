// hadi — the WIDE kernel: one solve spread over a TEAM of G co-resident CTAs (G up to one CTA per SM of the GPU).
//
// What bounds ONE large solve (401 x 201, 200 steps, Craig-Sneyd) is not memory or arithmetic throughput but the
// dependent FP64 chain of its line solves: a forward / backward Thomas sweep along 400 nodes is 400 x (16 + 40)
// cycles whoever runs it, and bit parity with the reference forbids re-associating it.  The lines of one sweep are
// independent, so the fastest schedule gives every line its own thread (the lines of a CTA packed into at most two warps
// per scheduler, one lane each), with every operand of the chain already in shared memory, and lets all other work of
// the step disappear behind it:
//
//   * rows (A1 sweeps) are dealt round-robin to the CTAs of the team, columns (A2 sweeps) in contiguous blocks;
//   * the point-wise stage that FEEDS a sweep is evaluated by the CTA that owns the line, straight into the
//     shared-memory line buffer (predictor / corrector / explicit stage before an A1 sweep, the A2 right-hand side
//     before an A2 sweep, the American projection after it) — no separate point-wise phase, no extra barrier;
//   * the A1 factors of a CTA's rows stay in shared memory for the whole solve when they fit;
//   * a step therefore has as many team barriers as it has row <-> column transpositions: 2 (Douglas) or 4
//     (Craig-Sneyd family), each a release / acquire counter in global memory (the grid is launched co-operatively,
//     every CTA is resident);
//   * the state arrays (U with halo, Y, and Y0 / R0 / R1 / R2 for the Craig-Sneyd family) live in the L2-resident
//     scratch block of the team and are read with ld.global.cg (other SMs write them between barriers).
//
// Arithmetic: the expressions of hadi_phases.cuh / hadi_phases_cs.cuh, operation for operation (the tables and the
// factorisation ARE those functions); results are bit-identical to every other variant.  A guarded division that leaves
// its operand range is redone with the IEEE division, one chunk of the chain at a time, so there is no re-solve pass.
//
// Replaces, for batches of a few solves, the thread-block-cluster kernel of round 1 (hadi_cluster_kernel, 8 CTAs, every
// phase through L2): 401 x 201 x 200 Craig-Sneyd 67.7 -> 12.3 ms for one solve.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdlib>

#include "hadi_launch.h"
#include "hadi_phases_cs.cuh"

namespace {

constexpr int kWideThreads = 512;
constexpr int kWideWarps = kWideThreads / 32;
#ifndef HADI_WIDE_KCH
#define HADI_WIDE_KCH 4   /* measured on B200: 6 and 8 nodes per chunk are no faster (tools/dev_timing_wide.py) */
#endif
constexpr int kCh = HADI_WIDE_KCH;   // nodes per register chunk of a chain
constexpr int kPad = 2 * kCh;   // the sweeps are unrolled by two chunks and prefetch one chunk ahead

__device__ __forceinline__ double wld(const double* p) { return __ldcg(p); }

// ---- per-stage cycle counters (HADI_PHASE_TIMING builds; thread 0 of every CTA, read with hadi_batch_prof_raw) ----------
// [0] set-up  [1] team barriers  [2] right-hand sides of the A1 sweeps  [3] A1 chains  [4] A2 chains
// [5] right-hand sides of the A2 sweeps  [6] solutions back to global memory (+ projection)  [7] factor reloads
struct WideProf {
#ifdef HADI_PHASE_TIMING
  long long* p;
  long long t;
#endif
};
__device__ __forceinline__ void wide_tick(WideProf& pf, int k, int tid) {
#ifdef HADI_PHASE_TIMING
  if (tid == 0 && pf.p != nullptr) {
    const long long now = clock64();
    pf.p[k] += now - pf.t;
    pf.t = now;
  }
#else
  (void)pf; (void)k; (void)tid;
#endif
}

// ---- team barrier ------------------------------------------------------------------------------------------------
struct WideTeam {
  unsigned* ctr;    // monotonic arrival counter of the team (zeroed by the host before the launch)
  unsigned epoch;   // arrivals after which the next barrier opens (thread 0 only)
  int G;
};
__device__ __forceinline__ void wide_sync(WideTeam& tm, int tid) {
  __syncthreads();
  if (tm.G > 1 && tid == 0) {
    // release: the CTA's writes (ordered before this point by the barrier above) are visible at GPU scope before the
    // arrival is; acquire: what the other CTAs wrote before arriving is visible to this CTA after the barrier below
    tm.epoch += (unsigned)tm.G;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(tm.ctr) : "memory");
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(tm.ctr) : "memory");
    } while ((int)(v - tm.epoch) < 0);
  }
  __syncthreads();
}

// ---- ownership and shared-memory line buffers -----------------------------------------------------------------------
struct WideGeo {
  int rank, G;
  int nrow;          // owned rows: rank + G*k, k < nrow
  int c0, ncol;      // owned columns [c0, c0 + ncol)
  int RB, CB;        // lines per batch
  int pr, pc;        // entries per row buffer (pairs) / per column buffer (doubles), padding included
  bool resident;     // every owned row fits one batch: its A1 factors are loaded once per solve
  // Row buffers, [RB][pr] pairs each, node i / element k at entry kPad + i: the chains run whole chunks and may touch
  // kPad entries either side of a line.
  double2 *rA;       // {right-hand side -> solution, Thomas multiplier m_i}
  double2 *rB;       // {forward result d_i, impl_upper_i}
  double2 *rC;       // element k of the back substitution (node m1 - k): {pivot, prepared reciprocal}
  double *cb, *cd, *cl;   // [CB][pc] right-hand side -> solution | forward result | lambda; row j at entry kPad + j
  double* t2;        // [n2 + 2 kPad][6] A2 factor records {F, G, MM, -, CP, C2P}, row j at record kPad + j
};

// the 11 neighbours of node (j, i) the explicit operators need (products: hadi_nb_* in hadi_phases_cs.cuh)
using WideNb = HadiNb;
__device__ __forceinline__ WideNb wide_nb(const double* X, int ld, int i, int j) {
  const double* p = X + j * ld + i;
  WideNb n;
  n.mm = wld(p - ld - 1); n.m0 = wld(p - ld); n.mp = wld(p - ld + 1);
  n.zm = wld(p - 1); n.z0 = wld(p); n.zp = wld(p + 1);
  n.pm = wld(p + ld - 1); n.p0 = wld(p + ld); n.pp = wld(p + ld + 1);
  n.m2 = wld(p - 2 * ld); n.p2 = wld(p + 2 * ld);
  return n;
}

// ---- right-hand sides of the A1 sweeps, one node ----------------------------------------------------------------------
// Douglas explicit stage (hadi_phase_explicit)
__device__ __forceinline__ double wide_node_explicit(const HadiItem& it, const HadiView& w, double e0, double e1, int i,
                                                     int j) {
  const WideNb n = wide_nb(w.U, w.ld, i, j);
  const double lam_cur = (it.style == 1) ? wld(w.lam + j * w.ld + i) : 0.0;
  return hadi_nb_explicit(it, w, n, lam_cur, e0, e1, i, j);
}
// Craig-Sneyd family, predictor (hadi_cs_predict): keeps R0, R1, R2, Y0 of the node
__device__ __forceinline__ double wide_node_predict(const HadiItem& it, const HadiView& w, const HadiCsView& cs, double e0,
                                                    double e1, int i, int j, int scheme) {
  const double dt = it.dt, c = w.c;
  const WideNb n = wide_nb(w.U, w.ld, i, j);
  const double x = n.z0;
  const double r0 = hadi_nb_a0(w, n, i, j);
  const double r1 = hadi_nb_a1_host(w, n, i, j);
  const double r2 = hadi_nb_a2(w, n, j);
  double b1p, b2p;
  hadi_cs_bounds(it, w, i, j, b1p, b2p);
  const double bb = 0.0 + b1p + b2p;
  const double y0 = x + dt * (r0 + r1 + r2 + bb * e0);
  const int q = j * w.ld + i;
  cs.R0[q] = r0;
  cs.R1[q] = r1;
  cs.R2[q] = r2;
  const double rhs = y0 + c * (b1p * e1 - (r1 + b1p * e0));
  cs.Y0[q] = (scheme == HADI_SCHEME_MCS) ? rhs : y0;
  return rhs;
}
// correctors (hadi_cs_correct, hadi_cs_correct2); Y2 sits in U
__device__ __forceinline__ double wide_node_correct(const HadiItem& it, const HadiView& w, const HadiCsView& cs, double e0,
                                                    double e1, int i, int j, int scheme) {
  const double dt = it.dt, c = w.c, theta = it.theta;
  const WideNb n = wide_nb(w.U, w.ld, i, j);
  const double a0y2 = hadi_nb_a0(w, n, i, j);
  double b1p, b2p;
  hadi_cs_bounds(it, w, i, j, b1p, b2p);
  const int q = j * w.ld + i;
  if (scheme == HADI_SCHEME_CS) {
    const double y0t = wld(cs.Y0 + q) + 0.5 * dt * ((a0y2 + 0.0 * e1) - (wld(cs.R0 + q) + 0.0 * e0));
    return y0t + c * (b1p * e1 - (wld(cs.R1 + q) + b1p * e0));
  }
  const double a1y2 = hadi_nb_a1_host(w, n, i, j);
  const double a2y2 = hadi_nb_a2(w, n, j);
  const double bb = 0.0 + b1p + b2p;
  const double R0 = wld(cs.R0 + q), R1 = wld(cs.R1 + q), R2 = wld(cs.R2 + q);
  const double prev = R0 + R1 + R2 + bb * e0;
  const double curr = a0y2 + a1y2 + a2y2 + bb * e1;
  if (scheme == HADI_SCHEME_MCS) {
    const double f0n = a0y2 + 0.0 * e1, f0m = R0 + 0.0 * e0;
    const double y0h = wld(cs.Y0 + q) + c * (f0n - f0m);
    const double y0t = y0h + (0.5 - theta) * dt * (curr - prev);
    return y0t + c * (b1p * e1 - (R1 + b1p * e0));
  }
  const double y0t = wld(cs.Y0 + q) + 0.5 * dt * (curr - prev);
  cs.R2[q] = a2y2;   // Hundsdorfer-Verwer: the second A2 right-hand side is centred on Y2
  return y0t + c * (b1p * e1 - (a1y2 + b1p * e1));
}

// ---- the chains: one thread per line, every operand in shared memory, next chunk's operands in flight --------------------
// The chain of a chunk is straight-line code: whole chunks only (nodes past the end of a line compute on the padding of
// the line buffers), operands arrive as 16-byte pairs, two chunk bodies per loop trip swap their register sets instead
// of copying them, and the guarded division raises a flag instead of branching; a chunk whose flag is up (operands
// below 2^-900 in the far out-of-the-money corner of a large grid) is redone with the IEEE division from the saved
// entry value — the same bits as dividing in line, without a branch per node on the chain.
// A1 (hadi_phase_solve_a1): forward x_i = y_i - m_i x_{i-1}; back x_i = (x_i - impl_upper_i x_{i+1}) / pivot_i
struct WideA1Fwd { double2 v[kCh]; };
struct WideA1Bwd { double2 c[kCh], b[kCh]; };
__device__ __forceinline__ void wide_a1_fwd_chunk(const WideA1Fwd& cur, WideA1Fwd& nxt, const double2* __restrict__ A,
                                                  double2* __restrict__ B, int ib, double& x) {
#pragma unroll
  for (int k = 0; k < kCh; ++k) nxt.v[k] = A[ib + kCh + k];
#pragma unroll
  for (int k = 0; k < kCh; ++k) {
    x = cur.v[k].x - cur.v[k].y * x;
    B[ib + k].x = x;
  }
}
__device__ __forceinline__ void wide_a1_bwd_chunk(const WideA1Bwd& cur, WideA1Bwd& nxt, double2* __restrict__ A,
                                                  const double2* __restrict__ B, const double2* __restrict__ C, int m1, int kb,
                                                  double& xn) {
#pragma unroll
  for (int k = 0; k < kCh; ++k) {
    nxt.c[k] = C[kb + kCh + k];
    nxt.b[k] = B[m1 - (kb + kCh + k)];
  }
  const double x_in = xn;
  double xo[kCh];
  unsigned bad = 0;
#pragma unroll
  for (int k = 0; k < kCh; ++k) {
    xn = hadi_div<false, false>(cur.b[k].x - cur.b[k].y * xn, cur.c[k].x, cur.c[k].y, bad);
    xo[k] = xn;
  }
  if (bad != 0u) {
    xn = x_in;
#pragma unroll
    for (int k = 0; k < kCh; ++k) {
      xn = (cur.b[k].x - cur.b[k].y * xn) / cur.c[k].x;
      xo[k] = xn;
    }
  }
#pragma unroll
  for (int k = 0; k < kCh; ++k) A[m1 - kb - k].x = xo[k];
}
// A, B, C point at node 0 / element 0 of the line (kPad entries of padding either side)
__device__ __forceinline__ void wide_chain_a1(double2* __restrict__ A, double2* __restrict__ B, const double2* __restrict__ C,
                                              int m1) {
  double x = A[0].x;
  B[0].x = x;
  {
    WideA1Fwd p, q;
#pragma unroll
    for (int k = 0; k < kCh; ++k) p.v[k] = A[1 + k];
#pragma unroll 1
    for (int ib = 1; ib <= m1; ib += 2 * kCh) {
      wide_a1_fwd_chunk(p, q, A, B, ib, x);
      wide_a1_fwd_chunk(q, p, A, B, ib + kCh, x);
    }
  }
  double xn = 0.0;
  {
    // element k of the back-substitution order is node i = m1 - k
    WideA1Bwd p, q;
#pragma unroll
    for (int k = 0; k < kCh; ++k) { p.c[k] = C[k]; p.b[k] = B[m1 - k]; }
#pragma unroll 1
    for (int kb = 0; kb < m1; kb += 2 * kCh) {
      wide_a1_bwd_chunk(p, q, A, B, C, m1, kb, xn);
      wide_a1_bwd_chunk(q, p, A, B, C, m1, kb + kCh, xn);
    }
  }
  A[0].x = B[0].x;
}
// A2 (hadi_phase_solve_a2): d_j = (b_j - f_j d_{j-1} - g_j d_{j-2}) m_j; x_j = d_j - c'_j x_{j+1} - c2'_j x_{j+2}
// T: factor records {F, G, MM, -, CP, C2P} of row 0 (kPad records of padding either side); b, d: row 0 of the column
struct WideA2Fwd { double b[kCh], mm[kCh]; double2 fg[kCh]; };
struct WideA2Bwd { double d[kCh]; double2 cc[kCh]; };
__device__ __forceinline__ void wide_a2_fwd_chunk(const WideA2Fwd& cur, WideA2Fwd& nxt, const double* __restrict__ b,
                                                  double* __restrict__ d, const double* __restrict__ T, int jb, double& d1,
                                                  double& d2) {
#pragma unroll
  for (int k = 0; k < kCh; ++k) {
    const int j = jb + kCh + k;
    nxt.b[k] = b[j];
    nxt.fg[k] = *reinterpret_cast<const double2*>(T + 6 * j);
    nxt.mm[k] = T[6 * j + 2];
  }
#pragma unroll
  for (int k = 0; k < kCh; ++k) {
    const double v = (cur.b[k] - cur.fg[k].x * d1 - cur.fg[k].y * d2) * cur.mm[k];
    d[jb + k] = v;
    d2 = d1;
    d1 = v;
  }
}
__device__ __forceinline__ void wide_a2_bwd_chunk(const WideA2Bwd& cur, WideA2Bwd& nxt, double* __restrict__ b,
                                                  const double* __restrict__ d, const double* __restrict__ T, int jt, double& x1,
                                                  double& x2) {
#pragma unroll
  for (int k = 0; k < kCh; ++k) {
    const int j = jt - kCh - k;
    nxt.d[k] = d[j];
    nxt.cc[k] = *reinterpret_cast<const double2*>(T + 6 * j + 4);
  }
#pragma unroll
  for (int k = 0; k < kCh; ++k) {
    const double x = cur.d[k] - cur.cc[k].x * x1 - cur.cc[k].y * x2;
    x2 = x1;
    x1 = x;
    b[jt - k] = x;
  }
}
__device__ __forceinline__ void wide_chain_a2(double* __restrict__ b, double* __restrict__ d, const double* __restrict__ T,
                                              const double* __restrict__ tj, int n2, int m2) {
  unsigned bad = 0;
  double d1 = hadi_div<false, true>(b[0], tj[TJ_MM * n2], tj[TJ_G * n2], bad);
  double d2 = 0.0;
  d[0] = d1;
  {
    WideA2Fwd p, q;
#pragma unroll
    for (int k = 0; k < kCh; ++k) {
      p.b[k] = b[1 + k];
      p.fg[k] = *reinterpret_cast<const double2*>(T + 6 * (1 + k));
      p.mm[k] = T[6 * (1 + k) + 2];
    }
#pragma unroll 1
    for (int jb = 1; jb <= m2; jb += 2 * kCh) {
      wide_a2_fwd_chunk(p, q, b, d, T, jb, d1, d2);
      wide_a2_fwd_chunk(q, p, b, d, T, jb + kCh, d1, d2);
    }
  }
  double x1 = 0.0, x2 = 0.0;
  {
    WideA2Bwd p, q;
#pragma unroll
    for (int k = 0; k < kCh; ++k) {
      p.d[k] = d[m2 - k];
      p.cc[k] = *reinterpret_cast<const double2*>(T + 6 * (m2 - k) + 4);
    }
#pragma unroll 1
    for (int jt = m2; jt >= 0; jt -= 2 * kCh) {
      wide_a2_bwd_chunk(p, q, b, d, T, jt, x1, x2);
      wide_a2_bwd_chunk(q, p, b, d, T, jt - kCh, x1, x2);
    }
  }
}
// Line r of a batch of nb lines runs on warp r % W, lane r / W, with W = min(nb, kChainWarps) warps: the lanes of a warp
// sweep in lockstep for free, while every extra WARP on a scheduler costs the others issue slots and FP64-pipe cycles
// (a chain node is ~22 instructions per 57-cycle chain step; an FP64 instruction takes the sub-partition's pipe for two
// cycles whatever its active lanes) — measured: 13 lines on 13 warps sweep 1.6 times slower than on 8.
constexpr int kChainWarps = 8;   // two per scheduler
__device__ __forceinline__ int wide_line_of_thread(int tid, int nb) {
  const int W = min(nb, kChainWarps);
  const int warp = tid >> 5, lane = tid & 31;
  return warp < W ? lane * W + warp : 0x7fffffff;
}

// ---- stages -------------------------------------------------------------------------------------------------------------
// A1 factor streams of owned row (batch slot r, row j) into the row buffers
__device__ __forceinline__ void wide_load_factors(const HadiItem& it, const HadiView& w, const WideGeo& g, int r, int j,
                                                  int tid) {
  const int m1 = w.m1;
  const double* fM = w.fM + (size_t)j * w.co_pi;
  const double* fB = w.fB + (size_t)j * w.co_pi * 2;
  const double vj = w.tj[TJ_V * w.n2 + j];
  const double theta = it.theta, dt = it.dt;
  for (int k = tid; k < m1; k += kWideThreads) {
    g.rA[r * g.pr + kPad + 1 + k].y = wld(fM + k);   // node i = k + 1 takes multiplier m_i
    g.rC[r * g.pr + kPad + k] = make_double2(wld(fB + 2 * k), wld(fB + 2 * k + 1));
  }
  for (int i = tid; i <= m1; i += kWideThreads) {
    const double a = hadi_ti(w, TI_HS2)[i] * vj;
    const double up = a * hadi_ti(w, TI_DSP)[i] + hadi_ti(w, TI_BBP)[i];
    g.rB[r * g.pr + kPad + i].y = -theta * dt * up;
  }
}

// kind: 0 Douglas explicit stage, 1 Craig-Sneyd-family predictor, 2 corrector
template <int KIND>
__device__ __forceinline__ void wide_rows(const HadiItem& it, const HadiView& w, const HadiCsView& cs, const WideGeo& g,
                                          double e0, double e1, int scheme, int tid, WideProf& pf) {
  const int m1 = w.m1, nc = m1 + 1;
  for (int b0 = 0; b0 < g.nrow; b0 += g.RB) {
    const int nb = min(g.RB, g.nrow - b0);
    if (!g.resident)
      for (int r = 0; r < nb; ++r) wide_load_factors(it, w, g, r, g.rank + g.G * (b0 + r), tid);
    wide_tick(pf, 7, tid);
    for (int idx = tid; idx < nb * nc; idx += kWideThreads) {
      const int r = idx / nc, i = idx - r * nc;
      const int j = g.rank + g.G * (b0 + r);
      double rhs;
      if (KIND == 0) rhs = wide_node_explicit(it, w, e0, e1, i, j);
      else if (KIND == 1) rhs = wide_node_predict(it, w, cs, e0, e1, i, j, scheme);
      else rhs = wide_node_correct(it, w, cs, e0, e1, i, j, scheme);
      g.rA[r * g.pr + kPad + i].x = rhs;
    }
    __syncthreads();
    wide_tick(pf, 2, tid);
    {
      const int r = wide_line_of_thread(tid, nb);
      if (r < nb)
        wide_chain_a1(g.rA + r * g.pr + kPad, g.rB + r * g.pr + kPad, g.rC + r * g.pr + kPad, m1);
    }
    __syncthreads();
    wide_tick(pf, 3, tid);
    for (int idx = tid; idx < nb * nc; idx += kWideThreads) {
      const int r = idx / nc, i = idx - r * nc;
      const int j = g.rank + g.G * (b0 + r);
      w.Y[j * w.ld + i] = g.rA[r * g.pr + kPad + i].x;
    }
    if (b0 + g.RB < g.nrow) __syncthreads();
    wide_tick(pf, 6, tid);
  }
}

// A2 sweeps of the owned columns.  douglas: the right-hand side re-derives A2 U from the old solution (hadi_phase_rhs2),
// the Dirichlet column of the put boundary set and the American projection follow the sweep (hadi_phase_project);
// otherwise the right-hand side takes the stored R2 (hadi_cs_rhs2).
__device__ __forceinline__ void wide_cols(const HadiItem& it, const HadiView& w, const HadiCsView& cs, const WideGeo& g,
                                          double e0, double e1, bool douglas, double g_dir, double rdt, int tid, WideProf& pf) {
  const int m2 = w.m2, ld = w.ld, n2 = w.n2, nr = m2 + 1;
  const double c = w.c, dt = it.dt;
  const bool am = douglas && it.style == 1;
  const double* tj = w.tj;
  for (int b0 = 0; b0 < g.ncol; b0 += g.CB) {
    const int nb = min(g.CB, g.ncol - b0);
    for (int idx = tid; idx < nb * nr; idx += kWideThreads) {
      const int j = idx / nb, cc = idx - j * nb;
      const int i = g.c0 + b0 + cc;
      const int q = j * ld + i;
      const double y = wld(w.Y + q);
      double v;
      if (douglas) {
        const double* p = w.U + q;
        double r2 = tj[TJ_L2 * n2 + j] * wld(p - 2 * ld) + tj[TJ_L1 * n2 + j] * wld(p - ld) + tj[TJ_D0 * n2 + j] * wld(p) +
                    tj[TJ_U1 * n2 + j] * wld(p + ld);
        r2 += tj[TJ_U2 * n2 + j] * wld(p + 2 * ld);
        const double b2 = (j == m2) ? hadi_ti(w, TI_B2V)[i] : 0.0;
        v = y + c * (b2 * e1 - (r2 + b2 * e0));
        if (am) g.cl[cc * g.pc + kPad + j] = wld(w.lam + q);
      } else {
        double b1p, b2p;
        hadi_cs_bounds(it, w, i, j, b1p, b2p);
        v = y + c * (b2p * e1 - (wld(cs.R2 + q) + b2p * e0));
      }
      g.cb[cc * g.pc + kPad + j] = v;
    }
    __syncthreads();
    wide_tick(pf, 5, tid);
    {
      const int cc = wide_line_of_thread(tid, nb);
      if (cc < nb)
        wide_chain_a2(g.cb + cc * g.pc + kPad, g.cd + cc * g.pc + kPad, g.t2 + 6 * kPad, tj, n2, m2);
    }
    __syncthreads();
    wide_tick(pf, 4, tid);
    for (int idx = tid; idx < nb * nr; idx += kWideThreads) {
      const int j = idx / nb, cc = idx - j * nb;
      const int i = g.c0 + b0 + cc;
      const int q = j * ld + i;
      double x = g.cb[cc * g.pc + kPad + j];
      if (douglas && it.bc && i == 0) x = g_dir;
      if (am) {
        unsigned bad = 0;
        const double u0 = hadi_ti(w, TI_PAY)[i];
        const double l = g.cl[cc * g.pc + kPad + j];
        const double ln = hadi_max(0.0, l + hadi_div<false, true>(u0 - x, dt, rdt, bad));
        w.lam[q] = (i == w.m1) ? 0.0 : ln;
        x = hadi_max(x - dt * l, u0);
      }
      w.U[q] = x;
    }
    if (b0 + g.CB < g.ncol) __syncthreads();
    wide_tick(pf, 6, tid);
  }
}

// Dividend jump (hadi_phase_div1 / div2) on the owned rows: the jump interpolates along s within a row, so every CTA jumps
// its rows through a shared-memory copy of the old row and writes U in place; the multiplier array is not touched.
// The caller follows with a team barrier (the explicit stage reads the neighbouring rows).
__device__ __forceinline__ void wide_dividend(const HadiView& w, const WideGeo& g, double amount, double pct, int tid) {
  const int m1 = w.m1, nc = m1 + 1;
  const double* s = hadi_ti(w, TI_S);
  for (int i = tid; i <= m1; i += kWideThreads) {
    int idx;
    double wt;
    hadi_dividend_index(s, m1, amount, pct, i, idx, wt);
    w.divk[i] = idx;
    hadi_ti(w, TI_DIVW)[i] = wt;
  }
  for (int b0 = 0; b0 < g.nrow; b0 += g.RB) {
    const int nb = min(g.RB, g.nrow - b0);
    for (int idx = tid; idx < nb * nc; idx += kWideThreads) {
      const int r = idx / nc, i = idx - r * nc;
      g.rA[r * g.pr + kPad + i].x = wld(w.U + (g.rank + g.G * (b0 + r)) * w.ld + i);
    }
    __syncthreads();
    for (int idx = tid; idx < nb * nc; idx += kWideThreads) {
      const int r = idx / nc, i = idx - r * nc;
      const double2* row = g.rA + r * g.pr + kPad;
      const int k = w.divk[i];
      const double wt = hadi_ti(w, TI_DIVW)[i];
      double val;
      if (k > 0)
        val = (1.0 - wt) * row[k - 1].x + wt * row[k].x;
      else if (k == 0)
        val = row[0].x;
      else
        val = 0.0;
      w.U[(g.rank + g.G * (b0 + r)) * w.ld + i] = val;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kWideThreads, 1) hadi_wide_kernel(const HadiLaunch L, const int G) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x;
  const int team = blockIdx.x / G, rank = blockIdx.x - team * G, n_teams = gridDim.x / G;
  const int m1 = L.m1, m2 = L.m2;
  HadiView w;
  w.m1 = m1; w.m2 = m2; w.P = (m1 + 1) * (m2 + 1);
  w.ld = L.ld; w.n1 = L.n1; w.n2 = L.n2; w.pj = L.pj;
  w.line_mul = G; w.line_off = rank;
  w.co_pi = (m1 + 7) & ~7;
  // shared memory: per-i tables | per-j tables | dividend index (unused) | arena (A2 assembly scratch during set-up, then line buffers)
  double* sp = smem;
  w.ti = sp; sp += (size_t)TI_COUNT * w.n1;
  w.tj = sp; sp += (size_t)TJ_COUNT * w.n2;
  w.divk = reinterpret_cast<int*>(sp); sp += w.n1 / 2;   // n1 ints (n1 is a multiple of four)
  double* arena = sp;
  unsigned dyn_smem;
  asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_smem));
  const int arena_doubles = (int)((dyn_smem - (unsigned)((char*)arena - (char*)smem)) / sizeof(double));   // what is left behind the tables
  double* scratch = L.scratch + (size_t)team * L.scratch_stride;
  const HadiScratchLayout gl = hadi_scratch_layout(m1, m2, w.ld, w.pj, true, L.scheme >= 1);
  double* Ualloc = scratch + gl.U;
  w.U = Ualloc + HADI_HALO * w.ld + 1;
  w.Y = scratch + gl.Y;
  w.fM = scratch + gl.fM;
  w.fB = scratch + gl.fB;
  w.lam = scratch + gl.lam;
  HadiCsView cs;
  cs.Y0 = scratch + gl.Y0; cs.R0 = scratch + gl.R0; cs.R1 = scratch + gl.R1; cs.R2 = scratch + gl.R2;

  WideGeo g;
  g.rank = rank; g.G = G;
  g.nrow = (rank <= m2) ? (m2 - rank) / G + 1 : 0;
  g.c0 = (int)(((long long)rank * (m1 + 1)) / G);
  g.ncol = (int)(((long long)(rank + 1) * (m1 + 1)) / G) - g.c0;
  // arena: A2 factor records | row buffers | column buffers (behind the row buffers when those stay resident)
  g.t2 = arena;
  const int t2_doubles = 6 * (w.n2 + 2 * kPad);
  double* lines = arena + t2_doubles;
  const int line_doubles = arena_doubles - t2_doubles;
  g.pr = (w.n1 + 2 * kPad) | 1;   // odd pitches: the lanes of a chain warp (one line each) hit different banks
  g.pc = (w.n2 + 2 * kPad) | 1;
  const int max_row = (m2 + G) / G, max_col = (m1 + G) / G;
  g.RB = min(max_row, line_doubles / (6 * g.pr));
  g.CB = min(max_col, line_doubles / (3 * g.pc));
  g.resident = max_row <= g.RB;
  g.rA = reinterpret_cast<double2*>(lines);
  g.rB = g.rA + (size_t)g.RB * g.pr;
  g.rC = g.rB + (size_t)g.RB * g.pr;
  g.cb = lines;
  // the row buffers of a resident team member hold its factors for the whole solve: the column buffers then sit behind them
  if (g.resident) {
    const int used = 6 * g.RB * g.pr;
    const int cb_behind = (line_doubles - used) / (3 * g.pc);
    if (cb_behind >= 1) {
      g.CB = min(max_col, cb_behind);
      g.cb = lines + used;
    } else {
      g.resident = false;   // no room left for even one column: the factors are reloaded per batch, the arena is shared
    }
  }
  g.cd = g.cb + (size_t)g.CB * g.pc;
  g.cl = g.cd + (size_t)g.CB * g.pc;

  WideTeam tm;
  tm.ctr = reinterpret_cast<unsigned*>(L.counter) + 16 + 8 * team;
  tm.epoch = 0;
  tm.G = G;
  const int gtid = rank * kWideThreads + tid, gnt = G * kWideThreads;
  for (int k = gtid; k < (m2 + 1 + 2 * HADI_HALO) * w.ld + 2; k += gnt) Ualloc[k] = 0.0;

  WideProf pf;
#ifdef HADI_PHASE_TIMING
  pf.p = L.prof ? L.prof + 8 * (size_t)blockIdx.x : nullptr;
  pf.t = clock64();
#endif
  for (int item = team; item < L.n_items; item += n_teams) {
    const HadiItem it = L.items[item];
    const double* sg = L.s_pool + it.s_off;
    const double* vg = L.v_pool + it.v_off;
    const double* eg = L.e_pool + it.e_off;
    w.c = it.theta * it.dt;
    const double rdt = hadi_rcp_prep(it.dt);
    {
      // tables and factorisation: hadi_phases.cuh, with the A2 assembly scratch in the (still unused) arena.  The view
      // is re-aimed in place: a by-value copy of it came out of nvcc 12.9 with line_mul / line_off / co_pi undefined.
      double* const y_keep = w.Y;
      w.Y = arena;
      w.ts_off = 0;
      hadi_phase_tables(it, w, sg, vg, tid, kWideThreads);
      __syncthreads();
      hadi_phase_factor(it, w, vg, tid, kWideThreads, kWideThreads - 1);
      __syncthreads();
      w.Y = y_keep;
      // A2 factor records of the column chains (the assembly scratch above is dead now; the records sit at the arena's start)
      for (int j = tid; j <= m2; j += kWideThreads) {
        double* rec = g.t2 + 6 * (kPad + j);
        rec[0] = w.tj[TJ_F * w.n2 + j]; rec[1] = w.tj[TJ_G * w.n2 + j]; rec[2] = w.tj[TJ_MM * w.n2 + j];
        rec[4] = w.tj[TJ_CP * w.n2 + j]; rec[5] = w.tj[TJ_C2P * w.n2 + j];
      }
    }
    for (int p = gtid; p < (m2 + 1) * (m1 + 1); p += gnt) {
      const int j = p / (m1 + 1), i = p - j * (m1 + 1);
      w.U[j * w.ld + i] = hadi_ti(w, TI_PAY)[i];
      if (it.style == 1) w.lam[j * w.ld + i] = 0.0;
    }
    if (g.resident)
      for (int r = 0; r < g.nrow; ++r) wide_load_factors(it, w, g, r, rank + G * r, tid);
    wide_tick(pf, 0, tid);
    wide_sync(tm, tid);
    wide_tick(pf, 1, tid);
    int div_cur = 0;
    for (int n = 1; n <= it.N; ++n) {
      if (it.nd > 0) {
        // device schedule: one dividend per step at most; extension: every dividend dated inside the step, in order
        for (;;) {
          const int hit = it.div_all ? hadi_dividend_next(n, it.dt, it.nd, L.div_dates, div_cur)
                                     : hadi_dividend_at(n, it.dt, it.nd, L.div_dates, div_cur);
          if (hit < 0) break;   // uniform across the team
          wide_dividend(w, g, L.div_amounts[hit], L.div_pcts[hit], tid);
          wide_sync(tm, tid);
          wide_tick(pf, 1, tid);
          if (!it.div_all) break;
        }
      }
      const double e0 = eg[n - 1], e1 = eg[n];
      if (L.scheme >= 1) {
        wide_rows<1>(it, w, cs, g, e0, e1, L.scheme, tid, pf);
        wide_sync(tm, tid);
        wide_tick(pf, 1, tid);
        wide_cols(it, w, cs, g, e0, e1, false, 0.0, rdt, tid, pf);   // Y2 -> U
        wide_sync(tm, tid);
        wide_tick(pf, 1, tid);
        wide_rows<2>(it, w, cs, g, e0, e1, L.scheme, tid, pf);
        wide_sync(tm, tid);
        wide_tick(pf, 1, tid);
        wide_cols(it, w, cs, g, L.scheme == HADI_SCHEME_HV ? e1 : e0, e1, false, 0.0, rdt, tid, pf);
        wide_sync(tm, tid);
        wide_tick(pf, 1, tid);
      } else {
        wide_rows<0>(it, w, cs, g, e0, e1, 0, tid, pf);
        wide_sync(tm, tid);
        wide_tick(pf, 1, tid);
        wide_cols(it, w, cs, g, e0, e1, true, it.bc ? it.K * eg[it.N + 1 + n] : 0.0, rdt, tid, pf);
        wide_sync(tm, tid);
        wide_tick(pf, 1, tid);
      }
    }
    if (gtid == 0) {
      if (L.out_stride <= 1) {
        L.out_values[it.out] = wld(w.U + it.idx_v * w.ld + it.idx_s);
      } else {
        double* o = L.out_values + (size_t)it.out * L.out_stride;
        o[0] = wld(w.U + it.idx_v * w.ld + it.idx_s);
        o[1] = wld(w.U + (it.aux & 0xffff) * w.ld + it.idx_s);
        o[2] = wld(w.U + ((it.aux >> 16) & 0xffff) * w.ld + it.idx_s);
      }
    }
    if (L.out_U != nullptr || L.out_lam != nullptr) {
      for (int p = gtid; p < (m2 + 1) * (m1 + 1); p += gnt) {
        const int j = p / (m1 + 1), i = p - j * (m1 + 1);
        const size_t o = (size_t)it.out * w.P + (size_t)p;
        if (L.out_U != nullptr) L.out_U[o] = wld(w.U + j * w.ld + i);
        if (L.out_lam != nullptr && it.style == 1) L.out_lam[o] = wld(w.lam + j * w.ld + i);
      }
    }
    wide_sync(tm, tid);   // everyone is done with U and the tables before the next item
  }
}

size_t wide_table_bytes(int n1, int n2) {
  return sizeof(double) * ((size_t)TI_COUNT * n1 + (size_t)TJ_COUNT * n2) + sizeof(int) * (size_t)n1;
}

}  // namespace

// Shared memory: the tables plus an arena that must take the A2 assembly scratch, one row batch and one column batch.
int hadi_wide_plan(int device, int m1, int m2, int ld, int n1, int n2, int pj, HadiPlan* plan) {
  (void)ld; (void)pj;
  int max_smem = 0, sms = 0;
  cudaError_t e = cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) return (int)e;
  if (m2 + 1 > kWideThreads - 1 || m1 + 1 > 4096) return -1;
  const size_t tables = wide_table_bytes(n1, n2);
  const size_t pr = ((size_t)n1 + 2 * kPad) | 1, pc = ((size_t)n2 + 2 * kPad) | 1;
  const size_t need = tables + sizeof(double) * std::max((size_t)TS_COUNT * n2, 6 * ((size_t)n2 + 2 * kPad) + 6 * pr + 3 * pc);
  if (need + 1024 > (size_t)max_smem) return -1;
  const size_t smem = ((size_t)max_smem - 1024) & ~size_t(127);   // take the SM: one CTA per SM, the arena as large as it gets
  e = cudaFuncSetAttribute((const void*)hadi_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)hadi_wide_kernel, kWideThreads, smem);
  if (e != cudaSuccess) return (int)e;
  if (occ < 1) return -1;
  plan->global_state = true;
  plan->variant = HADI_WIDE_VARIANT;
  plan->threads = kWideThreads;
  plan->ctas_per_sm = 1;
  plan->sm_count = sms;
  plan->smem_bytes = smem;
  plan->cluster = 1;   // team size: set per batch by the host layer (hadi_wide_team)
  plan->duo = 1;
  return 0;
}

// CTAs per solve for a batch of n_items on sm_count SMs: everything the GPU has, but no more CTAs than lines
int hadi_wide_team(int n_items, int sm_count, int m1, int m2) {
  const int lines = std::max(m1 + 1, m2 + 1);
  int G = std::max(1, std::min(sm_count / std::max(1, n_items), lines));
  if (const char* t = getenv("HADI_WIDE_TEAM"))   // development aid: cap the team size
    if (atoi(t) > 0) G = std::min(G, atoi(t));
  return G;
}

int hadi_launch_wide(const HadiLaunch& L_in, const HadiPlan& plan, int grid_ctas, void* stream) {
  HadiLaunch L = L_in;
  int G = plan.cluster;
  if (G < 1 || grid_ctas % G != 0) return (int)cudaErrorInvalidValue;
  cudaError_t e = cudaFuncSetAttribute((const void*)hadi_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);
  if (e != cudaSuccess) return (int)e;
  void* args[] = {(void*)&L, (void*)&G};
  // co-operative launch: the runtime refuses a grid whose CTAs cannot all be resident (the team barrier needs them)
  e = cudaLaunchCooperativeKernel((const void*)hadi_wide_kernel, dim3((unsigned)grid_ctas), dim3(kWideThreads), args, plan.smem_bytes,
                                  (cudaStream_t)stream);
  return (int)e;
}
