// hadi — grid-specialised (compile-time m1, m2), device-only forms of the two implicit line-solve phases.
//
// The arithmetic of every grid node is the arithmetic of hadi_phases.cuh (same operations, same order,
// -fmad=false), so results are bit-identical; what changes is how operands reach the dependent chains.
// Measured on B200 (tools/ubench_lat.cu): a dependent FP64 op issues every 8.1 cycles, a shared-memory
// load returns after ~29 cycles, an L2 hit after ~300.  A line solve is one dependent chain per line
// (2 + 5 ops per node along s, 4 + 3 along v), so any load whose latency lands on the chain multiplies
// the phase time.  Both phases are therefore written as fully unrolled straight-line code in which
// every operand is requested a fixed number of elements before the chain needs it:
//
//   S1 (A1, tridiagonal along s, one line per v-row; factors are per node and live in L2):
//      "co-operative warps".  NW = ceil((m2+1)/13) warps each own 13 v-rows.  Lanes 0..12 run the 13
//      dependent chains; lanes 0..25 double as loaders: each keeps 32 bytes per block of the factor
//      streams in flight in its registers (HADI_CO_DF / HADI_CO_DB blocks ahead — the register file of
//      the otherwise idle lanes is the latency buffer), drops a block into a 2-slot per-warp staging
//      area in shared memory when the chain lanes have left it, and the chain lanes pick their operands
//      up from there PF / PB elements ahead.  Only __syncwarp() is needed; there is no mbarrier, no
//      producer warp and no ring handshake on the chain.
//   S2 (A2, pentadiagonal along v, one line per s-column; factors are per row and live in shared
//      memory): one thread per column, operands PF elements ahead, no branches.
//
// Reference: src/hes_a1_kernels.hpp:139-161 (Thomas), src/hes_a2_shuffled_kernels.hpp:243-299.
#pragma once
#include "hadi_phases.cuh"

#if defined(__CUDACC__)

#define HADI_CO_NR 13       // v-rows (chains) per chain warp
#define HADI_CO_BLK 4       // nodes per staged block
#ifndef HADI_CO_SF
#define HADI_CO_SF 5        // forward staging slots  (5 x 4 nodes:  ~17 nodes of slack at 16 cycles per node)
#endif
#ifndef HADI_CO_SB
#define HADI_CO_SB 3        // backward staging slots (3 x 4 nodes:  ~ 9 nodes of slack at 40 cycles per node)
#endif
#ifndef HADI_CO_DP
#define HADI_CO_DP 6        // blocks a feeder warp keeps in flight in its registers
#endif
#define HADI_CO_FSLOT (HADI_CO_NR * 5)    // doubles: 13 rows x pitch 5  (4 multipliers)
#define HADI_CO_BSLOT (HADI_CO_NR * 10)   // doubles: 13 rows x pitch 10 (4 x {pivot, reciprocal})

// geometry of the factor streams in per-CTA global scratch
//   cM [rows][PI]     cM[j][i-1]      = Thomas multiplier m(j,i),            i = 1..m1
//   cB [rows][PI][2]  cB[j][m1-i][..] = pivot(j,i), prepared reciprocal,     i = m1..1  (consumption order)
// rows = 13 * NW (rows past m2 are never written; the lanes that read them never store).
HADI_HD constexpr int hadi_co_pi(int m1) { return (m1 + 7) & ~7; }
HADI_HD constexpr int hadi_co_warps(int m2) { return (m2 + HADI_CO_NR) / HADI_CO_NR; }
HADI_HD constexpr int hadi_co_rows(int m2) { return hadi_co_warps(m2) * HADI_CO_NR; }
HADI_HD constexpr int hadi_co_warp_doubles() {
  return (HADI_CO_SF * HADI_CO_FSLOT > HADI_CO_SB * HADI_CO_BSLOT) ? HADI_CO_SF * HADI_CO_FSLOT : HADI_CO_SB * HADI_CO_BSLOT;
}
// staging area: per chain warp max(forward ring, backward ring), then 2 ints per chain warp (staged, consumed)
HADI_HD constexpr int hadi_co_stage_bytes(int m2) {
  return hadi_co_warps(m2) * (hadi_co_warp_doubles() * 8 + 8);
}

__device__ __forceinline__ int hadi_ldv(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void hadi_stv(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// Block sequence of one A1 solve: NBF forward blocks (4 multipliers per row), then NBB backward blocks
// (4 x {pivot, reciprocal} per row).  Blocks are numbered globally over the item, g = solve * NU + u, and the
// two flags count them: staged = blocks published by the feeder, consumed = blocks the chain warp has left.
template <int M1>
struct HadiCoSeq {
  static constexpr int B = HADI_CO_BLK;
  static constexpr int NBF = (M1 + B - 1) / B, NBB = (M1 + B - 1) / B, NU = NBF + NBB;
};

// Feeder side.  `first`..`last` (block numbers within the solve, half-open) are staged; blocks before `first`
// were pre-staged by an earlier call (hadi_fast_prestage).
template <int M1, int M2>
__device__ __forceinline__ void hadi_co_feed(const HadiView& w, int c, int lane, int solve, int first, int last) {
  using Q = HadiCoSeq<M1>;
  constexpr int NR = HADI_CO_NR, PI = hadi_co_pi(M1), DP = HADI_CO_DP;
  constexpr int SF = HADI_CO_SF, SB = HADI_CO_SB;
  // loader lanes 0..25 cover 13 rows x 2 halves of a block (2 nodes = 16 B forward, 32 B backward); lanes 26..31
  // repeat lanes 24/25 (same addresses, same values: no extra traffic and no divergence)
  const int lrow = (lane < 2 * NR) ? (lane >> 1) : NR - 1, half = lane & 1;
  double* stg = w.stg + c * hadi_co_warp_doubles();
  int* flg = reinterpret_cast<int*>(w.stg + hadi_co_warps(M2) * hadi_co_warp_doubles()) + 2 * c;
  const double* gM = w.fM + (size_t)(c * NR + lrow) * PI + 2 * half;
  const double* gB = w.fB + ((size_t)(c * NR + lrow) * PI + 2 * half) * 2;
  double* sWf = stg + lrow * 5 + 2 * half;
  double* sWb = stg + lrow * 10 + 4 * half;
  const int base = solve * Q::NU;
  double2 R[DP][2];
#define HADI_FEED_LOAD(u)                                                                              \
  if ((u) < last) {                                                                                    \
    if ((u) < Q::NBF) {                                                                                \
      R[(u) % DP][0] = __ldcg(reinterpret_cast<const double2*>(gM + (u) * 4));                        \
    } else {                                                                                           \
      const double2* p_ = reinterpret_cast<const double2*>(gB + ((u) - Q::NBF) * 8);                  \
      R[(u) % DP][0] = __ldcg(p_);                                                                     \
      R[(u) % DP][1] = __ldcg(p_ + 1);                                                                 \
    }                                                                                                  \
  }
#pragma unroll
  for (int u = 0; u < Q::NU + DP; ++u) {
    if (u >= first && u < first + DP) { HADI_FEED_LOAD(u) }   // pipeline fill
  }
#pragma unroll
  for (int u = 0; u < Q::NU; ++u) {
    if (u >= first && u < last) {
      // the slot this block overwrites must have been left by the chain warp
      int need;
      if (u < Q::NBF) need = base + ((u - SF + 1 > 0) ? u - SF + 1 : 0);                  // forward block u - SF consumed
      else need = base + Q::NBF + ((u - Q::NBF - SB + 1 > 0) ? u - Q::NBF - SB + 1 : 0);  // all forward + block - SB
      while (hadi_ldv(flg + 1) < need) {}
      if (u < Q::NBF) {
        double* s_ = sWf + (u % SF) * HADI_CO_FSLOT;
        s_[0] = R[u % DP][0].x;
        s_[1] = R[u % DP][0].y;
      } else {
        double2* s_ = reinterpret_cast<double2*>(sWb + ((u - Q::NBF) % SB) * HADI_CO_BSLOT);
        s_[0] = R[u % DP][0];
        s_[1] = R[u % DP][1];
      }
      HADI_FEED_LOAD(u + DP)
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        hadi_stv(flg, base + u + 1);
      }
    }
  }
#undef HADI_FEED_LOAD
}

// Blocks of the NEXT solve that are staged ahead of it (while the feeder warps are idle in phase S2, or during
// item set-up): the whole forward ring.
template <int M1, int M2>
__device__ __forceinline__ void hadi_fast_prestage(const HadiView& w, int tid, int solve) {
  constexpr int NW = hadi_co_warps(M2);
  const int warp = tid >> 5, lane = tid & 31;
  if (warp < NW || warp >= 2 * NW) return;
  hadi_co_feed<M1, M2>(w, warp - NW, lane, solve, 0, HADI_CO_SF);
}

template <int M1, int M2, bool EXACT>
__device__ __forceinline__ void hadi_fast_solve_a1(const HadiItem& it, const HadiView& w, int solve, int tid,
                                                   unsigned& bad, long long* dbg = nullptr) {
  using Q = HadiCoSeq<M1>;
  constexpr int NR = HADI_CO_NR, B = HADI_CO_BLK, SF = HADI_CO_SF, SB = HADI_CO_SB;
  constexpr int PF = 3, PB = 2;                         // shared-memory lookahead of the chains (nodes)
  constexpr int NW = hadi_co_warps(M2);
  constexpr int LD = hadi_geo_ld(M1), N1 = hadi_geo_n1(M1), N2 = hadi_geo_n2(M2);
  const int warp = tid >> 5, lane = tid & 31;
  if (warp >= 2 * NW) return;
  if (warp >= NW) {
    hadi_co_feed<M1, M2>(w, warp - NW, lane, solve, HADI_CO_SF, Q::NU);
    return;
  }
  // ---- chain warp: lanes 0..12 own one v-row each.  Idle lanes shadow row 0 of the warp: same operands, same
  // results, stored to the same words (a store predicated on the lane makes the compiler split the sweep
  // into two divergent copies).
  const int c = warp;
  const bool chain = (lane < NR) && (c * NR + lane <= M2);
  const int jr = chain ? lane : 0;
  const int j = c * NR + jr;
  double* stg = w.stg + c * hadi_co_warp_doubles();
  int* flg = reinterpret_cast<int*>(w.stg + NW * hadi_co_warp_doubles()) + 2 * c;
  const double* sRf = stg + jr * 5;
  const double* sRb = stg + jr * 10;
  double* y = w.Y + j * LD;
  const double vj = w.tj[TJ_V * N2 + j];
  const double* hs2 = w.ti + TI_HS2 * N1;
  const double* dsp = w.ti + TI_DSP * N1;
  const double* bbp = w.ti + TI_BBP * N1;
  const double ntd = -it.theta * it.dt;
  const int base = solve * Q::NU;
#ifdef HADI_PHASE_TIMING
  const long long dbg_t0 = clock64();
#endif
  // ---- forward elimination: x_i = y_i - m_i x_{i-1}, i = 1..m1 (element e <-> i = e + 1, block e / 4)
  double x = y[0];
  {
    double mq[PF], yq[PF];
    int fl = hadi_ldv(flg);
    while (fl < base + 1) fl = hadi_ldv(flg);
    // Every load of a staged block is address-dependent on the flag value that admitted it (`zmask` is a
    // run-time zero): neither ptxas — which may move plain shared loads across volatile ones and did hoist
    // them above the wait loop — nor the hardware can then perform it before the flag load has returned.
    const double* sD = sRf + (fl & w.zmask);
#pragma unroll
    for (int e = 0; e < PF; ++e) {
      mq[e] = sD[((e / B) % SF) * HADI_CO_FSLOT + (e % B)];
      yq[e] = y[e + 1];
    }
#pragma unroll
    for (int e = 0; e < M1; ++e) {
      const int f = e + PF;          // element whose operands are requested now
      const double mc = mq[e % PF], yc = yq[e % PF];
      if (f < M1) {
        if (f % B == 0) {
          // first node of block f / B: the feeder must have published it (flag read two nodes ago)
          while (fl < base + f / B + 1) fl = hadi_ldv(flg);
          sD = sRf + (fl & w.zmask);
        }
        mq[e % PF] = sD[((f / B) % SF) * HADI_CO_FSLOT + (f % B)];
        yq[e % PF] = y[f + 1];
        if (f % B == B - 2 && f + 2 < M1) fl = hadi_ldv(flg);   // early probe for the next block
      }
      x = yc - mc * x;
      y[e + 1] = x;
      // block e / B has been left once its last node has been used (the published value depends on it)
      if (e % B == B - 1 || e == M1 - 1) {
        if (lane == 0) hadi_stv(flg + 1, base + e / B + 1 + (__double2hiint(x) & w.zmask));
      }
    }
  }
#ifdef HADI_PHASE_TIMING
  if (dbg) dbg[0] += clock64() - dbg_t0;
#endif
  // ---- back substitution: x_i = (x_i - impl_upper(j,i) x_{i+1}) / pivot(j,i), i = m1..1 (element e <-> i = m1 - e)
  // impl_upper(j,i) = -theta*dt*(a*delta_s(+1) + b*beta_s(+1)) from the zero-padded tables (exactly 0 at i = m1)
  {
    double tq[PB], rq[PB], yq[PB], hq[PB], dq[PB], bq[PB];
    const int bb = base + Q::NBF;
    int fl = hadi_ldv(flg);
    while (fl < bb + 1) fl = hadi_ldv(flg);
    const double* sD = sRb + (fl & w.zmask);
#pragma unroll
    for (int e = 0; e < PB; ++e) {
      const double2 tr = *reinterpret_cast<const double2*>(sD + ((e / B) % SB) * HADI_CO_BSLOT + 2 * (e % B));
      tq[e] = tr.x; rq[e] = tr.y;
      yq[e] = y[M1 - e];
      hq[e] = hs2[M1 - e]; dq[e] = dsp[M1 - e]; bq[e] = bbp[M1 - e];
    }
    double iu = ntd * ((hq[0] * vj) * dq[0] + bq[0]);
    double xn = 0.0;
    unsigned badl = 0;
#pragma unroll
    for (int e = 0; e < M1; ++e) {
      const int f = e + PB;
      const double tc = tq[e % PB], rc = rq[e % PB], yc = yq[e % PB];
      double iun = 0.0;
      if (e + 1 < M1) iun = ntd * ((hq[(e + 1) % PB] * vj) * dq[(e + 1) % PB] + bq[(e + 1) % PB]);
      if (f < M1) {
        if (f % B == 0) {
          while (fl < bb + f / B + 1) fl = hadi_ldv(flg);
          sD = sRb + (fl & w.zmask);
        }
        const double2 tr = *reinterpret_cast<const double2*>(sD + ((f / B) % SB) * HADI_CO_BSLOT + 2 * (f % B));
        tq[e % PB] = tr.x; rq[e % PB] = tr.y;
        yq[e % PB] = y[M1 - f];
        hq[e % PB] = hs2[M1 - f]; dq[e % PB] = dsp[M1 - f]; bq[e % PB] = bbp[M1 - f];
        if (f % B == B - 2 && f + 2 < M1) fl = hadi_ldv(flg);
      }
      x = hadi_div<EXACT>(yc - iu * xn, tc, rc, badl);
      xn = x;
      y[M1 - e] = x;
      iu = iun;
      if (e % B == B - 1 || e == M1 - 1) {
        if (lane == 0) hadi_stv(flg + 1, bb + e / B + 1 + (__double2hiint(x) & w.zmask));
      }
    }
    bad |= badl;
  }
}

// ----------------------------------------------------------------------------------------------
// R + S2 fused: (I - theta*dt*A2) U = Y1 + theta*dt*(b2*e1 - (A2 U + b2*e0)), one thread per s-column on the
// natural layout (stride LD).  Phase R (src/device_solver.hpp:254-260) is point-wise in the column, so the
// thread that is about to start the dependent forward chain of node j forms that node's right-hand side on
// the way: A2 U from a five-row register window of its own column of U (old solution: the back substitution
// below is what overwrites it, after the forward sweep has read all of it), with the per-row coefficients
// coming as packed 64-byte records {L2, L1, D0, U1, U2, F, G, MM} (4 broadcast LDS.128 per node instead of 8
// LDS.64).  This removes one pass over the grid and one __syncthreads() per step; the extra 15 FP64 operations
// per node ride in the issue slots the 4-deep dependent chain leaves free.  Same operations, same order as
// hadi_phase_rhs2 + hadi_phase_solve_a2.
#ifdef HADI_SPLIT_R
#define HADI_FUSE_R 0
#else
#define HADI_FUSE_R 1
#endif
template <int M1, int M2, bool EXACT>
__device__ __forceinline__ void hadi_fast_solve_a2(const HadiItem& it, const HadiView& w, double e0, double e1,
                                                   int tid, unsigned& bad) {
  constexpr int LD = hadi_geo_ld(M1), N1 = hadi_geo_n1(M1), N2 = hadi_geo_n2(M2);
  constexpr int PF = 2;
  if (tid > M1) return;
  const double* CP = w.tj + TJ_CP * N2;
  const double* C2P = w.tj + TJ_C2P * N2;
  const double2* rec = reinterpret_cast<const double2*>(w.tjp);
  double* Yc = w.Y + tid;
  double* Uc = w.U + tid;
  const double c = w.c;
  const double b2v = w.ti[TI_B2V * N1 + tid];
  (void)it;
  // ---- forward sweep: d_0 = b_0 / impl_main(0);  d_j = (b_j - f_j d_{j-1} - g_j d_{j-2}) * m_j
  // operand queue, PF nodes ahead: y_j, U[j+2] (the window's new row) and the row record
  double yq[PF + 1], uq[PF + 1];
  double2 ra[PF + 1], rb[PF + 1], rc[PF + 1], rd[PF + 1];
#pragma unroll
  for (int k = 0; k <= PF; ++k) {
    yq[k] = Yc[k * LD];
    uq[k] = HADI_FUSE_R ? Uc[(k + 2) * LD] : 0.0;
    ra[k] = rec[4 * k]; rb[k] = rec[4 * k + 1]; rc[k] = rec[4 * k + 2]; rd[k] = rec[4 * k + 3];
  }
  double um2 = 0.0, um1 = 0.0, u0 = 0.0, up1 = 0.0;     // U[j-2], U[j-1], U[j], U[j+1] (halo rows are zero)
  if (HADI_FUSE_R) { u0 = Uc[0]; up1 = Uc[LD]; }
  double d1 = 0.0, d2 = 0.0;
#pragma unroll
  for (int j = 0; j <= M2; ++j) {
    const int s = j % (PF + 1);
    const double yc = yq[s], up2 = uq[s];
    const double2 A = ra[s], B = rb[s], C = rc[s], D = rd[s];   // {L2,L1} {D0,U1} {U2,F} {G,MM}
    if (j + PF + 1 <= M2) {
      const int f = j + PF + 1;
      yq[s] = Yc[f * LD];
      uq[s] = HADI_FUSE_R ? Uc[(f + 2) * LD] : 0.0;
      ra[s] = rec[4 * f]; rb[s] = rec[4 * f + 1]; rc[s] = rec[4 * f + 2]; rd[s] = rec[4 * f + 3];
    }
    double bc = yc;
    if (HADI_FUSE_R) {
      double r2 = A.x * um2 + A.y * um1 + B.x * u0 + B.y * up1;
      r2 += C.x * up2;
      const double b2 = (j == M2) ? b2v : 0.0;
      bc = yc + c * (b2 * e1 - (r2 + b2 * e0));
      um2 = um1; um1 = u0; u0 = up1; up1 = up2;
    }
    double d;
    if (j == 0) {
      d = hadi_div<EXACT>(bc, D.y, D.x, bad);       // row 0: MM[0] holds impl_main(0), G[0] its prepared reciprocal
    } else {
      d = (bc - C.y * d1 - D.x * d2) * D.y;
    }
    Yc[j * LD] = d;
    d2 = d1;
    d1 = d;
  }
  // ---- back substitution: x_j = d_j - c'_j x_{j+1} - c2'_j x_{j+2}, j = m2..0 (element e <-> j = m2 - e)
  // (compiler fence: with every address static the compiler would otherwise forward all m2+1 stored d_j to
  //  the loads below, i.e. keep them live in registers and spill them to local memory)
  asm volatile("" ::: "memory");
  constexpr int PB = 3;
  double dq[PB], cq[PB], eq[PB];
  dq[0] = d1; cq[0] = CP[M2]; eq[0] = C2P[M2];
  dq[1] = d2; cq[1] = CP[M2 - 1]; eq[1] = C2P[M2 - 1];
#pragma unroll
  for (int k = 2; k < PB; ++k) {
    dq[k] = Yc[(M2 - k) * LD]; cq[k] = CP[M2 - k]; eq[k] = C2P[M2 - k];
  }
  double x1 = 0.0, x2 = 0.0;
#pragma unroll
  for (int e = 0; e <= M2; ++e) {
    const int s = e % PB;
    const double dc = dq[s], cc = cq[s], c2 = eq[s];
    if (e + PB <= M2) {
      dq[s] = Yc[(M2 - e - PB) * LD]; cq[s] = CP[M2 - e - PB]; eq[s] = C2P[M2 - e - PB];
    }
    const double xv = dc - cc * x1 - c2 * x2;
    x2 = x1;
    x1 = xv;
    Uc[(M2 - e) * LD] = xv;
  }
}

#endif  // __CUDACC__
