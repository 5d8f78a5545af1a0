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
// S1 with the back-substitution factors in TENSOR MEMORY (FEED 5).
//
// The 256 KB of tensor memory of a Blackwell SM are idle in this kernel (no tcgen05.mma anywhere), and the
// (pivot, prepared reciprocal) stream of the Thomas back substitution — 2 x (m2+1) x m1 doubles, 83 KB at
// 101 x 51 — is what phase S1 waits for when it lives in L2 (13 k of the 17 k cycles of S1).
// A CTA therefore allocates 256 (101 x 51) or 128 (51 x 26) TMEM columns at start-up — two / three CTAs per SM
// fit — and keeps that stream there for the life of an item: tcgen05.ld returns in a few tens of cycles and
// touches neither shared memory nor the L2.
//
// Layout.  TMEM is addressed as (lane 0..127, column): warp w reaches lanes 32*(w%4) .. +31, each thread its
// own lane (shape .32x32b).  Chain warp c (rows 13c .. 13c+12, as in the co-operative variant) uses lanes
// l = 0..12 and l + 16: lane l holds the {pivot, reciprocal} pairs of the FIRST half of row 13c+l's back
// substitution (nodes i = m1 .. m1/2+1) at columns 4e .. 4e+3, e = 0 .. m1/2-1, lane l + 16 the pairs of the
// SECOND half (i = m1/2 .. 1) at the same columns.  The sweep runs the same fully unrolled half-row code twice:
// first lanes 0..15 are live, then x is handed to lane l + 16 by one shuffle and lanes 16..31 are live; the
// lanes that are not live work on a dummy row of shared memory, so that no store is predicated (a predicated
// store makes the compiler split the sweep into divergent copies).  Per node the chain lane issues no shuffle
// and no global load; one tcgen05.ld.x16 per four nodes is requested a chunk ahead.
__device__ __forceinline__ void hadi_tm_alloc(unsigned* slot_smem, int cols) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(slot_smem);
  if (cols == 512) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(a) : "memory");
  else if (cols == 256) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(a) : "memory");
  else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(a) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void hadi_tm_free(unsigned addr, int cols) {
  if (cols == 512) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(addr) : "memory");
  else if (cols == 256) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(addr) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void hadi_tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void hadi_tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void hadi_tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// one {pivot, reciprocal} pair of this thread's lane
__device__ __forceinline__ void hadi_tm_st_pair(unsigned taddr, double t, double r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__double2loint(t)),
               "r"(__double2hiint(t)), "r"(__double2loint(r)), "r"(__double2hiint(r))
               : "memory");
}
// four pairs (16 columns) of this thread's lane; the registers may be read after hadi_tm_wait_ld(q)
struct HadiTmQuad { unsigned r[16]; };
__device__ __forceinline__ void hadi_tm_ld_quad(unsigned taddr, HadiTmQuad& q) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(q.r[0]), "=r"(q.r[1]), "=r"(q.r[2]), "=r"(q.r[3]), "=r"(q.r[4]), "=r"(q.r[5]), "=r"(q.r[6]), "=r"(q.r[7]),
        "=r"(q.r[8]), "=r"(q.r[9]), "=r"(q.r[10]), "=r"(q.r[11]), "=r"(q.r[12]), "=r"(q.r[13]), "=r"(q.r[14]), "=r"(q.r[15])
      : "r"(taddr)
      : "memory");
}
// the registers are operands of the wait: no use of them can be scheduled above it
__device__ __forceinline__ void hadi_tm_wait_ld(HadiTmQuad& q) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(q.r[0]), "+r"(q.r[1]), "+r"(q.r[2]), "+r"(q.r[3]), "+r"(q.r[4]), "+r"(q.r[5]), "+r"(q.r[6]), "+r"(q.r[7]),
                 "+r"(q.r[8]), "+r"(q.r[9]), "+r"(q.r[10]), "+r"(q.r[11]), "+r"(q.r[12]), "+r"(q.r[13]), "+r"(q.r[14]), "+r"(q.r[15])
               :
               : "memory");
}
__device__ __forceinline__ double hadi_tm_piv(const HadiTmQuad& q, int k) { return __hiloint2double((int)q.r[4 * k + 1], (int)q.r[4 * k]); }
__device__ __forceinline__ double hadi_tm_rcp(const HadiTmQuad& q, int k) { return __hiloint2double((int)q.r[4 * k + 3], (int)q.r[4 * k + 2]); }

HADI_HD constexpr int hadi_tm_half(int m1) { return m1 / 2; }
// columns a CTA allocates: 4 per node of a half row, rounded up to whole x16 loads, then to a power of two
HADI_HD constexpr int hadi_tm_cols(int m1) { return (16 * ((hadi_tm_half(m1) + 3) / 4) <= 128) ? 128 : 256; }
HADI_HD constexpr int hadi_tm_cols_feed(int feed, int m1) { return feed == 6 ? 512 : feed == 7 ? 256 : hadi_tm_cols(m1); }

// Thomas factors of I - theta*dt*A1 (src/hes_a1_kernels.hpp:145-152), chain warps only, all 32 lanes converged.
// Pass A runs nodes 1 .. m1/2 on every lane; in pass B lanes 0..15 go on with nodes m1/2+1 .. m1 while lanes
// 16..31 start the recurrence again at node 1, so that in iteration k both groups hold the pair that belongs
// at columns 4(m1/2 - k) of their own lane and one tcgen05.st serves both.  Multipliers go to fM (global, read
// by the forward sweep); lanes that own no row shadow row 13c (same values to the same words).
template <int M1, int M2>
__device__ __forceinline__ void hadi_tmem_factor(const HadiItem& it, const HadiView& w, const double* vg, int tid,
                                                 unsigned tmem) {
  constexpr int NR = HADI_CO_NR, NW = hadi_co_warps(M2), H = hadi_tm_half(M1);
  constexpr int N1 = hadi_geo_n1(M1), PJ = hadi_geo_pj(M2);
  static_assert(M1 % 2 == 0 && NW <= 4, "half-row layout needs an even m1 and at most four chain warps");
  const int warp = tid >> 5, lane = tid & 31;
  if (warp >= NW) return;
  const int g = lane >> 4, l = lane & 15;
  const bool own = (l < NR) && (warp * NR + l <= M2);
  const int j = warp * NR + (own ? l : 0);
  const unsigned mine = tmem + ((unsigned)(32 * warp) << 16);
  const double vj = vg[j];
  const double* hs2 = w.ti + TI_HS2 * N1;
  const double* dsm = w.ti + TI_DSM * N1;
  const double* ds0 = w.ti + TI_DS0 * N1;
  const double* dsp = w.ti + TI_DSP * N1;
  const double* sv = w.ti + TI_S * N1;
  const double* bsm = w.ti + TI_BSM * N1;
  const double* bs0 = w.ti + TI_BS0 * N1;
  const double* bbp = w.ti + TI_BBP * N1;
  const double theta = it.theta, dt = it.dt;
  const double rdiff = it.r_d - it.r_f;
  double t = 1.0;           // impl_main(j,0)
  double iu_prev = 0.0;     // impl_upper(j,0)
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    int i0 = 0;             // this lane handles node i0 + k in iteration k
    if (pass == 1) {
      if (g == 0) i0 = H;
      else { t = 1.0; iu_prev = 0.0; }
    }
#pragma unroll 1
    for (int k = 1; k <= H; ++k) {
      const int i = i0 + k;
      double il, im, iu;
      if (i < M1) {
        const double a = hs2[i] * vj;
        const double b = rdiff * sv[i];
        const double lo = a * dsm[i] + b * bsm[i];
        const double ma = a * ds0[i] + b * bs0[i] - 0.5 * it.r_d;
        const double up = a * dsp[i] + bbp[i];
        il = -theta * dt * lo;
        im = 1.0 - theta * dt * ma;
        iu = -theta * dt * up;
      } else {
        const double ma = -0.5 * it.r_d;
        il = 0.0;
        im = 1.0 - theta * dt * ma;
        iu = 0.0;
      }
      const double m = il / t;
      t = im - m * iu_prev;
      iu_prev = iu;
      w.fM[(size_t)(i - 1) * PJ + j] = m;
      // node i is element e = M1 - i of the back substitution: e - H*(1-g) = H - k for both groups
      if (pass == 1) hadi_tm_st_pair(mine + 4u * (unsigned)(H - k), t, hadi_rcp_prep(t));
    }
  }
  hadi_tm_wait_st();
}

// Phase S1 of the FEED 5 variants.  Forward sweep: multipliers from fM (L2) in chunks of HADI_KF nodes, as in
// hadi_phase_solve_a1.  Back substitution: two half rows fed from tensor memory (see above).
// `dummy`: m1 + 2 doubles of shared memory nobody reads.
template <int M1, int M2, bool EXACT>
__device__ __forceinline__ void hadi_tmem_solve_a1(const HadiItem& it, const HadiView& w, int tid, unsigned& bad,
                                                   unsigned tmem, double* dummy, long long* dbg = nullptr) {
  constexpr int NR = HADI_CO_NR, NW = hadi_co_warps(M2), H = hadi_tm_half(M1), NQ = (H + 3) / 4;
  constexpr int LD = hadi_geo_ld(M1), N1 = hadi_geo_n1(M1), N2 = hadi_geo_n2(M2), PJ = hadi_geo_pj(M2);
  constexpr int KF = HADI_KF;
  const int warp = tid >> 5, lane = tid & 31;
  if (warp >= NW) return;
  const int g = lane >> 4, l = lane & 15;
  const bool own = (l < NR) && (warp * NR + l <= M2);
  const int j = warp * NR + (own ? l : 0);
  const unsigned mine = tmem + ((unsigned)(32 * warp) << 16);
  double* y = w.Y + j * LD;
  const double vj = w.tj[TJ_V * N2 + j];
  const double ntd = -it.theta * it.dt;
#ifdef HADI_PHASE_TIMING
  const long long dbg_t0 = clock64();
#endif
  // first quad of the back substitution: in flight during the whole forward sweep
  HadiTmQuad qa, qb;
  hadi_tm_ld_quad(mine, qa);
  // ---- forward elimination: x_i = y_i - m_i x_{i-1}, i = 1..m1 (both lane groups run the row: same values)
  double x = y[0];
  {
    const double* pm = w.fM + j;
    constexpr int ncf = (M1 + KF - 1) / KF;
#pragma unroll
    for (int cc = 0; cc < ncf; ++cc) {
      const int ib = cc * KF + 1;
      double mm[KF], yy[KF];
#pragma unroll
      for (int k = 0; k < KF; ++k) {
        const int i = (ib + k <= M1) ? ib + k : M1;
        mm[k] = pm[(size_t)(i - 1) * PJ];
        yy[k] = y[i];
      }
#pragma unroll
      for (int k = 0; k < KF; ++k) {
        if (ib + k <= M1) {
          x = yy[k] - mm[k] * x;
          y[ib + k] = x;
        }
      }
    }
  }
#ifdef HADI_PHASE_TIMING
  if (dbg) dbg[0] += clock64() - dbg_t0;
#endif
  // ---- back substitution: x_i = (x_i - impl_upper(j,i) x_{i+1}) / pivot(j,i), i = m1..1
  // impl_upper(j,i) = -theta*dt*(a*delta_s(+1) + b*beta_s(+1)) from the zero-padded tables (exactly 0 at i = m1)
  const double* hs2 = w.ti + TI_HS2 * N1;
  const double* dsp = w.ti + TI_DSP * N1;
  const double* bbp = w.ti + TI_BBP * N1;
  constexpr int PB = 2;                                  // shared-memory lookahead of the chain (nodes)
  double xn = 0.0;
  unsigned badl = 0;
  hadi_tm_wait_ld(qa);
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    const int top = M1 - h * H;                          // first (largest) node of this half
    double* yh = ((g == h) ? y : dummy) + top;           // element e of the half <-> yh[-e]
    const double* hh = hs2 + top;
    const double* dh = dsp + top;
    const double* bh = bbp + top;
    if (h == 1) {
      xn = __shfl_sync(0xffffffffu, xn, l);              // lanes 16..31 take the row over from lane l
      hadi_tm_ld_quad(mine, qa);                         // their first quad (same columns, other lanes' data)
      hadi_tm_wait_ld(qa);
    }
    double yq[PB], hq[PB], dq[PB], bq[PB];
#pragma unroll
    for (int e = 0; e < PB; ++e) {
      yq[e] = yh[-e];
      hq[e] = hh[-e]; dq[e] = dh[-e]; bq[e] = bh[-e];
    }
    double iu = ntd * ((hq[0] * vj) * dq[0] + bq[0]);
#pragma unroll
    for (int e = 0; e < H; ++e) {
      const int f = e + PB;
      const int c = e / 4, k = e % 4;
      HadiTmQuad& cur = (c % 2 == 0) ? qa : qb;
      HadiTmQuad& nxt = (c % 2 == 0) ? qb : qa;
      if (k == 0 && c + 1 < NQ) hadi_tm_ld_quad(mine + 16u * (unsigned)(c + 1), nxt);   // a chunk ahead of the chain
      const double tc = hadi_tm_piv(cur, k), rc = hadi_tm_rcp(cur, k), yc = yq[e % PB];
      double iun = 0.0;
      if (e + 1 < H) iun = ntd * ((hq[(e + 1) % PB] * vj) * dq[(e + 1) % PB] + bq[(e + 1) % PB]);
      if (f < H) {
        yq[e % PB] = yh[-f];
        hq[e % PB] = hh[-f]; dq[e % PB] = dh[-f]; bq[e % PB] = bh[-f];
      }
      x = hadi_div<EXACT>(yc - iu * xn, tc, rc, badl);
      xn = x;
      yh[-e] = x;
      iu = iun;
      if ((k == 3 || e == H - 1) && c + 1 < NQ) hadi_tm_wait_ld(nxt);
    }
    if (g == h && own) bad |= badl;
    badl = 0;
  }
}

// ----------------------------------------------------------------------------------------------
// S1 entirely out of tensor memory (FEED 6, the "duo" kernel: two solves per CTA, one CTA per SM, all 512 columns).
//
// With 512 columns a lane holds 256 doubles: one lane per v-row carries the row's whole back-substitution stream
// — {pivot, prepared reciprocal} of node i at columns 4(m1-i) .. +3 — and the first HADI_TM_MT Thomas
// multipliers at columns 4 m1 + 2(i-1).  Two chain warps per solve (rows 0..25 and 26..50 at 101 x 51), as in the
// plain-load kernel, so the FP64 pipe sees no more instructions than before (an FP64 instruction costs the pipe
// the same whether 13 or 32 lanes are active; the pair layout of FEED 5 above doubles the chain warps and loses
// more in the other CTA's phases than S1 gains).  The multipliers that do not fit are re-formed on the way from
// the pair of the previous node: m_i = impl_lower(j,i) / pivot(j,i-1) with the prepared reciprocal — the bit
// pattern the factorisation stored (hadi_div) — 7 FP64 operations off the dependent chain.  Phase S1 then touches
// neither the L2 nor (beyond Y and three coefficient tables) shared memory.
// Team t of the CTA runs its chains on the two team-local warps whose CTA-wide warp index falls into TMEM quarters
// 2t and 2t + 1 (`wbase`: warps 0, 1 of team 0; warps 2, 3 of team 1 when a team has 8 warps, 0, 1 when it has 10).
#ifndef HADI_TM_MT
#define HADI_TM_MT 56
#endif
HADI_HD constexpr int hadi_tm2_rows(int m2) { return (m2 + 2) / 2; }   // v-rows per chain warp

template <int M1, int M2>
__device__ __forceinline__ void hadi_tmem2_factor(const HadiItem& it, const HadiView& w, const double* vg, int tid,
                                                  unsigned tmem, int wbase) {
  constexpr int NRW = hadi_tm2_rows(M2), MT = HADI_TM_MT;
  constexpr int N1 = hadi_geo_n1(M1);
  static_assert(NRW <= 32 && 4 * M1 + 2 * MT <= 512, "one TMEM lane per v-row: 4 columns per node + the stored multipliers");
  const int warp = (tid >> 5) - wbase, lane = tid & 31;   // chain warps: team-local warps wbase, wbase + 1
  if (warp < 0 || warp >= 2) return;
  const bool own = (lane < NRW) && (warp * NRW + lane <= M2);
  const int j = warp * NRW + (own ? lane : 0);
  const unsigned mine = tmem + ((unsigned)(32 * ((threadIdx.x >> 5) & 3)) << 16);
  const double vj = vg[j];
  const double* hs2 = w.ti + TI_HS2 * N1;
  const double* dsm = w.ti + TI_DSM * N1;
  const double* ds0 = w.ti + TI_DS0 * N1;
  const double* dsp = w.ti + TI_DSP * N1;
  const double* sv = w.ti + TI_S * N1;
  const double* bsm = w.ti + TI_BSM * N1;
  const double* bs0 = w.ti + TI_BS0 * N1;
  const double* bbp = w.ti + TI_BBP * N1;
  const double theta = it.theta, dt = it.dt;
  const double rdiff = it.r_d - it.r_f;
  double t = 1.0;           // impl_main(j,0)
  double iu_prev = 0.0;     // impl_upper(j,0)
#pragma unroll 1
  for (int i = 1; i <= M1; ++i) {
    double il, im, iu;
    if (i < M1) {
      const double a = hs2[i] * vj;
      const double b = rdiff * sv[i];
      const double lo = a * dsm[i] + b * bsm[i];
      const double ma = a * ds0[i] + b * bs0[i] - 0.5 * it.r_d;
      const double up = a * dsp[i] + bbp[i];
      il = -theta * dt * lo;
      im = 1.0 - theta * dt * ma;
      iu = -theta * dt * up;
    } else {
      const double ma = -0.5 * it.r_d;
      il = 0.0;
      im = 1.0 - theta * dt * ma;
      iu = 0.0;
    }
    const double m = il / t;
    t = im - m * iu_prev;
    iu_prev = iu;
    hadi_tm_st_pair(mine + 4u * (unsigned)(M1 - i), t, hadi_rcp_prep(t));
    if (i <= MT) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(mine + 4u * (unsigned)M1 + 2u * (unsigned)(i - 1)),
                   "r"(__double2loint(m)), "r"(__double2hiint(m))
                   : "memory");
    }
  }
  hadi_tm_wait_st();
}

// One quad (four nodes, elements 4q .. 4q+3 <-> i = M1-4q .. M1-4q-3) of the back substitution.  `cur` holds the
// quad's {pivot, reciprocal} pairs, yc / uc its right-hand sides and impl_upper values.  While the dependent chain
// of the four nodes runs, the operands of the NEXT quad are fetched (tensor memory -> nxt, y and the three
// coefficient tables -> registers) and its impl_upper values are formed stage by stage, four independent
// operations after each node of the chain, so that neither a load nor an off-chain FP64 latency lands on the chain.
template <int M1, bool EXACT, bool MORE>
__device__ __forceinline__ void hadi_tm2_bwd_quad(const HadiTmQuad& cur, HadiTmQuad& nxt, const double (&yc)[4],
                                                  const double (&uc)[4], double (&yn)[4], double (&un)[4],
                                                  double* yq, const double* hq, const double* dq, const double* bq,
                                                  unsigned next_cols, double vj, double ntd, double& xn,
                                                  unsigned& badl) {
  // yq, hq, dq, bq point at the first (largest-i) node of THIS quad: element k is [-k]; the next quad starts at [-4]
  double hn[4], dn[4], bn[4];
  if (MORE) {
    hadi_tm_ld_quad(next_cols, nxt);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      yn[k] = yq[-4 - k];
      hn[k] = hq[-4 - k]; dn[k] = dq[-4 - k]; bn[k] = bq[-4 - k];
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double x = hadi_div<EXACT>(yc[k] - uc[k] * xn, hadi_tm_piv(cur, k), hadi_tm_rcp(cur, k), badl);
    xn = x;
    yq[-k] = x;
    if (MORE) {
      // stage k of the next quad's impl_upper = ntd * ((hs2 * v) * dsp + bbp), all four nodes
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (k == 0) un[q] = hn[q] * vj;
        if (k == 1) un[q] = un[q] * dn[q];
        if (k == 2) un[q] = un[q] + bn[q];
        if (k == 3) un[q] = ntd * un[q];
      }
    }
  }
  if (MORE) hadi_tm_wait_ld(nxt);
}

// Forward counterpart for the nodes whose multiplier is re-formed: quad `cur` holds the pairs of elements
// 4c .. 4c+3, i.e. of nodes i-1 for i = M1-4c+1 .. M1-4c-2 (descending slots 3..0 serve ascending i).
// m_i = impl_lower(j,i) / pivot(j,i-1), impl_lower = ntd * ((hs2 v) dsm + (rdiff s) bsm): formed for the four
// nodes stage by stage (nine stages of four independent operations), then the chain x_i = y_i - m_i x_{i-1}.
template <int M1, bool EXACT>
__device__ __forceinline__ void hadi_tm2_fwd_quad(const HadiTmQuad& cur, int i0, int nn, int slot0, double* y,
                                                  const double* hs2, const double* dsm, const double* sv,
                                                  const double* bsm, double vj, double ntd, double rdiff, double& x,
                                                  unsigned& badl) {
  // nodes i0 .. i0+nn-1 (nn <= 4, compile-time after inlining); node i0+k uses slot slot0-k of `cur`
  double il[4], yy[4], mm[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < nn) {
      const int i = i0 + k;
      yy[k] = y[i];
      il[k] = hs2[i] * vj;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < nn) il[k] = il[k] * dsm[i0 + k];
  double bb[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < nn) bb[k] = rdiff * sv[i0 + k];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < nn) bb[k] = bb[k] * bsm[i0 + k];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < nn) il[k] = il[k] + bb[k];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < nn) il[k] = (i0 + k < M1) ? ntd * il[k] : 0.0;     // impl_lower(j, m1) = 0
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < nn) mm[k] = hadi_div<EXACT>(il[k], hadi_tm_piv(cur, slot0 - k), hadi_tm_rcp(cur, slot0 - k), badl);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < nn) {
      x = yy[k] - mm[k] * x;
      y[i0 + k] = x;
    }
  }
}

template <int M1, int M2, bool EXACT>
__device__ __forceinline__ void hadi_tmem2_solve_a1(const HadiItem& it, const HadiView& w, int tid, unsigned& bad,
                                                    unsigned tmem, int wbase, long long* dbg = nullptr) {
  constexpr int NRW = hadi_tm2_rows(M2), MT = HADI_TM_MT;
  constexpr int LD = hadi_geo_ld(M1), N1 = hadi_geo_n1(M1), N2 = hadi_geo_n2(M2);
  static_assert(MT % 8 == 0 && MT >= 8 && MT < M1 && M1 % 4 == 0 && (M1 - MT) % 4 == 0,
                "stored multipliers come in x16 loads of eight, pairs in quads of four");
  const int warp = (tid >> 5) - wbase, lane = tid & 31;
  if (warp < 0 || warp >= 2) return;
  const bool own = (lane < NRW) && (warp * NRW + lane <= M2);
  const int j = warp * NRW + (own ? lane : 0);
  const unsigned mine = tmem + ((unsigned)(32 * ((threadIdx.x >> 5) & 3)) << 16);
  double* y = w.Y + j * LD;
  const double vj = w.tj[TJ_V * N2 + j];
  const double ntd = -it.theta * it.dt;
  const double* hs2 = w.ti + TI_HS2 * N1;
  const double* dsm = w.ti + TI_DSM * N1;
  const double* dsp = w.ti + TI_DSP * N1;
  const double* sv = w.ti + TI_S * N1;
  const double* bsm = w.ti + TI_BSM * N1;
  const double* bbp = w.ti + TI_BBP * N1;
  const double rdiff = it.r_d - it.r_f;
  unsigned badl = 0;
#ifdef HADI_PHASE_TIMING
  const long long dbg_t0 = clock64();
#endif
  HadiTmQuad qa, qb;
  // ---- forward elimination: x_i = y_i - m_i x_{i-1}, i = 1..m1.  The loops are rolled (a few hundred instructions
  // in all): with tensor memory there is no load to hoist a long way ahead, and the fully unrolled sweeps of the
  // plain-load kernels are what overflows the instruction cache once two teams run different phases.
  double x = y[0];
  {
    // nodes 1..MT: stored multipliers, eight per load, one load ahead of the chain
    constexpr int NQ = MT / 8;
    hadi_tm_ld_quad(mine + 4u * (unsigned)M1, qa);
    double yc[8], yn[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) yc[k] = y[1 + k];
    hadi_tm_wait_ld(qa);
#pragma unroll 1
    for (int c = 0; c < NQ; ++c) {
      // next load: the following eight multipliers, or (last round) the first pair quad of the re-formed part
      const unsigned ncol = (c + 1 < NQ) ? 4u * (unsigned)M1 + 16u * (unsigned)(c + 1) : 16u * (unsigned)((M1 - MT) / 4);
      hadi_tm_ld_quad(mine + ncol, qb);
      double* yb = y + 8 * c + 1;
      if (c + 1 < NQ) {
#pragma unroll
        for (int k = 0; k < 8; ++k) yn[k] = yb[8 + k];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double mc = __hiloint2double((int)qa.r[2 * k + 1], (int)qa.r[2 * k]);
        x = yc[k] - mc * x;
        yb[k] = x;
      }
      hadi_tm_wait_ld(qb);
      qa = qb;
#pragma unroll
      for (int k = 0; k < 8; ++k) yc[k] = yn[k];
    }
    // nodes MT+1..m1, multipliers re-formed from the pair of node i-1 = element M1-i+1: node MT+1 takes slot 0 of
    // quad C0 = (M1-MT)/4 (in qa now), then quads C0-1 .. 1 serve four nodes each (slots 3..0), quad 0 the last three
    constexpr int C0 = (M1 - MT) / 4;
    hadi_tm_ld_quad(mine + 16u * (unsigned)(C0 - 1), qb);
    hadi_tm2_fwd_quad<M1, EXACT>(qa, MT + 1, 1, 0, y, hs2, dsm, sv, bsm, vj, ntd, rdiff, x, badl);
    hadi_tm_wait_ld(qb);
#pragma unroll 1
    for (int c = C0 - 1; c >= 1; --c) {
      qa = qb;
      hadi_tm_ld_quad(mine + 16u * (unsigned)(c - 1), qb);
      hadi_tm2_fwd_quad<M1, EXACT>(qa, M1 - 4 * c - 2, 4, 3, y, hs2, dsm, sv, bsm, vj, ntd, rdiff, x, badl);
      hadi_tm_wait_ld(qb);
    }
    hadi_tm2_fwd_quad<M1, EXACT>(qb, M1 - 2, 3, 3, y, hs2, dsm, sv, bsm, vj, ntd, rdiff, x, badl);
  }
#ifdef HADI_PHASE_TIMING
  if (dbg) dbg[0] += clock64() - dbg_t0;
#endif
  // ---- back substitution: x_i = (x_i - impl_upper(j,i) x_{i+1}) / pivot(j,i), i = m1..1 (element e <-> i = m1 - e)
  // impl_upper(j,i) = -theta*dt*(a*delta_s(+1) + b*beta_s(+1)) from the zero-padded tables (exactly 0 at i = m1)
  {
    constexpr int NQ = M1 / 4;
    hadi_tm_ld_quad(mine, qa);
    double ya[4], ua[4], yb[4], ub[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ya[k] = y[M1 - k];
      ua[k] = ntd * ((hs2[M1 - k] * vj) * dsp[M1 - k] + bbp[M1 - k]);
    }
    double xn = 0.0;
    hadi_tm_wait_ld(qa);
    int q = 0;
#pragma unroll 1
    for (; q + 2 < NQ; q += 2) {
      const int i0 = M1 - 4 * q;
      hadi_tm2_bwd_quad<M1, EXACT, true>(qa, qb, ya, ua, yb, ub, y + i0, hs2 + i0, dsp + i0, bbp + i0,
                                         mine + 16u * (unsigned)(q + 1), vj, ntd, xn, badl);
      hadi_tm2_bwd_quad<M1, EXACT, true>(qb, qa, yb, ub, ya, ua, y + i0 - 4, hs2 + i0 - 4, dsp + i0 - 4, bbp + i0 - 4,
                                         mine + 16u * (unsigned)(q + 2), vj, ntd, xn, badl);
    }
    // one or two quads are left
    {
      const int i0 = M1 - 4 * q;
      if (q + 2 == NQ) {
        hadi_tm2_bwd_quad<M1, EXACT, true>(qa, qb, ya, ua, yb, ub, y + i0, hs2 + i0, dsp + i0, bbp + i0,
                                           mine + 16u * (unsigned)(q + 1), vj, ntd, xn, badl);
        hadi_tm2_bwd_quad<M1, EXACT, false>(qb, qa, yb, ub, ya, ua, y + i0 - 4, hs2 + i0 - 4, dsp + i0 - 4, bbp + i0 - 4,
                                            0u, vj, ntd, xn, badl);
      } else {
        hadi_tm2_bwd_quad<M1, EXACT, false>(qa, qb, ya, ua, yb, ub, y + i0, hs2 + i0, dsp + i0, bbp + i0, 0u, vj, ntd,
                                            xn, badl);
      }
    }
  }
  if (own) bad |= badl;
}

// ----------------------------------------------------------------------------------------------
// S1 with the back-substitution stream in tensor memory, two CTAs per SM (FEED 7, "relay").
//
// 256 columns per CTA hold 64 {pivot, reciprocal} pairs per lane, and a v-row needs 100.  The sweep is therefore run
// as a relay: warps 0 and 1 (rows 0..25 / 26..50, one lane per row, as the plain-load kernel assigns them) carry
// the chain through elements 0..63 out of their own lanes, then warps 2 and 3 — idle in this phase otherwise —
// take the rows over for elements 64..99, whose pairs sit in THEIR quarters of tensor memory.  The hand-over costs
// one named barrier per sweep (x of the last node is already where the next node reads it: in Y); every node is
// still processed by exactly one warp, so the FP64 pipe sees the instruction count of the plain-load kernel while
// no back-substitution operand comes from L2 any more.  The forward sweep keeps its multipliers in L2 (fM).
#define HADI_RELAY_SPLIT 64       /* elements served by warps 0, 1; a multiple of 8 */
template <int M1, int M2>
__device__ __forceinline__ void hadi_relay_factor(const HadiItem& it, const HadiView& w, const double* vg, int tid,
                                                  unsigned tmem) {
  constexpr int NRW = hadi_tm2_rows(M2), SP = HADI_RELAY_SPLIT;
  constexpr int N1 = hadi_geo_n1(M1), PJ = hadi_geo_pj(M2);
  static_assert(NRW <= 32 && 4 * SP <= 256 && 4 * (M1 - SP) <= 256 && M1 > SP, "64 pairs per lane and 256 columns");
  const int warp = tid >> 5, lane = tid & 31;
  if (warp >= 4) return;
  // all four warps run the recurrence of their rows (warps 2, 3 repeat rows of warps 0, 1: they are idle in this
  // phase, and each warp can only store into its own quarter of tensor memory)
  const int grp = warp & 1, late = warp >> 1;
  const bool own = (lane < NRW) && (grp * NRW + lane <= M2);
  const int j = grp * NRW + (own ? lane : 0);
  const unsigned mine = tmem + ((unsigned)(32 * warp) << 16);
  const double vj = vg[j];
  const double* hs2 = w.ti + TI_HS2 * N1;
  const double* dsm = w.ti + TI_DSM * N1;
  const double* ds0 = w.ti + TI_DS0 * N1;
  const double* dsp = w.ti + TI_DSP * N1;
  const double* sv = w.ti + TI_S * N1;
  const double* bsm = w.ti + TI_BSM * N1;
  const double* bs0 = w.ti + TI_BS0 * N1;
  const double* bbp = w.ti + TI_BBP * N1;
  const double theta = it.theta, dt = it.dt;
  const double rdiff = it.r_d - it.r_f;
  double t = 1.0;           // impl_main(j,0)
  double iu_prev = 0.0;     // impl_upper(j,0)
#pragma unroll 1
  for (int i = 1; i <= M1; ++i) {
    double il, im, iu;
    if (i < M1) {
      const double a = hs2[i] * vj;
      const double b = rdiff * sv[i];
      const double lo = a * dsm[i] + b * bsm[i];
      const double ma = a * ds0[i] + b * bs0[i] - 0.5 * it.r_d;
      const double up = a * dsp[i] + bbp[i];
      il = -theta * dt * lo;
      im = 1.0 - theta * dt * ma;
      iu = -theta * dt * up;
    } else {
      const double ma = -0.5 * it.r_d;
      il = 0.0;
      im = 1.0 - theta * dt * ma;
      iu = 0.0;
    }
    const double m = il / t;
    t = im - m * iu_prev;
    iu_prev = iu;
    const int e = M1 - i;                                   // element of the back substitution
    if (late == 0) w.fM[(size_t)(i - 1) * PJ + j] = m;
    if ((e >= SP) == (late == 1)) hadi_tm_st_pair(mine + 4u * (unsigned)(late ? e - SP : e), t, hadi_rcp_prep(t));
  }
  hadi_tm_wait_st();
}

// quads [0, NQ) of a stage whose first element is E0 (node M1 - E0), pairs at columns 0.. of this warp's lanes
template <int M1, bool EXACT, int E0, int NQ>
__device__ __forceinline__ void hadi_relay_stage(unsigned mine, double* y, const double* hs2, const double* dsp,
                                                 const double* bbp, double vj, double ntd, double& xn, unsigned& badl) {
  HadiTmQuad qa, qb;
  hadi_tm_ld_quad(mine, qa);
  constexpr int I0 = M1 - E0;
  double ya[4], ua[4], yb[4], ub[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    ya[k] = y[I0 - k];
    ua[k] = ntd * ((hs2[I0 - k] * vj) * dsp[I0 - k] + bbp[I0 - k]);
  }
  hadi_tm_wait_ld(qa);
  int q = 0;
#pragma unroll 1
  for (; q + 2 < NQ; q += 2) {
    const int i0 = I0 - 4 * q;
    hadi_tm2_bwd_quad<M1, EXACT, true>(qa, qb, ya, ua, yb, ub, y + i0, hs2 + i0, dsp + i0, bbp + i0,
                                       mine + 16u * (unsigned)(q + 1), vj, ntd, xn, badl);
    hadi_tm2_bwd_quad<M1, EXACT, true>(qb, qa, yb, ub, ya, ua, y + i0 - 4, hs2 + i0 - 4, dsp + i0 - 4, bbp + i0 - 4,
                                       mine + 16u * (unsigned)(q + 2), vj, ntd, xn, badl);
  }
  const int i0 = I0 - 4 * q;
  if (NQ % 2 == 0) {
    hadi_tm2_bwd_quad<M1, EXACT, true>(qa, qb, ya, ua, yb, ub, y + i0, hs2 + i0, dsp + i0, bbp + i0,
                                       mine + 16u * (unsigned)(q + 1), vj, ntd, xn, badl);
    hadi_tm2_bwd_quad<M1, EXACT, false>(qb, qa, yb, ub, ya, ua, y + i0 - 4, hs2 + i0 - 4, dsp + i0 - 4, bbp + i0 - 4, 0u,
                                        vj, ntd, xn, badl);
  } else {
    hadi_tm2_bwd_quad<M1, EXACT, false>(qa, qb, ya, ua, yb, ub, y + i0, hs2 + i0, dsp + i0, bbp + i0, 0u, vj, ntd, xn,
                                        badl);
  }
}

template <int M1, int M2, bool EXACT>
__device__ __forceinline__ void hadi_relay_solve_a1(const HadiItem& it, const HadiView& w, int tid, unsigned& bad,
                                                    unsigned tmem, long long* dbg = nullptr) {
  constexpr int NRW = hadi_tm2_rows(M2), SP = HADI_RELAY_SPLIT;
  constexpr int LD = hadi_geo_ld(M1), N1 = hadi_geo_n1(M1), N2 = hadi_geo_n2(M2), PJ = hadi_geo_pj(M2);
  constexpr int KF = HADI_KF;
  static_assert(M1 % 4 == 0 && SP % 4 == 0, "quads of four nodes");
  const int warp = tid >> 5, lane = tid & 31;
  if (warp >= 4) return;
  const int grp = warp & 1, late = warp >> 1;
  const bool own = (lane < NRW) && (grp * NRW + lane <= M2);
  const int j = grp * NRW + (own ? lane : 0);
  const unsigned mine = tmem + ((unsigned)(32 * warp) << 16);
  double* y = w.Y + j * LD;
  const double vj = w.tj[TJ_V * N2 + j];
  const double ntd = -it.theta * it.dt;
  const double* hs2 = w.ti + TI_HS2 * N1;
  const double* dsp = w.ti + TI_DSP * N1;
  const double* bbp = w.ti + TI_BBP * N1;
  unsigned badl = 0;
  if (late == 0) {
#ifdef HADI_PHASE_TIMING
    const long long dbg_t0 = clock64();
#endif
    // ---- forward elimination: x_i = y_i - m_i x_{i-1}, i = 1..m1, multipliers from L2 in chunks
    double x = y[0];
    {
      const double* pm = w.fM + j;
      constexpr int ncf = (M1 + KF - 1) / KF;
#pragma unroll
      for (int cc = 0; cc < ncf; ++cc) {
        const int ib = cc * KF + 1;
        double mm[KF], yy[KF];
#pragma unroll
        for (int k = 0; k < KF; ++k) {
          const int i = (ib + k <= M1) ? ib + k : M1;
          mm[k] = pm[(size_t)(i - 1) * PJ];
          yy[k] = y[i];
        }
#pragma unroll
        for (int k = 0; k < KF; ++k) {
          if (ib + k <= M1) {
            x = yy[k] - mm[k] * x;
            y[ib + k] = x;
          }
        }
      }
    }
#ifdef HADI_PHASE_TIMING
    if (dbg) dbg[0] += clock64() - dbg_t0;
#endif
    // ---- back substitution, elements 0 .. SP-1 (nodes m1 .. m1-SP+1)
    double xn = 0.0;
    hadi_relay_stage<M1, EXACT, 0, SP / 4>(mine, y, hs2, dsp, bbp, vj, ntd, xn, badl);
    // hand the rows to warps 2, 3: x of node m1-SP+1 is in Y
    asm volatile("bar.arrive 1, 128;" ::: "memory");
  } else {
    asm volatile("bar.sync 1, 128;" ::: "memory");
    // ---- elements SP .. m1-1 (nodes m1-SP .. 1) out of this warp's own quarter of tensor memory
    double xn = y[M1 - SP + 1];
    hadi_relay_stage<M1, EXACT, SP, (M1 - SP) / 4>(mine, y, hs2, dsp, bbp, vj, ntd, xn, badl);
  }
  if (own) bad |= badl;
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
