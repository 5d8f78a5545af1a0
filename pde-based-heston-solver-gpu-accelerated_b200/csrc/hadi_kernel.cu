// hadi — fused, persistent Douglas ADI kernel for sm_100a (B200).
//
// One CTA solves one work item (an option or one Jacobian bump) from payoff to price without
// leaving the SM: the solution U and the stage vector Y stay in shared memory for all N time steps;
// the explicit A0/A1/A2 products, both implicit line solves, the boundary terms, the dividend jump
// and the American projection are fused into three barrier-separated phases per step.  CTAs are
// persistent and pull items from a global counter (longest items first), so a batch of any size
// runs as one launch; two CTAs share an SM so that one option's latency-bound line solves overlap
// the other's throughput-bound explicit stage.
//
// Replaces the reference's "Base_Price_computation" / "Jacobian_computation" Kokkos kernels
// (src/jacobian_computation.cpp:232,391,...; src/heston_calibration.cpp:2206,2366) together with
// the 12 global-memory work arrays of DO_Workspace (src/DO_solver_workspace.hpp:5-44).
//
// Compile with -fmad=false: parity with the reference requires un-fused multiplies and adds.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "hadi_launch.h"
#include "hadi_phases_cs.cuh"

namespace {

#ifdef HADI_PHASE_TIMING
#define HADI_TICK(k) { const long long tn = clock64(); tacc[k] += tn - tlast; tlast = tn; }
#else
#define HADI_TICK(k)
#endif
// development aid: end the item after phase k of step L.dbg_step (state is then dumped by the caller)
#if defined(HADI_DEBUG_TRACE)
// development aid: per-step, per-phase position-weighted checksums of U, Y and lambda into out_U
__device__ __forceinline__ void hadi_dbg_hash(const HadiLaunch& L, const HadiItem& it, const HadiView& w, int n, int k,
                                              int tid, int nt) {
  __shared__ unsigned long long hs[3];
  if (L.out_U == nullptr) return;
  if (tid < 3) hs[tid] = 0ULL;
  __syncthreads();
  unsigned long long a = 0, b = 0, c = 0;
  for (int p = tid; p < (w.m2 + 1) * (w.m1 + 1); p += nt) {
    const int j = p / (w.m1 + 1), i = p - j * (w.m1 + 1);
    const unsigned long long wt = 2ULL * (unsigned long long)p + 1ULL;
    a += (unsigned long long)__double_as_longlong(w.U[j * w.ld + i]) * wt;
    b += (unsigned long long)__double_as_longlong(w.Y[j * w.ld + i]) * wt;
    if (it.style == 1) c += (unsigned long long)__double_as_longlong(hadi_lam_ld(&w.lam[j * w.ld + i])) * wt;
  }
  atomicAdd(&hs[0], a); atomicAdd(&hs[1], b); atomicAdd(&hs[2], c);
  __syncthreads();
  if (tid < 3) {
    unsigned long long* tr = reinterpret_cast<unsigned long long*>(L.out_U + (size_t)it.out * w.P + (size_t)(24 * n + 3 * k));
    tr[tid] = hs[tid];
  }
  __syncthreads();
}
#define HADI_STOP(k) hadi_dbg_hash(L, it, w, n, k, tid, NT);
#elif defined(HADI_DEBUG_STOP)
#define HADI_STOP(k) if (L.dbg_step == n && (L.dbg_phase % 100) == (k)) { stopped = true; break; }
#else
#define HADI_STOP(k)
#endif

#ifdef HADI_DEBUG_LAM
// development aid: before phase E the multiplier must be identical in Y (shared) and in the global
// scratch copy; record disagreements in L.prof[blockIdx*8 ..]: count, step, i, j, bits(global),
// bits(shared), smid, item
__device__ __forceinline__ void hadi_dbg_check_lam(const HadiLaunch& L, const HadiView& w, int n, int item, int tid, int nt) {
  const HadiMap mp = hadi_map(w.m1, w.m2, tid, nt);
  if (!mp.active) return;
  for (int j = mp.j0; j < mp.j1; ++j) {
    const double lg = hadi_lam_ld(&w.lam[j * w.ld + mp.i]);
    const double ls = w.Y[j * w.ld + mp.i];
    if (__double_as_longlong(lg) != __double_as_longlong(ls)) {
      long long* rec = L.prof + (size_t)blockIdx.x * 8;
      if (atomicAdd((unsigned long long*)&rec[0], 1ULL) == 0ULL) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        rec[1] = n; rec[2] = mp.i; rec[3] = j;
        rec[4] = __double_as_longlong(lg); rec[5] = __double_as_longlong(ls);
        rec[6] = smid; rec[7] = item;
      }
    }
  }
}
#endif

// Item result: the price at (S0, V0) (src/jacobian_computation.cpp:275-287,330).  Batches that take the V0
// column of the Jacobian by interpolation (src/device_solver.cpp:1725-1829) publish three values per item:
// the price and U at the two v-rows bracketing V0 + eps on the S0 column (HadiItem::aux packs them).
__device__ __forceinline__ void hadi_publish(const HadiLaunch& L, const HadiItem& it, const HadiView& w) {
  if (L.out_stride <= 1) {
    L.out_values[it.out] = w.U[it.idx_v * w.ld + it.idx_s];
  } else {
    double* o = L.out_values + (size_t)it.out * L.out_stride;
    o[0] = w.U[it.idx_v * w.ld + it.idx_s];
    o[1] = w.U[(it.aux & 0xffff) * w.ld + it.idx_s];
    o[2] = w.U[((it.aux >> 16) & 0xffff) * w.ld + it.idx_s];
  }
}


// CTA-wide or sub-CTA barrier.  The "duo" variant runs TWO solves per CTA (640 threads, one SM, all 512 columns of
// tensor memory): threads [0, NT) and [NT, 2 NT) are two independent teams with their own shared-memory working
// set, item loop and named barrier (bar.sync 1 + team, NT).  id 0 / all threads is __syncthreads().
struct HadiBar {
  unsigned id, nt;
  __device__ __forceinline__ void sync() const {
    if (id == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nt) : "memory");
  }
  __device__ __forceinline__ bool sync_or(int pred) const {
    if (id == 0) return __syncthreads_or(pred) != 0;
    unsigned r;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.s32 p, %3, 0;\n"
        "bar.red.or.pred q, %1, %2, p;\n"
        "selp.u32 %0, 1, 0, q;\n"
        "}\n"
        : "=r"(r)
        : "r"(id), "r"(nt), "r"(pred)
        : "memory");
    return r != 0;
  }
};
// Turn-taking between the two teams of a duo CTA.  Both teams run the same phase sequence on equal work: left
// alone they fall into lockstep (explicit stage against explicit stage on the FP64 pipe, line solves against line
// solves with the pipe idle), which is the worst pairing.  A team takes this lock for its FP64-bound stages, so that
// one team's explicit stage runs beside the other's latency-bound line solves.
struct HadiTurn {
  int* lock;      // shared memory, nullptr: no turn-taking
  __device__ __forceinline__ void acquire(const HadiBar& bar, int tid) const {
    if (lock == nullptr) return;
    if (tid == 0) {
      while (atomicCAS(lock, 0, 1) != 0) __nanosleep(64);
    }
    bar.sync();
  }
  // call after the team barrier that ends the stage
  __device__ __forceinline__ void release(int tid) const {
    if (lock != nullptr && tid == 0) atomicExch(lock, 0);
  }
};
#define HADI_SYNC() bar.sync()

// One item, payoff to price.  Returns (CTA-uniform) whether any guarded division left its fast-path
// range; EXACT = true compiles every division as IEEE '/'.
// FEED == 4: co-operative S1 / pipelined S2 of hadi_phases_fast.cuh (grid-specialised variants only)
struct HadiCoopFeed : HadiDirectFeed {
  static constexpr bool kCoop = true;
};
template <class F> struct hadi_is_coop { static constexpr bool value = false; };
template <> struct hadi_is_coop<HadiCoopFeed> { static constexpr bool value = true; };
// FEED == 5: back-substitution factors of S1 in tensor memory (hadi_phases_fast.cuh)
struct HadiTmemFeed : HadiDirectFeed {
  unsigned tmem = 0;        // base address of the CTA's tensor-memory allocation
  double* dummy = nullptr;  // a row of shared memory nobody reads (lanes that are not live in a half sweep)
};
template <class F> struct hadi_is_tmem { static constexpr bool value = false; };
template <> struct hadi_is_tmem<HadiTmemFeed> { static constexpr bool value = true; };
// FEED == 6: the whole S1 factor stream in tensor memory, one lane per v-row (duo kernel: 512 columns per CTA)
struct HadiTmem2Feed : HadiDirectFeed {
  unsigned tmem = 0;
  int wbase = 0;   // first of the team's two chain warps (team-local index)
};
// FEED == 7: back-substitution stream in tensor memory, relayed from warps 0, 1 to warps 2, 3 (two CTAs per SM)
struct HadiRelayFeed : HadiDirectFeed {
  unsigned tmem = 0;
};
template <class F> struct hadi_is_relay { static constexpr bool value = false; };
template <> struct hadi_is_relay<HadiRelayFeed> { static constexpr bool value = true; };
template <class F> struct hadi_is_tmem2 { static constexpr bool value = false; };
template <> struct hadi_is_tmem2<HadiTmem2Feed> { static constexpr bool value = true; };

// n0..n1: the time steps to run (1..N for a whole solve); hin: state to resume from (nullptr: start from the payoff)
template <int NT, int M1, int M2, bool EXACT, class Feed>
__device__ __forceinline__ bool hadi_solve_item(const HadiLaunch& L, const HadiItem& it, HadiView& w, Feed& feed,
                                                int tid, long long* tacc, long long& tlast, int n0, int n1,
                                                const double* hin, const HadiBar& bar, const HadiTurn& turn = HadiTurn{nullptr}) {
  const int m1 = M1 ? M1 : L.m1, m2 = M2 ? M2 : L.m2;
  const double* sg = L.s_pool + it.s_off;
  const double* vg = L.v_pool + it.v_off;
  const double* eg = L.e_pool + it.e_off;
  w.c = it.theta * it.dt;
  const double rdt = hadi_rcp_prep(it.dt);
  unsigned bad = 0;

  hadi_phase_tables(it, w, sg, vg, tid, NT);
  if constexpr (hadi_is_coop<Feed>::value && M1 > 0) {
    // staged / consumed block counters of the co-operative S1 (hadi_phases_fast.cuh)
    if (tid < 2 * hadi_co_warps(M2))
      reinterpret_cast<int*>(w.stg + hadi_co_warps(M2) * hadi_co_warp_doubles())[tid] = 0;
  }
  HADI_SYNC();
  if constexpr (hadi_is_tmem<Feed>::value && M1 > 0) {
    hadi_phase_factor(it, w, vg, tid, NT, NT - 1, false);   // A2 only; the A1 rows go to tensor memory
    hadi_tmem_factor<M1, M2>(it, w, vg, tid, feed.tmem);
  } else if constexpr (hadi_is_tmem2<Feed>::value && M1 > 0) {
    hadi_phase_factor(it, w, vg, tid, NT, NT - 1, false);
    hadi_tmem2_factor<M1, M2>(it, w, vg, tid, feed.tmem, feed.wbase);
  } else if constexpr (hadi_is_relay<Feed>::value && M1 > 0) {
    hadi_phase_factor(it, w, vg, tid, NT, NT - 1, false);
    hadi_relay_factor<M1, M2>(it, w, vg, tid, feed.tmem);
  } else {
    hadi_phase_factor(it, w, vg, tid, NT, NT - 1);
  }
  if constexpr (Feed::kTma) {
    // the factor streams were written with generic stores and will be read by TMA (async proxy)
#ifndef HADI_NO_THREADFENCE
    __threadfence();
#endif
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  HADI_SYNC();  // the A2 assembly keeps scratch tables in Y: finish it before Y is initialised
  // initial condition U = payoff (the reference's U_0 input array), lambda = 0 — or the state another CTA
  // left after step n0 - 1 (split schedule)
  {
    const HadiMap mp = hadi_map(m1, m2, tid, NT);
    if (mp.active) {
      const double pay = hadi_ti(w, TI_PAY)[mp.i];
      for (int j = mp.j0; j < mp.j1; ++j) {
        if (hin == nullptr) {
          w.U[j * w.ld + mp.i] = pay;
          if (it.style == 1) {
            hadi_lam_st(&w.lam[j * w.ld + mp.i], 0.0);
            w.Y[j * w.ld + mp.i] = 0.0;   // phase E reads lambda from Y
          }
        } else {
          const size_t p = (size_t)j * (m1 + 1) + mp.i;
          w.U[j * w.ld + mp.i] = __ldcg(hin + p);
          if (it.style == 1) {
            const double lm = __ldcg(hin + (size_t)w.P + p);
            hadi_lam_st(&w.lam[j * w.ld + mp.i], lm);
            w.Y[j * w.ld + mp.i] = lm;
          }
        }
      }
    }
  }
  if constexpr (hadi_is_coop<Feed>::value && M1 > 0) hadi_fast_prestage<M1, M2>(w, tid, 0);
  HADI_SYNC();
  if constexpr (Feed::kTma) {
    if (feed.producer(tid)) feed.produce(0, it.N, 0);  // first chunks are in flight before step 1
  }
  if (tid <= m2) feed.begin_item(it.N, tid);
  HADI_TICK(0)

  int div_cur = 0;
  bool stopped = false;
  if (it.nd > 0)   // dividend queue position after the steps another CTA ran
    for (int n = 1; n < n0; ++n) {
      if (it.div_all) { while (hadi_dividend_next(n, it.dt, it.nd, L.div_dates, div_cur) >= 0) {} }
      else hadi_dividend_at(n, it.dt, it.nd, L.div_dates, div_cur);
    }
  for (int n = n0; n <= n1; ++n) {
    if (it.nd > 0) {
      // device schedule: one dividend per step at most; extension: every dividend dated inside the step, in order
      for (;;) {
        const int hit = it.div_all ? hadi_dividend_next(n, it.dt, it.nd, L.div_dates, div_cur)
                                   : hadi_dividend_at(n, it.dt, it.nd, L.div_dates, div_cur);
        if (hit < 0) break;  // uniform across the CTA
        hadi_phase_div1(w, L.div_amounts[hit], L.div_pcts[hit], tid, NT);
        HADI_SYNC();
        HADI_STOP(1)
        hadi_phase_div2(w, tid, NT);
        HADI_SYNC();
        HADI_STOP(2)
        if (it.style == 1) {
          hadi_phase_div3(w, tid, NT);
          HADI_SYNC();
        }
        HADI_STOP(3)
        if (!it.div_all) break;
      }
#ifdef HADI_DEBUG_STOP
      if (stopped) break;
#endif
    }
    HADI_TICK(0)
#ifdef HADI_DEBUG_LAM
    if (it.style == 1) hadi_dbg_check_lam(L, w, n, it.out, tid, NT);
#endif
    const double e0 = eg[n - 1], e1 = eg[n];
    turn.acquire(bar, tid);
    hadi_phase_explicit<M1, M2>(it, w, e0, e1, tid, NT);
    HADI_SYNC();
    turn.release(tid);
    HADI_TICK(2)
    HADI_STOP(4)
    if constexpr (hadi_is_coop<Feed>::value && M1 > 0) {
      hadi_fast_solve_a1<M1, M2, EXACT>(it, w, n - 1, tid, bad, &tacc[1]);
    } else if constexpr (hadi_is_tmem<Feed>::value && M1 > 0) {
      hadi_tmem_solve_a1<M1, M2, EXACT>(it, w, tid, bad, feed.tmem, feed.dummy, &tacc[1]);
    } else if constexpr (hadi_is_tmem2<Feed>::value && M1 > 0) {
      hadi_tmem2_solve_a1<M1, M2, EXACT>(it, w, tid, bad, feed.tmem, feed.wbase, &tacc[1]);
    } else if constexpr (hadi_is_relay<Feed>::value && M1 > 0) {
      hadi_relay_solve_a1<M1, M2, EXACT>(it, w, tid, bad, feed.tmem, &tacc[1]);
    } else {
      hadi_phase_solve_a1<M1, M2, EXACT>(it, w, e0, e1, n, tid, NT, feed, bad, &tacc[1]);
    }
    HADI_SYNC();
    HADI_TICK(3)
    HADI_STOP(5)
#ifdef HADI_SPLIT_R
    hadi_phase_rhs2<M1, M2>(it, w, e0, e1, tid, NT);
    HADI_SYNC();
#else
    if constexpr (M1 == 0) {
      if (!w.gstate) {   // uniform: the global-state kernels fold phase R into the column solve below
        hadi_phase_rhs2<M1, M2>(it, w, e0, e1, tid, NT);
        HADI_SYNC();
      }
    }
#endif
    HADI_TICK(7)
    HADI_STOP(6)
    if constexpr (M1 > 0) {
      // grid-specialised variants: phase R is folded into the forward sweep of the column solve
      hadi_fast_solve_a2<M1, M2, EXACT>(it, w, e0, e1, tid, bad);
      if constexpr (hadi_is_coop<Feed>::value) {
        if (n < it.N) hadi_fast_prestage<M1, M2>(w, tid, n);   // the feeder warps are idle in this phase
      }
    } else {
      // global-state kernels: phase R rides on the operand fetch of the forward sweep (one pass over Y and a barrier less)
      hadi_phase_solve_a2<M1, M2, EXACT>(it, w, tid, NT, bad, HadiRhs2{w.gstate ? 1 : 0, nullptr, e0, e1});
    }
    if (it.bc && tid * w.line_mul + w.line_off == 0) hadi_dirichlet_col0(w, it.K * eg[it.N + 1 + n]);
    HADI_SYNC();
    HADI_TICK(4)
    HADI_STOP(7)
    if (it.style == 1) {
      // all multiplier loads of a thread's rows are issued before the first is used (one L2 round trip)
      constexpr int CHP = (M1 == 100 && M2 == 50 && NT / 101 == 3) ? 17 : (M1 == 100 && M2 == 50 && NT / 101 == 2) ? 13 : HADI_CHP;
      if (L.flags & 2) turn.acquire(bar, tid);
      hadi_phase_project<M1, M2, EXACT, CHP>(it, w, rdt, tid, NT, bad);
      HADI_SYNC();
      if (L.flags & 2) turn.release(tid);
    }
    HADI_TICK(5)
    HADI_STOP(8)
  }
  if constexpr (Feed::kTma) {
#ifdef HADI_DEBUG_STOP
    if (stopped) {
      // drain the chunks the producer had in flight so that the ring counters stay consistent
      __shared__ unsigned s_issued;
      if (feed.producer(tid)) s_issued = feed.issued;
      HADI_SYNC();
      const unsigned upto = s_issued;
      if (tid <= m2) {
        while (feed.consumed != upto) {
          feed.probe = 0;
          feed.wait_slot();
          hadi_mbar_arrive(&feed.empty[feed.consumed % HADI_NS]);
          feed.consumed++;
        }
      }
      feed.issued = feed.consumed = feed.base = upto;
      HADI_SYNC();
    } else
#endif
    {
      // every chunk issued for this item has been consumed; keep both counters in step on all threads
      const unsigned nc = (unsigned)(feed.ncf() + feed.ncb());
      feed.issued = feed.consumed = feed.base = feed.base + (unsigned)it.N * nc;
    }
  }
  (void)stopped;
  return bar.sync_or((int)bad);
}

// Craig-Sneyd (European, no dividends) on the global-state working set: src/solver.hpp:781-907.
template <int NT, bool EXACT, class Feed>
__device__ __forceinline__ bool hadi_solve_item_cs(const HadiLaunch& L, const HadiItem& it, HadiView& w,
                                                   const HadiCsView& cs, Feed& feed, int tid, const HadiBar& bar) {
  const int m1 = L.m1, m2 = L.m2;
  const double* sg = L.s_pool + it.s_off;
  const double* vg = L.v_pool + it.v_off;
  const double* eg = L.e_pool + it.e_off;
  w.c = it.theta * it.dt;
  unsigned bad = 0;
  hadi_phase_tables(it, w, sg, vg, tid, NT);
  HADI_SYNC();
  hadi_phase_factor(it, w, vg, tid, NT, NT - 1);
  if constexpr (Feed::kTma) {
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  HADI_SYNC();
  {
    const HadiMap mp = hadi_map(m1, m2, tid, NT);
    if (mp.active) {
      const double pay = hadi_ti(w, TI_PAY)[mp.i];
      for (int j = mp.j0; j < mp.j1; ++j) w.U[j * w.ld + mp.i] = pay;
    }
  }
  HADI_SYNC();
  // two A1 solves per step: the feed sees 2N "steps"
  if constexpr (Feed::kTma) {
    if (feed.producer(tid)) feed.produce(0, 2 * it.N, 0);
  }
  if (tid <= m2) feed.begin_item(2 * it.N, tid);
  for (int n = 1; n <= it.N; ++n) {
    const double e0 = eg[n - 1], e1 = eg[n];
    hadi_cs_predict(it, w, cs, e0, e1, tid, NT, L.scheme);
    HADI_SYNC();
    hadi_phase_solve_a1<0, 0, EXACT>(it, w, e0, e1, 2 * n - 1, tid, NT, feed, bad, nullptr, 2 * it.N);
    HADI_SYNC();
    // hadi_cs_rhs2 rides on the operand fetch of the A2 forward sweep (HadiRhs2 mode 2)
    hadi_phase_solve_a2<0, 0, EXACT>(it, w, tid, NT, bad, HadiRhs2{2, cs.R2, e0, e1});   // Y2 -> U
    HADI_SYNC();
    if (L.scheme == HADI_SCHEME_CS) hadi_cs_correct(it, w, cs, e0, e1, tid, NT);
    else hadi_cs_correct2(it, w, cs, e0, e1, tid, NT, L.scheme);
    HADI_SYNC();
    hadi_phase_solve_a1<0, 0, EXACT>(it, w, e0, e1, 2 * n, tid, NT, feed, bad, nullptr, 2 * it.N);
    HADI_SYNC();
    hadi_phase_solve_a2<0, 0, EXACT>(it, w, tid, NT, bad, HadiRhs2{2, cs.R2, L.scheme == HADI_SCHEME_HV ? e1 : e0, e1});
    HADI_SYNC();
  }
  if constexpr (Feed::kTma) {
    const unsigned nc = (unsigned)(feed.ncf() + feed.ncb());
    feed.issued = feed.consumed = feed.base = feed.base + (unsigned)(2 * it.N) * nc;
  }
  return bar.sync_or((int)bad);
}

// DUO: solves in flight per CTA (teams of NT threads; see HadiBar).  Team t of CTA b is "virtual block" b*DUO + t:
// scratch slot, segment list and profile slot are indexed by it, L.vgrid is their count.
template <int NT, int MINB, int M1, int M2, int FEED, bool GLOBAL = false, int DUO = 1>
__global__ void __launch_bounds__(NT * DUO, MINB) hadi_douglas_kernel(const HadiLaunch L) {
  extern __shared__ double smem[];
  __shared__ int s_item_team[DUO];
  const int team = (DUO > 1) ? (int)threadIdx.x / NT : 0;
  const int tid = (int)threadIdx.x - team * NT;
  // team 0 of every CTA is served first: a batch of up to one solve per SM keeps the second teams idle
  const int vb = team * (int)gridDim.x + (int)blockIdx.x;
  const int vgrid = (DUO > 1) ? L.vgrid : (int)gridDim.x;
  __shared__ int s_turn;
  if (threadIdx.x == 0) s_turn = 0;
  const HadiTurn turn{(DUO > 1 && (L.flags & 1)) ? &s_turn : nullptr};
  int& s_item = s_item_team[team];
  const HadiBar bar{(DUO > 1) ? 1u + (unsigned)team : 0u, (unsigned)NT};
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = 0;
#ifdef HADI_PHASE_TIMING
  tlast = clock64();
#endif

  const int m1 = M1 ? M1 : L.m1, m2 = M2 ? M2 : L.m2;
  HadiView w;
  w.m1 = m1; w.m2 = m2; w.P = (m1 + 1) * (m2 + 1);
  w.ld = hadi_geo_ld(m1); w.n1 = hadi_geo_n1(m1); w.n2 = hadi_geo_n2(m2); w.pj = hadi_geo_pj(m2);
  const HadiSmemLayout lay = hadi_smem_layout(m1, m2, w.ld, w.n1, w.n2, w.pj, FEED == 1, GLOBAL, FEED == 4, M1 > 0, FEED == 5);
  if constexpr (M1 > 0) w.nti = HADI_LEAN ? TI_CORE : TI_COUNT;
  char* sbase = reinterpret_cast<char*>(smem) + (size_t)team * ((lay.total + 127) & ~size_t(127));
  w.zmask = (L.n_items < 0) ? ~0u : 0u;   // a zero the compiler cannot fold
  if constexpr (FEED == 4) {
    w.co_pi = hadi_co_pi(m1);
    w.stg = reinterpret_cast<double*>(sbase + lay.ring);
  }
  double* scratch = L.scratch + (size_t)(vb < vgrid ? vb : 0) * L.scratch_stride;
  const HadiScratchLayout gl = hadi_scratch_layout(m1, m2, w.ld, w.pj, GLOBAL, GLOBAL && L.scheme >= 1);
  // the working set: shared memory, or (GLOBAL) L2-resident global scratch for grids beyond it
  double* Ualloc = GLOBAL ? scratch + gl.U : reinterpret_cast<double*>(sbase + lay.U);
  w.U = Ualloc + HADI_HALO * w.ld + 1;
  w.Y = GLOBAL ? scratch + gl.Y : reinterpret_cast<double*>(sbase + lay.Y);
  w.gstate = GLOBAL;
  w.ti = reinterpret_cast<double*>(sbase + lay.ti);
  w.tj = reinterpret_cast<double*>(sbase + lay.tj);
  w.divk = reinterpret_cast<int*>(sbase + lay.divk);
  if constexpr (M1 > 0) w.tjp = reinterpret_cast<double*>(sbase + lay.tjp);
  w.fM = scratch + gl.fM;
  w.fB = scratch + gl.fB;
  w.lam = scratch + gl.lam;
  HadiCsView cs;
  cs.Y0 = scratch + gl.Y0; cs.R0 = scratch + gl.R0; cs.R1 = scratch + gl.R1; cs.R2 = scratch + gl.R2;
  // zero the halo of U once (payoff initialisation and the sweeps only ever write rows 0..m2)
  for (int k = tid; k < (m2 + 1 + 2 * HADI_HALO) * w.ld + 2; k += NT) Ualloc[k] = 0.0;

  // factor feed of phase S1
  const int ncw = (m2 + 1 + 31) / 32;  // solver warps
  // FEED: 0 plain loads, 1 TMA ring + mbarriers (one producer thread), 3 plain loads behind L1 prefetches,
  //       4 co-operative warps (hadi_phases_fast.cuh)
  typename std::conditional<FEED == 1, HadiRingFeed,
      typename std::conditional<FEED == 3, HadiPrefetchFeed,
          typename std::conditional<FEED == 4, HadiCoopFeed,
              typename std::conditional<FEED == 5, HadiTmemFeed,
                  typename std::conditional<FEED == 6, HadiTmem2Feed,
                      typename std::conditional<FEED == 7, HadiRelayFeed, HadiDirectFeed>::type>::type>::type>::type>::type>::type feed;
  feed.fM = w.fM;
  feed.fB = w.fB;
  feed.pj = w.pj;
  if constexpr (FEED == 3) {
    feed.m1 = m1;
    feed.j = 0;
  }
  if constexpr (FEED == 5 || FEED == 6 || FEED == 7) {
    // tensor memory: 256 (101 x 51) / 128 (51 x 26) columns per CTA — all 512 in the duo kernel —, held until
    // the CTA exits
    static_assert(M1 > 0, "the tensor-memory feeds exist for the grid-specialised variants only");
    static_assert(FEED != 6 || (DUO == 2 && MINB == 1), "FEED 6 needs every column of the SM: one CTA, two teams");
    __shared__ unsigned s_tmem;
    if (threadIdx.x < 32) hadi_tm_alloc(&s_tmem, hadi_tm_cols_feed(FEED, M1));
    hadi_tm_fence_before();
    __syncthreads();
    hadi_tm_fence_after();
    feed.tmem = s_tmem;
    if constexpr (FEED == 6) {
      // the team's chain warps are the two whose CTA-wide warp index lies in tensor-memory quarters 2 team, 2 team + 1
      static_assert(NT % 32 == 0, "teams are whole warps");
      feed.wbase = ((2 * team - team * (NT / 32)) % 4 + 4) % 4;
    }
    if constexpr (FEED == 5) {
      feed.dummy = reinterpret_cast<double*>(sbase + lay.ring);
      for (int k = tid; k < m1 + 2; k += NT) feed.dummy[k] = 0.0;
    }
  }
  if constexpr (FEED == 1) {
    feed.m1 = m1;
    feed.prod_tid = 32 * ncw;
    feed.ring = reinterpret_cast<double*>(sbase + lay.ring);
    feed.full = reinterpret_cast<unsigned long long*>(sbase + lay.bars);
    feed.empty = feed.full + HADI_NS;
    feed.issued = feed.consumed = feed.base = 0;
#ifdef HADI_PHASE_TIMING
    feed.wait_cycles = 0;
#endif
    feed.probe = 0;
    feed.zmask = (L.n_items < 0) ? ~0u : 0u;   // a zero the compiler cannot fold (see HadiRingFeed::release)
    if (tid == 0) {
      for (int k = 0; k < HADI_NS; ++k) {
        hadi_mbar_init(&feed.full[k], 1);
        hadi_mbar_init(&feed.empty[k], m2 + 1);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  HADI_SYNC();

  // Work source: whole items pulled from a global counter, or (split schedule) this CTA's static list of
  // segments — McNaughton's wrap-around rule on the host cuts the one solve that straddles the end of a CTA's
  // share of the batch in two, so that every CTA finishes at the same time instead of after a whole number of
  // solves (hadi_host.cpp: build_split_schedule).
  const bool split = (L.segs != nullptr) && !GLOBAL && FEED != 1 && FEED != 4;
  int seg_q = (split && vb < vgrid) ? L.seg_off[vb] : 0;
  const int seg_end = (split && vb < vgrid) ? L.seg_off[vb + 1] : 0;
  for (;;) {
    if (DUO > 1 && vb >= vgrid) break;   // a team beyond the virtual grid (odd number of work slots) has nothing to do
    HadiSegment sg;
    if (split) {
      if (seg_q >= seg_end) break;
      sg = L.segs[seg_q++];
    } else {
      if (tid == 0) s_item = atomicAdd(L.counter, 1);
      HADI_SYNC();
      sg.item = s_item;
      if (sg.item >= L.n_items) break;
      sg.n0 = 1; sg.n1 = 0; sg.hin = -1; sg.hout = -1;
    }
    const HadiItem it = L.items[sg.item];
    if (!split) sg.n1 = it.N;
    HADI_TICK(0)
    // state left by the CTA that ran steps 1..n0-1: wait for it (bounded: after 100 ms — fifty times a config-2 launch —
    // or if that CTA's guarded divisions left their range, this CTA solves the item from the payoff instead)
    bool whole_exact = false;
    const double* hin = nullptr;
    if (sg.hin >= 0) {
      if (tid == 0) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        int st;
        for (;;) {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(st) : "l"(L.hand_state + sg.hin) : "memory");
          if (st != HADI_HAND_PENDING) break;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
          if (t1 - t0 > 100000000ull) { st = HADI_HAND_BAD; break; }
          __nanosleep(200);
        }
        s_item = st;
      }
      HADI_SYNC();
      if (s_item == HADI_HAND_READY) hin = L.hand_data + (size_t)sg.hin * 2 * (size_t)w.P;
      else whole_exact = true;
      HADI_SYNC();
    }
    // fast pass; if any guarded division left its range (never observed on option data at these grids), the
    // item is re-solved with IEEE divisions so that the published value is exact in every case
    bool cs_done = false;
    if constexpr (GLOBAL) {
      if (L.scheme >= 1) {
        if (hadi_solve_item_cs<NT, false>(L, it, w, cs, feed, tid, bar)) hadi_solve_item_cs<NT, true>(L, it, w, cs, feed, tid, bar);
        cs_done = true;
      }
    }
    if (!cs_done) {
#ifdef HADI_FORCE_EXACT
      if (!whole_exact) hadi_solve_item<NT, M1, M2, true>(L, it, w, feed, tid, tacc, tlast, sg.n0, sg.n1, hin, bar, turn);
#else
      if (!whole_exact) {
#ifdef HADI_NO_RERUN   /* timing experiments only */
        whole_exact = hadi_solve_item<NT, M1, M2, false>(L, it, w, feed, tid, tacc, tlast, sg.n0, sg.n1, hin, bar, turn) && L.n_items < 0;
#else
        whole_exact = hadi_solve_item<NT, M1, M2, false>(L, it, w, feed, tid, tacc, tlast, sg.n0, sg.n1, hin, bar, turn);
#endif
      }
#endif
      if (L.dbg_step == -7) whole_exact = true;   // test hook (HADI_DEBUG_STOP=-7:0): treat every fast pass as out of range
      if (whole_exact) {
        if (tid == 0 && L.reruns != nullptr) atomicAdd(L.reruns, 1ULL);
        // only the CTA that holds the last steps publishes: it re-solves the whole item
        if (sg.hout < 0) hadi_solve_item<NT, M1, M2, true>(L, it, w, feed, tid, tacc, tlast, 1, it.N, nullptr, bar, turn);
      }
    }

    if (sg.hout >= 0) {
      // hand the state after step n1 to the CTA that runs the remaining steps
      double* ho = L.hand_data + (size_t)sg.hout * 2 * (size_t)w.P;
      if (!whole_exact) {
        const HadiMap mp = hadi_map(m1, m2, tid, NT);
        if (mp.active) {
          for (int j = mp.j0; j < mp.j1; ++j) {
            const size_t p = (size_t)j * (m1 + 1) + mp.i;
            __stcg(ho + p, w.U[j * w.ld + mp.i]);
            if (it.style == 1) __stcg(ho + (size_t)w.P + p, w.Y[j * w.ld + mp.i]);   // lambda sits in Y between steps
          }
        }
      }
      HADI_SYNC();
      if (tid == 0) {
        __threadfence();
        const int st = whole_exact ? HADI_HAND_BAD : HADI_HAND_READY;
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(L.hand_state + sg.hout), "r"(st) : "memory");
      }
      HADI_SYNC();
      HADI_TICK(6)
      continue;
    }

    if (tid == 0) hadi_publish(L, it, w);
#ifdef HADI_DEBUG_TRACE
    if (false) {
#else
    if (L.out_U != nullptr || L.out_lam != nullptr) {
#endif
      const HadiMap mp = hadi_map(m1, m2, tid, NT);
      if (mp.active) {
        for (int j = mp.j0; j < mp.j1; ++j) {
          const size_t p = (size_t)it.out * w.P + (size_t)j * (m1 + 1) + mp.i;
          if (L.out_U != nullptr) L.out_U[p] = w.U[j * w.ld + mp.i];
#ifdef HADI_DEBUG_STOP
          if (L.out_lam != nullptr && L.dbg_step > 0) {
            L.out_lam[p] = (L.dbg_phase >= 100) ? w.lam[j * w.ld + mp.i] : w.Y[j * w.ld + mp.i];
            continue;
          }
#endif
          if (L.out_lam != nullptr && it.style == 1) L.out_lam[p] = w.lam[j * w.ld + mp.i];
        }
      }
    }
    HADI_SYNC();  // everyone is done with s_item, U and the tables before the next item
    HADI_TICK(6)
  }
  if constexpr (FEED == 5 || FEED == 6 || FEED == 7) {
    hadi_tm_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) hadi_tm_free(feed.tmem, hadi_tm_cols_feed(FEED, M1));
  }
#ifdef HADI_PHASE_TIMING
  if constexpr (FEED == 1) tacc[6] = feed.wait_cycles;   // slot 6 reports the ring wait of solver thread 0
  if (tid == 0 && L.prof != nullptr && vb < vgrid)
    for (int k = 0; k < 8; ++k) L.prof[(size_t)vb * 8 + k] = tacc[k];
#endif
}

// ---- cluster kernel -------------------------------------------------------------------------------
// One solve on a thread-block CLUSTER: grids beyond shared memory (e.g. 401 x 201, 645 KB per array) keep their
// working arrays in global scratch, so nothing ties a solve to one SM.  The HADI_CLUSTER CTAs of a cluster share
// one scratch block; point-wise phases run over all their threads (same (i, q) mapping, 8 x 256 threads with up to 255 registers each), the
// lines of the implicit solves are dealt round-robin to the CTAs (HadiView::line_mul / line_off), every CTA
// keeps its own copy of the small coefficient tables in shared memory, and the phases are separated by
// barrier.cluster (release / acquire at cluster scope: global writes of the other CTAs become visible and the
// L1 is invalidated).  Same phase functions, same arithmetic: results are bit-identical to variant 5.
__device__ __forceinline__ void hadi_csync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int NT, bool EXACT>
__device__ __forceinline__ bool hadi_cluster_solve(const HadiLaunch& L, const HadiItem& it, HadiView& w,
                                                   const HadiCsView& cs, int tid, int gtid, int gnt, int* mail) {
  const int m1 = L.m1, m2 = L.m2;
  const double* sg = L.s_pool + it.s_off;
  const double* vg = L.v_pool + it.v_off;
  const double* eg = L.e_pool + it.e_off;
  w.c = it.theta * it.dt;
  const double rdt = hadi_rcp_prep(it.dt);
  unsigned bad = 0;
  HadiDirectFeed feed;
  feed.fM = w.fM; feed.fB = w.fB; feed.pj = w.pj;
  hadi_phase_tables(it, w, sg, vg, tid, NT);          // per-CTA tables (the A2 scratch tables live in Y: every
  hadi_csync();                                        // CTA writes the same values there)
  hadi_phase_factor(it, w, vg, tid, NT, NT - 1);
  hadi_csync();
  {
    const HadiMap mp = hadi_map(m1, m2, gtid, gnt);
    if (mp.active) {
      const double pay = hadi_ti(w, TI_PAY)[mp.i];
      for (int j = mp.j0; j < mp.j1; ++j) {
        w.U[j * w.ld + mp.i] = pay;
        if (it.style == 1) {
          w.lam[j * w.ld + mp.i] = 0.0;
          w.Y[j * w.ld + mp.i] = 0.0;
        }
      }
    }
  }
  hadi_csync();
  for (int n = 1; n <= it.N; ++n) {
    const double e0 = eg[n - 1], e1 = eg[n];
    if (L.scheme >= 1) {
      hadi_cs_predict(it, w, cs, e0, e1, gtid, gnt, L.scheme);
      hadi_csync();
      hadi_phase_solve_a1<0, 0, EXACT>(it, w, e0, e1, 2 * n - 1, tid, NT, feed, bad, nullptr, 2 * it.N);
      hadi_csync();
      hadi_cs_rhs2(it, w, cs, e0, e1, gtid, gnt);
      hadi_csync();
      hadi_phase_solve_a2<0, 0, EXACT>(it, w, tid, NT, bad);
      hadi_csync();
      if (L.scheme == HADI_SCHEME_CS) hadi_cs_correct(it, w, cs, e0, e1, gtid, gnt);
      else hadi_cs_correct2(it, w, cs, e0, e1, gtid, gnt, L.scheme);
      hadi_csync();
      hadi_phase_solve_a1<0, 0, EXACT>(it, w, e0, e1, 2 * n, tid, NT, feed, bad, nullptr, 2 * it.N);
      hadi_csync();
      hadi_cs_rhs2(it, w, cs, L.scheme == HADI_SCHEME_HV ? e1 : e0, e1, gtid, gnt);
      hadi_csync();
      hadi_phase_solve_a2<0, 0, EXACT>(it, w, tid, NT, bad);
      hadi_csync();
    } else {
      hadi_phase_explicit<0, 0>(it, w, e0, e1, gtid, gnt);
      hadi_csync();
      hadi_phase_solve_a1<0, 0, EXACT>(it, w, e0, e1, n, tid, NT, feed, bad, nullptr, 0);
      hadi_csync();
      hadi_phase_rhs2<0, 0>(it, w, e0, e1, gtid, gnt);
      hadi_csync();
      hadi_phase_solve_a2<0, 0, EXACT>(it, w, tid, NT, bad);
      if (it.bc && tid * w.line_mul + w.line_off == 0) hadi_dirichlet_col0(w, it.K * eg[it.N + 1 + n]);
      hadi_csync();
      if (it.style == 1) {
        hadi_phase_project<0, 0, EXACT, HADI_CHP>(it, w, rdt, gtid, gnt, bad);
        hadi_csync();
      }
    }
  }
  // cluster-wide vote on the guarded divisions
  if (__syncthreads_or((int)bad) != 0 && tid == 0) atomicOr(mail + 1, 1);
  hadi_csync();
  return *reinterpret_cast<volatile int*>(mail + 1) != 0;
}

template <int NT>
__global__ void __cluster_dims__(HADI_CLUSTER, 1, 1) __launch_bounds__(NT, 1) hadi_cluster_kernel(const HadiLaunch L) {
  extern __shared__ double smem[];
  const int tid = threadIdx.x;
  unsigned rank_u;
  asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank_u));
  const int rank = (int)rank_u;
  const int cid = blockIdx.x / HADI_CLUSTER;
  const int gtid = rank * NT + tid, gnt = HADI_CLUSTER * NT;
  const int m1 = L.m1, m2 = L.m2;
  HadiView w;
  w.m1 = m1; w.m2 = m2; w.P = (m1 + 1) * (m2 + 1);
  w.ld = hadi_geo_ld(m1); w.n1 = hadi_geo_n1(m1); w.n2 = hadi_geo_n2(m2); w.pj = hadi_geo_pj(m2);
  w.line_mul = HADI_CLUSTER; w.line_off = rank;
  // per-CTA A2 assembly scratch: its own region of the scratch block (hadi_ts() addresses it relative to Y)
  const HadiSmemLayout lay = hadi_smem_layout(m1, m2, w.ld, w.n1, w.n2, w.pj, false, true);
  char* sbase = reinterpret_cast<char*>(smem);
  double* scratch = L.scratch + (size_t)cid * L.scratch_stride;
  const HadiScratchLayout gl = hadi_scratch_layout(m1, m2, w.ld, w.pj, true, L.scheme >= 1);
  double* Ualloc = scratch + gl.U;
  w.U = Ualloc + HADI_HALO * w.ld + 1;
  w.Y = scratch + gl.Y;
  w.ts_off = (int)(gl.ts - gl.Y) + rank * TS_COUNT * w.n2;
  w.ti = reinterpret_cast<double*>(sbase + lay.ti);
  w.tj = reinterpret_cast<double*>(sbase + lay.tj);
  w.divk = reinterpret_cast<int*>(sbase + lay.divk);
  w.fM = scratch + gl.fM;
  w.fB = scratch + gl.fB;
  w.lam = scratch + gl.lam;
  HadiCsView cs;
  cs.Y0 = scratch + gl.Y0; cs.R0 = scratch + gl.R0; cs.R1 = scratch + gl.R1; cs.R2 = scratch + gl.R2;
  int* mail = reinterpret_cast<int*>(scratch + gl.mail);
  for (int k = gtid; k < (m2 + 1 + 2 * HADI_HALO) * w.ld + 2; k += gnt) Ualloc[k] = 0.0;
  for (;;) {
    if (gtid == 0) {
      mail[0] = atomicAdd(L.counter, 1);
      mail[1] = 0;
    }
    hadi_csync();
    const int item = *reinterpret_cast<volatile int*>(mail);
    if (item >= L.n_items) break;
    const HadiItem it = L.items[item];
    if (hadi_cluster_solve<NT, false>(L, it, w, cs, tid, gtid, gnt, mail)) {
      if (gtid == 0 && L.reruns != nullptr) atomicAdd(L.reruns, 1ULL);
      hadi_cluster_solve<NT, true>(L, it, w, cs, tid, gtid, gnt, mail);
    }
    if (gtid == 0) hadi_publish(L, it, w);
    if (L.out_U != nullptr || L.out_lam != nullptr) {
      const HadiMap mp = hadi_map(m1, m2, gtid, gnt);
      if (mp.active) {
        for (int j = mp.j0; j < mp.j1; ++j) {
          const size_t p = (size_t)it.out * w.P + (size_t)j * (m1 + 1) + mp.i;
          if (L.out_U != nullptr) L.out_U[p] = w.U[j * w.ld + mp.i];
          if (L.out_lam != nullptr && it.style == 1) L.out_lam[p] = w.lam[j * w.ld + mp.i];
        }
      }
    }
    hadi_csync();   // everyone is done with the mailbox, U and the tables before the next item
  }
}

// ---- variants ----------------------------------------------------------------------------------
// 0: 101 x 51 nodes (BASELINE configs 1, 2, 5): 320 threads = 3 row-chunks x 101 columns (+17),
//    2 CTAs/SM, <= 102 registers (no spills: L1 is all but gone at this shared-memory carve-out)
// 1:  51 x 26 nodes (the reference's own test / benchmark grid): 256 threads = 5 x 51 (+1), 3 CTAs/SM
// 11: the same grid with 128 threads = 2 x 51 (+26) and 6 CTAs/SM: a solve takes 40 % longer, twice as many fill the chain
//     phases of the others — 9 to 12 % more throughput on batches of two thousand solves (tools/time_51x26.py); chosen by
//     hadi_douglas_plan for batches that fill its 888 slots
// 2: any grid with m1+1 <= 416 that fits shared memory, run-time dimensions, direct factor loads
// 3: any grid with m1+1 <= 1024 that fits shared memory, run-time dimensions, one CTA per SM
// 5: any grid with m1+1 <= 512: U and Y in L2-resident global scratch, tables in shared memory, TMA ring for
//    the A1 factors, one CTA of 512 threads (128 registers) per SM (grids beyond shared memory, e.g. 401 x 201;
//    all Craig-Sneyd solves)
// 6: the same with 1024 threads (64 registers: the generic phases spill) for m1+1 <= 1024
// 7: the cluster kernel (hadi_cluster_kernel), chosen by hadi_douglas_plan when there are few items
// Factor feed of the grid-specialised variants, measured on B200 (round 1): at 101x51 plain loads (2.68 ms for
// config 2) beat the TMA ring (2.78 ms: mbarrier try_wait costs ~90 cycles per chunk on the dependent chain),
// per-thread cp.async stages (2.97 ms) and L1 prefetches (2.87 ms); at 51x26 the L1 prefetch wins.
#ifndef HADI_FEED0
#define HADI_FEED0 7   /* 101 x 51: back-substitution stream in tensor memory, relayed between warp pairs */
#endif
#ifndef HADI_FEED1
#define HADI_FEED1 3
#endif
#ifndef HADI_V1_NT
#define HADI_V1_NT 256     /* 51 x 26: threads per CTA and CTAs per SM */
#define HADI_V1_MINB 3
#endif
#ifndef HADI_DUO_NT
#define HADI_DUO_NT 320   /* threads per team of the duo kernel (2 x 256 threads with 128 registers each measured slower) */
#endif
#ifndef HADI_DUO
#define HADI_DUO 0   /* 1: also build variant 8 (two solves per CTA, all of S1 out of tensor memory) and let the planner
                        prefer it; measured slower than variant 0 with the relay feed */
#endif
#if HADI_DUO
#define HADI_DUO_VARIANT(X) X(8, HADI_DUO_NT, 1, 100, 50, 6, false, 2)
#else
#define HADI_DUO_VARIANT(X)
#endif
// X(id, threads per team, min CTAs/SM, m1, m2, feed, global state, teams per CTA)
#define HADI_VARIANTS(X)                    \
  X(0, 320, 2, 100, 50, HADI_FEED0, false, 1) \
  X(1, HADI_V1_NT, HADI_V1_MINB, 50, 25, HADI_FEED1, false, 1)  \
  X(11, 128, 6, 50, 25, HADI_FEED1, false, 1)  \
  X(2, 416, 2, 0, 0, 0, false, 1)          \
  X(3, 1024, 1, 0, 0, 0, false, 1)         \
  X(4, 320, 2, 100, 50, 4, false, 1)       \
  X(5, 512, 1, 0, 0, 1, true, 1)           \
  X(6, 1024, 1, 0, 0, 1, true, 1)          \
  HADI_DUO_VARIANT(X)

struct VariantInfo {
  int id;
  int threads, minb, m1, m2;
  int feed;   // 0 plain loads, 1 TMA ring, 3 plain loads behind L1 prefetches, 4 co-operative warps, 5 / 6 tensor memory
  bool global_state;
  int duo;    // teams (solves in flight) per CTA
  const void* fn;
};
const VariantInfo* variants() {
  static const VariantInfo v[] = {
#define X(id, nt, minb, a, b, r, g, d) {id, nt, minb, a, b, r, g, d, (const void*)hadi_douglas_kernel<nt, minb, a, b, r, g, d>},
      HADI_VARIANTS(X)
#undef X
  };
  return v;
}
constexpr int kNumVariants = 8 + HADI_DUO;   // entries of the table above
constexpr int kClusterVariant = 7;   // hadi_cluster_kernel: one solve per thread-block cluster
constexpr int kDuoVariant = 8;
const VariantInfo* variant_by_id(int id) {
  const VariantInfo* v = variants();
  for (int k = 0; k < kNumVariants; ++k)
    if (v[k].id == id) return &v[k];
  return nullptr;
}
constexpr int kClusterThreads = 256; // few threads, many registers: the generic phases spill badly at 64 registers

}  // namespace

int hadi_douglas_plan(int device, int m1, int m2, int ld, int n1, int n2, int pj, bool need_global, HadiPlan* plan,
                      bool want_cluster, bool many) {
  int max_smem = 0, sms = 0;
  cudaError_t e = cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  if (e != cudaSuccess) return (int)e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) return (int)e;
  const VariantInfo* v = variants();
  int pick = -1;
  size_t smem = 0;
  {
    // the cluster kernel: global working set, HADI_CLUSTER CTAs of 1024 threads per solve
    const char* fv = getenv("HADI_FORCE_VARIANT");
    const bool forced = fv && atoi(fv) == kClusterVariant;
    if ((want_cluster && need_global && !(fv && !forced)) || forced) {
      const size_t sm = hadi_smem_layout(m1, m2, ld, n1, n2, pj, false, true).total;
      if (sm <= (size_t)max_smem && m1 + 1 <= kClusterThreads * HADI_CLUSTER && m2 + 1 < kClusterThreads * HADI_CLUSTER) {
        e = cudaFuncSetAttribute((const void*)hadi_cluster_kernel<kClusterThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return (int)e;
        plan->global_state = true;
        plan->variant = kClusterVariant;
        plan->threads = kClusterThreads;
        plan->ctas_per_sm = 1;
        plan->sm_count = sms;
        plan->smem_bytes = sm;
        plan->cluster = HADI_CLUSTER;
        return 0;
      }
    }
  }
  // development aid: HADI_FORCE_VARIANT=<id> restricts the choice (e.g. 2 = run-time dims, direct loads)
  const char* force = getenv("HADI_FORCE_VARIANT");
  const char* noduo = getenv("HADI_NO_DUO");
  auto eligible = [&](const VariantInfo& q, size_t* bytes) -> bool {
    if (force && atoi(force) != q.id) return false;
    if (need_global && !q.global_state) return false;
    if (q.m1 != 0 && (q.m1 != m1 || q.m2 != m2)) return false;
    if (q.threads < m1 + 1 || q.threads - 1 <= m2) return false;
    if (q.feed == 1 && q.threads <= 32 * ((m2 + 1 + 31) / 32)) return false;   // needs a producer thread past the solver warps
    const size_t one = hadi_smem_layout(m1, m2, ld, n1, n2, pj, q.feed == 1, q.global_state, q.feed == 4, q.m1 != 0, q.feed == 5).total;
    *bytes = q.duo > 1 ? (size_t)q.duo * ((one + 127) & ~size_t(127)) : one;
    return *bytes <= (size_t)max_smem;
  };
  // first choice where it exists: two solves per CTA with phase S1 fed from tensor memory (variant 8)
  if (HADI_DUO && !(noduo && atoi(noduo) != 0)) {
    for (int k = 0; k < kNumVariants && pick < 0; ++k)
      if (v[k].duo > 1 && eligible(v[k], &smem)) pick = k;
  }
  for (int k = 0; k < kNumVariants && pick < 0; ++k) {
    if (v[k].duo > 1 && !(force && atoi(force) == v[k].id)) continue;
    // two instantiations for the 51 x 26 grid: the one with more, smaller CTAs only when the batch fills them
    if (!force && v[k].id == 1 && many) continue;
    if (!force && v[k].id == 11 && !many) continue;
    if (eligible(v[k], &smem)) pick = k;
  }
  if (pick < 0) return -1;
  e = cudaFuncSetAttribute(v[pick].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v[pick].fn, v[pick].threads * v[pick].duo, smem);
  if (e != cudaSuccess) return (int)e;
  if (occ < 1) return -1;
  if (v[pick].feed == 5 || v[pick].feed == 7) {
    // The occupancy calculator knows nothing about tensor memory and answers 1 for a kernel that executes
    // tcgen05.alloc; the hardware co-schedules CTAs as long as registers and shared memory fit, and each CTA's
    // allocation of hadi_tm_cols() columns then succeeds while the SM's 512 columns last (measured on B200
    // with tools/ubench_tmem.cu: two CTAs x 256 columns resident on all 148 SMs).  Residency from the launch
    // bounds (registers), the shared-memory footprint and the column budget:
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, v[pick].fn);
    if (e != cudaSuccess) return (int)e;
    int smem_sm = 0, regs_sm = 0;
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
    cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, device);
    const int by_smem = (int)((size_t)smem_sm / (smem + fa.sharedSizeBytes + 1024));
    const int by_regs = regs_sm / (((fa.numRegs + 7) & ~7) * v[pick].threads);
    const int by_tmem = 512 / hadi_tm_cols_feed(v[pick].feed, v[pick].m1);
    occ = std::min(std::min(by_smem, by_regs), std::min(by_tmem, v[pick].minb));
    if (occ < 1) return -1;
  }
  if (v[pick].feed == 6) occ = 1;   // the CTA owns all 512 tensor-memory columns of its SM
  plan->global_state = v[pick].global_state;
  plan->variant = v[pick].id;
  plan->threads = v[pick].threads * v[pick].duo;
  plan->ctas_per_sm = occ * v[pick].duo;   // work slots (solves in flight) per SM
  plan->sm_count = sms;
  plan->smem_bytes = smem;
  plan->cluster = 1;
  plan->duo = v[pick].duo;
  return 0;
}

int hadi_launch_douglas(const HadiLaunch& L_in, const HadiPlan& plan, int grid_ctas, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  HadiLaunch L = L_in;
  L.vgrid = grid_ctas;   // work slots; the duo kernel packs two per CTA
  // the run-time-dimension kernels serve many grid shapes and plans are cached by the host layer: the dynamic
  // shared-memory limit of the function must be the one of THIS plan, not of the plan made last
  {
    const VariantInfo* q = variant_by_id(plan.variant);
    if (plan.variant != kClusterVariant && q == nullptr) return (int)cudaErrorInvalidValue;
    const void* fn = plan.variant == kClusterVariant ? (const void*)hadi_cluster_kernel<kClusterThreads> : q->fn;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);
    if (e != cudaSuccess) return (int)e;
  }
  switch (plan.variant) {
#define X(id, nt, minb, a, b, r, g, d) \
  case id:                             \
    hadi_douglas_kernel<nt, minb, a, b, r, g, d><<<d > 1 ? std::min(grid_ctas, std::max(plan.sm_count, (grid_ctas + d - 1) / d)) : grid_ctas, nt * d, plan.smem_bytes, st>>>(L); \
    break;
    HADI_VARIANTS(X)
#undef X
    case kClusterVariant:
      hadi_cluster_kernel<kClusterThreads><<<grid_ctas, kClusterThreads, plan.smem_bytes, st>>>(L);
      break;
    default:
      return (int)cudaErrorInvalidValue;
  }
  return (int)cudaGetLastError();
}
