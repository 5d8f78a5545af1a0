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

// One item, payoff to price.  Returns (CTA-uniform) whether any guarded division left its fast-path
// range; EXACT = true compiles every division as IEEE '/'.
// FEED == 4: co-operative S1 / pipelined S2 of hadi_phases_fast.cuh (grid-specialised variants only)
struct HadiCoopFeed : HadiDirectFeed {
  static constexpr bool kCoop = true;
};
template <class F> struct hadi_is_coop { static constexpr bool value = false; };
template <> struct hadi_is_coop<HadiCoopFeed> { static constexpr bool value = true; };

// n0..n1: the time steps to run (1..N for a whole solve); hin: state to resume from (nullptr: start from the payoff)
template <int NT, int M1, int M2, bool EXACT, class Feed>
__device__ __forceinline__ bool hadi_solve_item(const HadiLaunch& L, const HadiItem& it, HadiView& w, Feed& feed,
                                                int tid, long long* tacc, long long& tlast, int n0, int n1,
                                                const double* hin) {
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
  __syncthreads();
  hadi_phase_factor(it, w, vg, tid, NT, NT - 1);
  if constexpr (Feed::kTma) {
    // the factor streams were written with generic stores and will be read by TMA (async proxy)
#ifndef HADI_NO_THREADFENCE
    __threadfence();
#endif
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();  // the A2 assembly keeps scratch tables in Y: finish it before Y is initialised
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
  __syncthreads();
  if constexpr (Feed::kTma) {
    if (feed.producer(tid)) feed.produce(0, it.N, 0);  // first chunks are in flight before step 1
  }
  if (tid <= m2) feed.begin_item(it.N, tid);
  HADI_TICK(0)

  int div_cur = 0;
  bool stopped = false;
  if (it.nd > 0)   // dividend queue position after the steps another CTA ran
    for (int n = 1; n < n0; ++n) hadi_dividend_at(n, it.dt, it.nd, L.div_dates, div_cur);
  for (int n = n0; n <= n1; ++n) {
    if (it.nd > 0) {
      const int hit = hadi_dividend_at(n, it.dt, it.nd, L.div_dates, div_cur);
      if (hit >= 0) {  // uniform across the CTA
        hadi_phase_div1(w, L.div_amounts[hit], L.div_pcts[hit], tid, NT);
        __syncthreads();
        HADI_STOP(1)
        hadi_phase_div2(w, tid, NT);
        __syncthreads();
        HADI_STOP(2)
        if (it.style == 1) {
          hadi_phase_div3(w, tid, NT);
          __syncthreads();
        }
        HADI_STOP(3)
      }
    }
    HADI_TICK(0)
#ifdef HADI_DEBUG_LAM
    if (it.style == 1) hadi_dbg_check_lam(L, w, n, it.out, tid, NT);
#endif
    const double e0 = eg[n - 1], e1 = eg[n];
    hadi_phase_explicit<M1, M2>(it, w, e0, e1, tid, NT);
    __syncthreads();
    HADI_TICK(2)
    HADI_STOP(4)
    if constexpr (hadi_is_coop<Feed>::value && M1 > 0) {
      hadi_fast_solve_a1<M1, M2, EXACT>(it, w, n - 1, tid, bad, &tacc[1]);
    } else {
      hadi_phase_solve_a1<M1, M2, EXACT>(it, w, e0, e1, n, tid, NT, feed, bad, &tacc[1]);
    }
    __syncthreads();
    HADI_TICK(3)
    HADI_STOP(5)
#ifdef HADI_SPLIT_R
    hadi_phase_rhs2<M1, M2>(it, w, e0, e1, tid, NT);
    __syncthreads();
#else
    if constexpr (M1 == 0) {
      hadi_phase_rhs2<M1, M2>(it, w, e0, e1, tid, NT);
      __syncthreads();
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
      hadi_phase_solve_a2<M1, M2, EXACT>(it, w, tid, NT, bad);
    }
    __syncthreads();
    HADI_TICK(4)
    HADI_STOP(7)
    if (it.style == 1) {
      // all multiplier loads of a thread's rows are issued before the first is used (one L2 round trip)
      constexpr int CHP = (M1 == 100 && M2 == 50 && NT / 101 == 3) ? 17 : HADI_CHP;
      hadi_phase_project<M1, M2, EXACT, CHP>(it, w, rdt, tid, NT, bad);
      __syncthreads();
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
      __syncthreads();
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
      __syncthreads();
    } else
#endif
    {
      // every chunk issued for this item has been consumed; keep both counters in step on all threads
      const unsigned nc = (unsigned)(feed.ncf() + feed.ncb());
      feed.issued = feed.consumed = feed.base = feed.base + (unsigned)it.N * nc;
    }
  }
  (void)stopped;
  return __syncthreads_or((int)bad) != 0;
}

// Craig-Sneyd (European, no dividends) on the global-state working set: src/solver.hpp:781-907.
template <int NT, bool EXACT, class Feed>
__device__ __forceinline__ bool hadi_solve_item_cs(const HadiLaunch& L, const HadiItem& it, HadiView& w,
                                                   const HadiCsView& cs, Feed& feed, int tid) {
  const int m1 = L.m1, m2 = L.m2;
  const double* sg = L.s_pool + it.s_off;
  const double* vg = L.v_pool + it.v_off;
  const double* eg = L.e_pool + it.e_off;
  w.c = it.theta * it.dt;
  unsigned bad = 0;
  hadi_phase_tables(it, w, sg, vg, tid, NT);
  __syncthreads();
  hadi_phase_factor(it, w, vg, tid, NT, NT - 1);
  if constexpr (Feed::kTma) {
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();
  {
    const HadiMap mp = hadi_map(m1, m2, tid, NT);
    if (mp.active) {
      const double pay = hadi_ti(w, TI_PAY)[mp.i];
      for (int j = mp.j0; j < mp.j1; ++j) w.U[j * w.ld + mp.i] = pay;
    }
  }
  __syncthreads();
  // two A1 solves per step: the feed sees 2N "steps"
  if constexpr (Feed::kTma) {
    if (feed.producer(tid)) feed.produce(0, 2 * it.N, 0);
  }
  if (tid <= m2) feed.begin_item(2 * it.N, tid);
  for (int n = 1; n <= it.N; ++n) {
    const double e0 = eg[n - 1], e1 = eg[n];
    hadi_cs_predict(it, w, cs, e0, e1, tid, NT);
    __syncthreads();
    hadi_phase_solve_a1<0, 0, EXACT>(it, w, e0, e1, 2 * n - 1, tid, NT, feed, bad, nullptr, 2 * it.N);
    __syncthreads();
    hadi_cs_rhs2(it, w, cs, e0, e1, tid, NT);
    __syncthreads();
    hadi_phase_solve_a2<0, 0, EXACT>(it, w, tid, NT, bad);   // Y2 -> U
    __syncthreads();
    hadi_cs_correct(it, w, cs, e0, e1, tid, NT);
    __syncthreads();
    hadi_phase_solve_a1<0, 0, EXACT>(it, w, e0, e1, 2 * n, tid, NT, feed, bad, nullptr, 2 * it.N);
    __syncthreads();
    hadi_cs_rhs2(it, w, cs, e0, e1, tid, NT);
    __syncthreads();
    hadi_phase_solve_a2<0, 0, EXACT>(it, w, tid, NT, bad);
    __syncthreads();
  }
  if constexpr (Feed::kTma) {
    const unsigned nc = (unsigned)(feed.ncf() + feed.ncb());
    feed.issued = feed.consumed = feed.base = feed.base + (unsigned)(2 * it.N) * nc;
  }
  return __syncthreads_or((int)bad) != 0;
}

template <int NT, int MINB, int M1, int M2, int FEED, bool GLOBAL = false>
__global__ void __launch_bounds__(NT, MINB) hadi_douglas_kernel(const HadiLaunch L) {
  extern __shared__ double smem[];
  __shared__ int s_item;
  const int tid = threadIdx.x;
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = 0;
#ifdef HADI_PHASE_TIMING
  tlast = clock64();
#endif

  const int m1 = M1 ? M1 : L.m1, m2 = M2 ? M2 : L.m2;
  HadiView w;
  w.m1 = m1; w.m2 = m2; w.P = (m1 + 1) * (m2 + 1);
  w.ld = hadi_geo_ld(m1); w.n1 = hadi_geo_n1(m1); w.n2 = hadi_geo_n2(m2); w.pj = hadi_geo_pj(m2);
  const HadiSmemLayout lay = hadi_smem_layout(m1, m2, w.ld, w.n1, w.n2, w.pj, FEED == 1, GLOBAL, FEED == 4, M1 > 0);
  if constexpr (M1 > 0) w.nti = HADI_LEAN ? TI_CORE : TI_COUNT;
  char* sbase = reinterpret_cast<char*>(smem);
  w.zmask = (L.n_items < 0) ? ~0u : 0u;   // a zero the compiler cannot fold
  if constexpr (FEED == 4) {
    w.co_pi = hadi_co_pi(m1);
    w.stg = reinterpret_cast<double*>(sbase + lay.ring);
  }
  double* scratch = L.scratch + (size_t)blockIdx.x * L.scratch_stride;
  const HadiScratchLayout gl = hadi_scratch_layout(m1, m2, w.ld, w.pj, GLOBAL, GLOBAL && L.scheme == 1);
  // the working set: shared memory, or (GLOBAL) L2-resident global scratch for grids beyond it
  double* Ualloc = GLOBAL ? scratch + gl.U : reinterpret_cast<double*>(sbase + lay.U);
  w.U = Ualloc + HADI_HALO * w.ld + 1;
  w.Y = GLOBAL ? scratch + gl.Y : reinterpret_cast<double*>(sbase + lay.Y);
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
          typename std::conditional<FEED == 4, HadiCoopFeed, HadiDirectFeed>::type>::type>::type feed;
  feed.fM = w.fM;
  feed.fB = w.fB;
  feed.pj = w.pj;
  if constexpr (FEED == 3) {
    feed.m1 = m1;
    feed.j = 0;
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
  __syncthreads();

  // Work source: whole items pulled from a global counter, or (split schedule) this CTA's static list of
  // segments — McNaughton's wrap-around rule on the host cuts the one solve that straddles the end of a CTA's
  // share of the batch in two, so that every CTA finishes at the same time instead of after a whole number of
  // solves (hadi_host.cpp: build_split_schedule).
  const bool split = (L.segs != nullptr) && !GLOBAL && FEED != 1 && FEED != 4;
  int seg_q = split ? L.seg_off[blockIdx.x] : 0;
  const int seg_end = split ? L.seg_off[blockIdx.x + 1] : 0;
  for (;;) {
    HadiSegment sg;
    if (split) {
      if (seg_q >= seg_end) break;
      sg = L.segs[seg_q++];
    } else {
      if (tid == 0) s_item = atomicAdd(L.counter, 1);
      __syncthreads();
      sg.item = s_item;
      if (sg.item >= L.n_items) break;
      sg.n0 = 1; sg.n1 = 0; sg.hin = -1; sg.hout = -1;
    }
    const HadiItem it = L.items[sg.item];
    if (!split) sg.n1 = it.N;
    HADI_TICK(0)
    // state left by the CTA that ran steps 1..n0-1: wait for it (bounded: after two seconds, or if that CTA's
    // guarded divisions left their range, this CTA solves the item from the payoff instead)
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
          if (t1 - t0 > 2000000000ull) { st = HADI_HAND_BAD; break; }
          __nanosleep(200);
        }
        s_item = st;
      }
      __syncthreads();
      if (s_item == HADI_HAND_READY) hin = L.hand_data + (size_t)sg.hin * 2 * (size_t)w.P;
      else whole_exact = true;
      __syncthreads();
    }
    // fast pass; if any guarded division left its range (never observed on option data at these grids), the
    // item is re-solved with IEEE divisions so that the published value is exact in every case
    bool cs_done = false;
    if constexpr (GLOBAL) {
      if (L.scheme == 1) {
        if (hadi_solve_item_cs<NT, false>(L, it, w, cs, feed, tid)) hadi_solve_item_cs<NT, true>(L, it, w, cs, feed, tid);
        cs_done = true;
      }
    }
    if (!cs_done) {
#ifdef HADI_FORCE_EXACT
      if (!whole_exact) hadi_solve_item<NT, M1, M2, true>(L, it, w, feed, tid, tacc, tlast, sg.n0, sg.n1, hin);
#else
      if (!whole_exact) {
#ifdef HADI_NO_RERUN   /* timing experiments only */
        whole_exact = hadi_solve_item<NT, M1, M2, false>(L, it, w, feed, tid, tacc, tlast, sg.n0, sg.n1, hin) && L.n_items < 0;
#else
        whole_exact = hadi_solve_item<NT, M1, M2, false>(L, it, w, feed, tid, tacc, tlast, sg.n0, sg.n1, hin);
#endif
      }
#endif
      if (L.dbg_step == -7) whole_exact = true;   // test hook (HADI_DEBUG_STOP=-7:0): treat every fast pass as out of range
      if (whole_exact) {
        if (tid == 0 && L.prof != nullptr) atomicAdd((unsigned long long*)&L.prof[(size_t)gridDim.x * 8], 1ULL);
        // only the CTA that holds the last steps publishes: it re-solves the whole item
        if (sg.hout < 0) hadi_solve_item<NT, M1, M2, true>(L, it, w, feed, tid, tacc, tlast, 1, it.N, nullptr);
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
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        const int st = whole_exact ? HADI_HAND_BAD : HADI_HAND_READY;
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(L.hand_state + sg.hout), "r"(st) : "memory");
      }
      __syncthreads();
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
    __syncthreads();  // everyone is done with s_item, U and the tables before the next item
    HADI_TICK(6)
  }
#ifdef HADI_PHASE_TIMING
  if constexpr (FEED == 1) tacc[6] = feed.wait_cycles;   // slot 6 reports the ring wait of solver thread 0
  if (tid == 0 && L.prof != nullptr)
    for (int k = 0; k < 8; ++k) L.prof[(size_t)blockIdx.x * 8 + k] = tacc[k];
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
    if (L.scheme == 1) {
      hadi_cs_predict(it, w, cs, e0, e1, gtid, gnt);
      hadi_csync();
      hadi_phase_solve_a1<0, 0, EXACT>(it, w, e0, e1, 2 * n - 1, tid, NT, feed, bad, nullptr, 2 * it.N);
      hadi_csync();
      hadi_cs_rhs2(it, w, cs, e0, e1, gtid, gnt);
      hadi_csync();
      hadi_phase_solve_a2<0, 0, EXACT>(it, w, tid, NT, bad);
      hadi_csync();
      hadi_cs_correct(it, w, cs, e0, e1, gtid, gnt);
      hadi_csync();
      hadi_phase_solve_a1<0, 0, EXACT>(it, w, e0, e1, 2 * n, tid, NT, feed, bad, nullptr, 2 * it.N);
      hadi_csync();
      hadi_cs_rhs2(it, w, cs, e0, e1, gtid, gnt);
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
  w.ts_off = rank * TS_COUNT * w.n2;   // (8 x 18 x n2 doubles: far inside Y for any grid the kernel is chosen for)
  const HadiSmemLayout lay = hadi_smem_layout(m1, m2, w.ld, w.n1, w.n2, w.pj, false, true);
  char* sbase = reinterpret_cast<char*>(smem);
  double* scratch = L.scratch + (size_t)cid * L.scratch_stride;
  const HadiScratchLayout gl = hadi_scratch_layout(m1, m2, w.ld, w.pj, true, L.scheme == 1);
  double* Ualloc = scratch + gl.U;
  w.U = Ualloc + HADI_HALO * w.ld + 1;
  w.Y = scratch + gl.Y;
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
      if (gtid == 0 && L.prof != nullptr) atomicAdd((unsigned long long*)&L.prof[(size_t)gridDim.x * 8], 1ULL);
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
#define HADI_FEED0 0
#endif
#ifndef HADI_FEED1
#define HADI_FEED1 3
#endif
#define HADI_VARIANTS(X)                 \
  X(0, 320, 2, 100, 50, HADI_FEED0, false) \
  X(1, 256, 3, 50, 25, HADI_FEED1, false)  \
  X(2, 416, 2, 0, 0, 0, false)          \
  X(3, 1024, 1, 0, 0, 0, false)         \
  X(4, 320, 2, 100, 50, 4, false)       \
  X(5, 512, 1, 0, 0, 1, true)           \
  X(6, 1024, 1, 0, 0, 1, true)

struct VariantInfo {
  int threads, m1, m2;
  int feed;   // 0 plain loads, 1 TMA ring, 3 plain loads behind L1 prefetches, 4 co-operative warps
  bool global_state;
  const void* fn;
};
const VariantInfo* variants() {
  static const VariantInfo v[] = {
#define X(id, nt, minb, a, b, r, g) {nt, a, b, r, g, (const void*)hadi_douglas_kernel<nt, minb, a, b, r, g>},
      HADI_VARIANTS(X)
#undef X
  };
  return v;
}
constexpr int kNumVariants = 7;
constexpr int kClusterVariant = 7;   // hadi_cluster_kernel: one solve per thread-block cluster
constexpr int kClusterThreads = 256; // few threads, many registers: the generic phases spill badly at 64 registers

}  // namespace

int hadi_douglas_plan(int device, int m1, int m2, int ld, int n1, int n2, int pj, bool need_global, HadiPlan* plan,
                      bool want_cluster) {
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
  for (int k = 0; k < kNumVariants; ++k) {
    if (force && atoi(force) != k) continue;
    if (need_global && !v[k].global_state) continue;
    if (v[k].m1 != 0 && (v[k].m1 != m1 || v[k].m2 != m2)) continue;
    if (v[k].threads < m1 + 1 || v[k].threads - 1 <= m2) continue;
    if (v[k].feed == 1 && v[k].threads <= 32 * ((m2 + 1 + 31) / 32)) continue;   // needs a producer thread past the solver warps
    smem = hadi_smem_layout(m1, m2, ld, n1, n2, pj, v[k].feed == 1, v[k].global_state, v[k].feed == 4, v[k].m1 != 0).total;
    if (smem > (size_t)max_smem) continue;
    pick = k;
    break;
  }
  if (pick < 0) return -1;
  e = cudaFuncSetAttribute(v[pick].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v[pick].fn, v[pick].threads, smem);
  if (e != cudaSuccess) return (int)e;
  if (occ < 1) return -1;
  plan->global_state = v[pick].global_state;
  plan->variant = pick;
  plan->threads = v[pick].threads;
  plan->ctas_per_sm = occ;
  plan->sm_count = sms;
  plan->smem_bytes = smem;
  plan->cluster = 1;
  return 0;
}

int hadi_launch_douglas(const HadiLaunch& L, const HadiPlan& plan, int grid_ctas, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  // the run-time-dimension kernels serve many grid shapes and plans are cached by the host layer: the dynamic
  // shared-memory limit of the function must be the one of THIS plan, not of the plan made last
  {
    const void* fn = plan.variant == kClusterVariant ? (const void*)hadi_cluster_kernel<kClusterThreads>
                                                     : variants()[plan.variant].fn;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem_bytes);
    if (e != cudaSuccess) return (int)e;
  }
  switch (plan.variant) {
#define X(id, nt, minb, a, b, r, g) \
  case id:                          \
    hadi_douglas_kernel<nt, minb, a, b, r, g><<<grid_ctas, nt, plan.smem_bytes, st>>>(L); \
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
