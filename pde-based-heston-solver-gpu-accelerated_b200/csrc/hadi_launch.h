// hadi — launch descriptor shared by the host layer (hadi_host.cpp) and the kernel (hadi_kernel.cu).
#pragma once
#include <cstddef>
#include "hadi_phases.cuh"
#include "hadi_phases_fast.cuh"
#ifndef HADI_LEAN
#define HADI_LEAN 1   /* grid-specialised variants keep only the TI_CORE per-column tables in shared memory */
#endif

// One entry of a CTA's static work list (split schedule): time steps n0..n1 (1-based, inclusive) of items[item].
// A solve cut in two hands its state (U, lambda) from the CTA that ran steps 1..k to the CTA that runs
// k+1..N through hand-off slot `hout` of the first == `hin` of the second (-1: none).
struct HadiSegment {
  int item, n0, n1, hin, hout, pad0, pad1, pad2;
};
#define HADI_FLAGS_DEFAULT 1
#define HADI_HAND_PENDING 0
#define HADI_HAND_READY 1
#define HADI_HAND_BAD 2      /* a guarded division left its range: the final segment re-solves the whole item with IEEE '/' */

struct HadiLaunch {
  int m1, m2, ld, n1, n2, pj;
  int n_items;
  const HadiItem* items;     // [n_items], device
  const double* s_pool;      // device
  const double* v_pool;      // device
  const double* e_pool;      // device
  int nd;
  const double* div_dates;   // device, [nd]
  const double* div_amounts;
  const double* div_pcts;
  double* scratch;           // device, per-CTA-slot scratch (A1 factors, lambda)
  size_t scratch_stride;     // doubles per CTA slot
  int* counter;              // device work counter (zeroed before launch)
  double* out_values;        // [n_items][out_stride] price at (S0, V0) per item (slot item.out)
  int out_stride;            // values per item: 1, or 3 = {price, U(S0, v_lower), U(S0, v_upper)} (interpolated-V0 Jacobian)
  double* out_U;             // optional [n_items][P] natural layout
  double* out_lam;           // optional [n_items][P]
  long long* prof;           // optional [work slots][8] phase cycle counters (HADI_PHASE_TIMING builds only)
  unsigned long long* reruns;  // items of this launch re-solved with IEEE divisions (guarded division left its range,
                             // or a hand-off timed out); zeroed before the launch, read back with the values
  int dbg_step, dbg_phase;   // HADI_DEBUG_STOP builds only: end every item after phase dbg_phase of step dbg_step
  int scheme;                // 0 Douglas, 1 Craig-Sneyd (global-state kernel only)
  // split schedule (nullptr: CTAs pull whole items from `counter`): CTA b runs segs[seg_off[b] .. seg_off[b+1])
  const HadiSegment* segs;
  const int* seg_off;        // [gridDim.x + 1]
  int* hand_state;           // [n_hand] HADI_HAND_* (zeroed before launch)
  double* hand_data;         // [n_hand][2 * P]: U then lambda, natural layout
  int flags;                 // scheduling switches (HADI_FLAGS, default HADI_FLAGS_DEFAULT): bit 0 — in the duo kernel the two
                             // teams of a CTA take turns in the FP64-bound explicit stage; bit 1 — and in the projection
  int vgrid;                 // work slots of this launch (set by hadi_launch_douglas; = CTAs except in the duo kernel)
};

// Kernel variant chosen for a grid shape (hadi_kernel.cu).
struct HadiPlan {
  bool global_state;  // U and Y live in L2-resident global scratch (grids beyond shared memory, Craig-Sneyd)
  int variant;        // index of the template instantiation
  int threads;        // CTA size
  int ctas_per_sm;    // resident CTAs per SM at this shared-memory footprint
  int sm_count;
  size_t smem_bytes;  // dynamic shared memory per CTA
  int cluster;        // CTAs that share one solve (1, or HADI_CLUSTER in the cluster kernel)
  int duo = 1;        // solves in flight per CTA (2 in the duo kernel; ctas_per_sm counts work slots)
};

// shared memory layout of the Douglas kernel (offsets in bytes from the dynamic smem base)
struct HadiSmemLayout {
  size_t U, Y, ti, tj, tjp, divk, ring, bars, total;
};
HADI_HD HadiSmemLayout hadi_smem_layout(int m1, int m2, int ld, int n1, int n2, int pj, bool ring,
                                        bool global_state = false, bool coop = false, bool lean = false,
                                        bool tmem = false) {
  (void)m1;
  HadiSmemLayout s;
  size_t off = 0;
  // U carries HADI_HALO zero rows above and below and one spare word at either end
  s.U = off; if (!global_state) off += sizeof(double) * ((size_t)(m2 + 1 + 2 * HADI_HALO) * ld + 2);
  s.Y = off; if (!global_state) off += sizeof(double) * (size_t)(m2 + 1) * ld;
  s.ti = off; off += sizeof(double) * (size_t)((lean && HADI_LEAN) ? TI_CORE : TI_COUNT) * n1;   // lean tables in the grid-specialised variants
  s.tj = off; off += sizeof(double) * (size_t)TJ_COUNT * n2;
  // packed per-row records {L2, L1, D0, U1, U2, F, G, MM} of the fused R + S2 phase (grid-specialised variants)
  off = (off + 15) & ~size_t(15);
  s.tjp = off; if (lean) off += sizeof(double) * (size_t)8 * n2;
  s.divk = off; off += sizeof(int) * (size_t)n1;
  off = (off + 127) & ~size_t(127);
  s.ring = off;
  if (ring) off += sizeof(double) * (size_t)HADI_NS * HADI_KF * pj;
  if (coop) off += (size_t)hadi_co_stage_bytes(m2);
  if (tmem) off += sizeof(double) * (size_t)(n1 + 4);   // dummy row of the half sweeps (hadi_tmem_solve_a1)
  s.bars = off;
  if (ring) off += sizeof(unsigned long long) * 2 * HADI_NS;
  s.total = off + 16;
  return s;
}
// per-CTA global scratch: fM [m1][pj], fB [m1][2*pj], lambda [m2+1][ld]; the global-state kernel adds
// U (with halo) and Y, and for Craig-Sneyd Y0, R0, R1, R2 — every array starts on a 128-byte boundary
#define HADI_CLUSTER 8
struct HadiScratchLayout {
  size_t fM, fB, lam, U, Y, Y0, R0, R1, R2, ts, mail, total;   // offsets in doubles
};
HADI_HD HadiScratchLayout hadi_scratch_layout(int m1, int m2, int ld, int pj, bool global_state, bool cs) {
  HadiScratchLayout s;
  size_t off = 0;
  const size_t arr = (size_t)(m2 + 1) * (size_t)ld;
  // sized for either factor layout: classic [m1][pj] or co-operative [13*ceil((m2+1)/13)][(m1+7)&~7]
  const size_t co = (size_t)(((m2 + 13) / 13) * 13) * (size_t)((m1 + 7) & ~7);
  const size_t fsz = ((size_t)m1 * pj > co) ? (size_t)m1 * pj : co;
  s.fM = off; off += fsz; off = (off + 15) & ~size_t(15);
  s.fB = off; off += 2 * fsz; off = (off + 15) & ~size_t(15);
  s.lam = off; off += arr; off = (off + 15) & ~size_t(15);
  s.U = off; if (global_state) off += (size_t)(m2 + 1 + 2 * HADI_HALO) * ld + 2; off = (off + 15) & ~size_t(15);
  s.Y = off; if (global_state) off += arr; off = (off + 15) & ~size_t(15);
  s.Y0 = off; if (cs) off += arr; off = (off + 15) & ~size_t(15);
  s.R0 = off; if (cs) off += arr; off = (off + 15) & ~size_t(15);
  s.R1 = off; if (cs) off += arr; off = (off + 15) & ~size_t(15);
  s.R2 = off; if (cs) off += arr; off = (off + 15) & ~size_t(15);
  // cluster kernel: one A2 assembly scratch region per CTA (TS_COUNT rows of n2 doubles).  Round 1 kept them inside Y,
  // which they overran on grids below about 35 s-nodes (into the mailbox behind the arrays: garbage results).
  s.ts = off; if (global_state) off += (size_t)HADI_CLUSTER * TS_COUNT * hadi_geo_n2(m2); off = (off + 15) & ~size_t(15);
  s.mail = off; off += 16;   // cluster kernel: work-item mailbox and vote word
  s.total = (off + 31) & ~size_t(31);
  return s;
}

// defined in hadi_kernel.cu; return 0 or a cudaError_t
int hadi_douglas_plan(int device, int m1, int m2, int ld, int n1, int n2, int pj, bool need_global, HadiPlan* plan,
                      bool want_cluster = false, bool many = false);
int hadi_launch_douglas(const HadiLaunch& L, const HadiPlan& plan, int grid_ctas, void* stream);

// defined in hadi_wide.cu: one solve spread over a team of co-resident CTAs (co-operative launch); HadiPlan::cluster
// carries the team size, set per batch with hadi_wide_team().  HadiLaunch::counter must hold HADI_COUNTER_INTS zeroed ints.
#define HADI_WIDE_VARIANT 9
#define HADI_WIDE_MAX_ITEMS_DEFAULT 40   /* batches up to this size of global-state solves go to the wide kernel (twice as many on small grids) */
#define HADI_COUNTER_INTS 2048   /* work counter, cluster mailboxes, one barrier word per team (16 + 8 * team) */
int hadi_wide_plan(int device, int m1, int m2, int ld, int n1, int n2, int pj, HadiPlan* plan);
int hadi_wide_team(int n_items, int sm_count, int m1, int m2);
int hadi_launch_wide(const HadiLaunch& L, const HadiPlan& plan, int grid_ctas, void* stream);
