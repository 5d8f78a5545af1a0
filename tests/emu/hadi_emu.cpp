// TEST INFRASTRUCTURE — CPU emulation of the CUDA kernel's phase sequence.
//
// Compiles the product's device header (csrc/hadi_phases.cuh) with g++ and executes the phases of
// hadi_douglas_kernel (csrc/hadi_kernel.cu) with a serial loop over thread ids in place of the CTA's
// threads and __syncthreads().  It lets the CPU test-suite check the kernel's index logic and
// operation order bit-for-bit against the oracle without a GPU.  Not shipped, not a fallback.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "hadi_phases.cuh"
#include "hadi_phases_cs.cuh"

extern "C" int hadi_emu_solve(const HadiItem* item, int m1, int m2, int nt, const double* sg, const double* vg,
                              const double* eg, int nd, const double* dd, const double* da, const double* dp,
                              double* price, double* U_out, double* lam_out) {
  HadiItem it = *item;
  HadiView w;
  w.m1 = m1; w.m2 = m2; w.ld = hadi_geo_ld(m1); w.P = (m1 + 1) * (m2 + 1);
  w.n1 = hadi_geo_n1(m1); w.n2 = hadi_geo_n2(m2); w.pj = hadi_geo_pj(m2);
  const int rows = m2 + 1;
  if (nt < m1 + 1 || nt - 1 <= m2) return -1;
  std::vector<double> U((rows + 2 * HADI_HALO) * w.ld + 2, 0.0), Y(rows * w.ld, 1e300), ti(TI_COUNT * w.n1, 1e300), tj(TJ_COUNT * w.n2, 1e300);
  std::vector<double> fM(m1 * w.pj, 1e300), fB(m1 * 2 * w.pj, 1e300), lam(rows * w.ld, 1e300);
  std::vector<int> divk(w.n1, -7);
  w.U = U.data() + HADI_HALO * w.ld + 1; w.Y = Y.data(); w.ti = ti.data(); w.tj = tj.data(); w.divk = divk.data();
  w.fM = fM.data(); w.fB = fB.data(); w.lam = lam.data();
  HadiDirectFeed feed;
  feed.fM = w.fM; feed.fB = w.fB; feed.pj = w.pj;
  w.c = it.theta * it.dt;
  const double rdt = hadi_rcp_prep(it.dt);
  const bool spec = true;
  unsigned bad = 0;
#define PHASE(call) for (int tid = 0; tid < nt; ++tid) { call; }
  PHASE(hadi_phase_tables(it, w, sg, vg, tid, nt));
  PHASE(hadi_phase_factor(it, w, vg, tid, nt, nt - 1));
  {
    for (int tid = 0; tid < nt; ++tid) {
      const HadiMap mp = hadi_map(m1, m2, tid, nt);
      if (!mp.active) continue;
      const double pay = hadi_ti(w, TI_PAY)[mp.i];
      for (int j = mp.j0; j < mp.j1; ++j) {
        w.U[j * w.ld + mp.i] = pay;
        if (it.style == 1) { w.lam[j * w.ld + mp.i] = 0.0; w.Y[j * w.ld + mp.i] = 0.0; }
      }
    }
  }
  int div_cur = 0;
  for (int n = 1; n <= it.N; ++n) {
    if (it.nd > 0) {
      const int hit = hadi_dividend_at(n, it.dt, nd, dd, div_cur);
      if (hit >= 0) {
        PHASE(hadi_phase_div1(w, da[hit], dp[hit], tid, nt));
        PHASE(hadi_phase_div2(w, tid, nt));
        if (it.style == 1) PHASE(hadi_phase_div3(w, tid, nt));
      }
    }
    const double e0 = eg[n - 1], e1 = eg[n];
    if (spec && m1 == 100 && m2 == 50) {
      PHASE((hadi_phase_explicit<100, 50>(it, w, e0, e1, tid, nt)));
      PHASE((hadi_phase_solve_a1<100, 50, true>(it, w, e0, e1, n, tid, nt, feed, bad)));
      PHASE((hadi_phase_rhs2<100, 50>(it, w, e0, e1, tid, nt)));
      PHASE((hadi_phase_solve_a2<100, 50, true>(it, w, tid, nt, bad)));
      if (it.style == 1) PHASE((hadi_phase_project<100, 50, true>(it, w, rdt, tid, nt, bad)));
    } else {
      PHASE((hadi_phase_explicit<0, 0>(it, w, e0, e1, tid, nt)));
      PHASE((hadi_phase_solve_a1<0, 0, true>(it, w, e0, e1, n, tid, nt, feed, bad)));
      PHASE((hadi_phase_rhs2<0, 0>(it, w, e0, e1, tid, nt)));
      PHASE((hadi_phase_solve_a2<0, 0, true>(it, w, tid, nt, bad)));
      if (it.style == 1) PHASE((hadi_phase_project<0, 0, true>(it, w, rdt, tid, nt, bad)));
    }
  }
  *price = w.U[it.idx_v * w.ld + it.idx_s];
  for (int p = 0; p < w.P; ++p) {
    const int j = p / (m1 + 1), i = p - j * (m1 + 1);
    if (U_out) U_out[p] = w.U[j * w.ld + i];
    if (lam_out) lam_out[p] = (it.style == 1) ? w.lam[j * w.ld + i] : 0.0;
  }
  return 0;
}
// Craig-Sneyd: the phase sequence of hadi_solve_item_cs (csrc/hadi_kernel.cu)
extern "C" int hadi_emu_solve_cs(const HadiItem* item, int m1, int m2, int nt, const double* sg, const double* vg,
                                 const double* eg, double* price, double* U_out) {
  HadiItem it = *item;
  HadiView w;
  w.m1 = m1; w.m2 = m2; w.ld = hadi_geo_ld(m1); w.P = (m1 + 1) * (m2 + 1);
  w.n1 = hadi_geo_n1(m1); w.n2 = hadi_geo_n2(m2); w.pj = hadi_geo_pj(m2);
  const int rows = m2 + 1;
  if (nt < m1 + 1 || nt - 1 <= m2) return -1;
  std::vector<double> U((rows + 2 * HADI_HALO) * w.ld + 2, 0.0), Y(rows * w.ld, 1e300), ti(TI_COUNT * w.n1, 1e300), tj(TJ_COUNT * w.n2, 1e300);
  std::vector<double> fM(m1 * w.pj, 1e300), fB(m1 * 2 * w.pj, 1e300), lam(rows * w.ld, 1e300);
  std::vector<double> Y0(rows * w.ld, 1e300), R0(rows * w.ld, 1e300), R1(rows * w.ld, 1e300), R2(rows * w.ld, 1e300);
  std::vector<int> divk(w.n1, -7);
  w.U = U.data() + HADI_HALO * w.ld + 1; w.Y = Y.data(); w.ti = ti.data(); w.tj = tj.data(); w.divk = divk.data();
  w.fM = fM.data(); w.fB = fB.data(); w.lam = lam.data();
  HadiCsView cs{Y0.data(), R0.data(), R1.data(), R2.data()};
  HadiDirectFeed feed;
  feed.fM = w.fM; feed.fB = w.fB; feed.pj = w.pj;
  w.c = it.theta * it.dt;
  unsigned bad = 0;
  PHASE(hadi_phase_tables(it, w, sg, vg, tid, nt));
  PHASE(hadi_phase_factor(it, w, vg, tid, nt, nt - 1));
  for (int tid = 0; tid < nt; ++tid) {
    const HadiMap mp = hadi_map(m1, m2, tid, nt);
    if (!mp.active) continue;
    for (int j = mp.j0; j < mp.j1; ++j) w.U[j * w.ld + mp.i] = hadi_ti(w, TI_PAY)[mp.i];
  }
  for (int n = 1; n <= it.N; ++n) {
    const double e0 = eg[n - 1], e1 = eg[n];
    PHASE(hadi_cs_predict(it, w, cs, e0, e1, tid, nt));
    PHASE((hadi_phase_solve_a1<0, 0, true>(it, w, e0, e1, 2 * n - 1, tid, nt, feed, bad, nullptr, 2 * it.N)));
    PHASE(hadi_cs_rhs2(it, w, cs, e0, e1, tid, nt));
    PHASE((hadi_phase_solve_a2<0, 0, true>(it, w, tid, nt, bad)));
    PHASE(hadi_cs_correct(it, w, cs, e0, e1, tid, nt));
    PHASE((hadi_phase_solve_a1<0, 0, true>(it, w, e0, e1, 2 * n, tid, nt, feed, bad, nullptr, 2 * it.N)));
    PHASE(hadi_cs_rhs2(it, w, cs, e0, e1, tid, nt));
    PHASE((hadi_phase_solve_a2<0, 0, true>(it, w, tid, nt, bad)));
  }
  *price = w.U[it.idx_v * w.ld + it.idx_s];
  for (int p = 0; p < w.P; ++p) {
    const int j = p / (m1 + 1), i = p - j * (m1 + 1);
    if (U_out) U_out[p] = w.U[j * w.ld + i];
  }
  return 0;
}
extern "C" int hadi_emu_item_size() { return (int)sizeof(HadiItem); }
