// hadi — fused, persistent Douglas ADI kernel for sm_100a (B200).
//
// One CTA solves one work item (an option or one Jacobian bump) from payoff to price without
// leaving the SM: the solution U and the stage vector Y stay in shared memory for all N time steps;
// the explicit A0/A1/A2 products, both implicit line solves, the boundary terms, the dividend jump
// and the American projection are fused between __syncthreads().  CTAs are persistent and pull
// items from a global counter (longest items first), so a batch of any size runs as one launch.
//
// Replaces the reference's "Base_Price_computation" / "Jacobian_computation" Kokkos kernels
// (src/jacobian_computation.cpp:232,391,...; src/heston_calibration.cpp:2206,2366) together with
// the 12 global-memory work arrays of DO_Workspace (src/DO_solver_workspace.hpp:5-44).
//
// Compile with -fmad=false: parity with the reference requires un-fused multiplies and adds.
#include <cuda_runtime.h>
#include <cstdio>

#include "hadi_launch.h"

#ifndef HADI_NT
#define HADI_NT 384
#endif

namespace {

template <int NT>
__global__ void __launch_bounds__(NT, 2) hadi_douglas_kernel(const HadiLaunch L) {
  extern __shared__ double smem[];
  __shared__ int s_item;
  const int tid = threadIdx.x;

  HadiView w;
  w.m1 = L.m1; w.m2 = L.m2; w.ld = L.ld; w.P = (L.m1 + 1) * (L.m2 + 1);
  w.n1 = L.n1; w.n2 = L.n2; w.pj = L.pj;
  const int rows = L.m2 + 1;
  w.U = smem;
  w.Y = w.U + rows * L.ld;
  w.ti = w.Y + rows * L.ld;
  w.tj = w.ti + TI_COUNT * L.n1;
  w.divk = reinterpret_cast<int*>(w.tj + TJ_COUNT * L.n2);
  double* scratch = L.scratch + (size_t)blockIdx.x * L.scratch_stride;
  w.fM = scratch;
  w.fT = w.fM + (size_t)(L.m1 + 1) * L.pj;
  w.lam = w.fT + (size_t)(L.m1 + 1) * L.pj;

  for (;;) {
    if (tid == 0) s_item = atomicAdd(L.counter, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= L.n_items) break;
    const HadiItem it = L.items[item];
    const double* sg = L.s_pool + it.s_off;
    const double* vg = L.v_pool + it.v_off;
    const double* eg = L.e_pool + it.e_off;
    w.c = it.theta * it.dt;

    hadi_phase_tables(it, w, sg, vg, tid, NT);
    __syncthreads();
    hadi_phase_factor(it, w, vg, tid, NT, NT - 1);
    // initial condition U = payoff (the reference's U_0 input array), lambda = 0
    {
      const double* pay = hadi_ti(w, TI_PAY);
      for (int p = tid; p < rows * (L.m1 + 1); p += NT) {
        const int j = p / (L.m1 + 1), i = p - j * (L.m1 + 1);
        w.U[j * L.ld + i] = pay[i];
        if (it.style == 1) w.lam[j * L.ld + i] = 0.0;
      }
    }
    __syncthreads();

    int div_cur = 0;
    for (int n = 1; n <= it.N; ++n) {
      if (it.nd > 0) {
        const int hit = hadi_dividend_at(n, it.dt, it.nd, L.div_dates, div_cur);
        if (hit >= 0) {  // uniform across the CTA
          hadi_phase_div1(w, L.div_amounts[hit], L.div_pcts[hit], tid, NT);
          __syncthreads();
          hadi_phase_div2(w, tid, NT);
          __syncthreads();
        }
      }
      const double e0 = eg[n - 1], e1 = eg[n];
      hadi_phase_explicit(it, w, e0, e1, tid, NT);
      __syncthreads();
      hadi_phase_solve_a1(it, w, tid, NT);
      __syncthreads();
      hadi_phase_solve_a2(it, w, e0, e1, tid, NT);
      __syncthreads();
      if (it.style == 1) {
        hadi_phase_project(it, w, tid, NT);
        __syncthreads();
      }
    }

    if (tid == 0) L.out_values[it.out] = w.U[it.idx_v * L.ld + it.idx_s];
    if (L.out_U != nullptr) {
      double* dst = L.out_U + (size_t)it.out * w.P;
      for (int p = tid; p < w.P; p += NT) {
        const int j = p / (L.m1 + 1), i = p - j * (L.m1 + 1);
        dst[p] = w.U[j * L.ld + i];
      }
    }
    if (L.out_lam != nullptr && it.style == 1) {
      double* dst = L.out_lam + (size_t)it.out * w.P;
      for (int p = tid; p < w.P; p += NT) {
        const int j = p / (L.m1 + 1), i = p - j * (L.m1 + 1);
        dst[p] = w.lam[j * L.ld + i];
      }
    }
    __syncthreads();  // everyone is done with s_item, U and the tables before the next item
  }
}

}  // namespace

int hadi_douglas_config(int* threads, int* max_smem_optin, int* sm_count, int device) {
  int v = 0;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return -1;
  *max_smem_optin = v;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return -1;
  *sm_count = v;
  *threads = HADI_NT;
  return 0;
}

int hadi_launch_douglas(const HadiLaunch& L, int grid_ctas, size_t smem_bytes, void* stream) {
  cudaError_t e = cudaFuncSetAttribute(hadi_douglas_kernel<HADI_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem_bytes);
  if (e != cudaSuccess) return (int)e;
  hadi_douglas_kernel<HADI_NT><<<grid_ctas, HADI_NT, smem_bytes, (cudaStream_t)stream>>>(L);
  return (int)cudaGetLastError();
}
