// hadi — launch descriptor shared by the host layer (hadi_host.cpp) and the kernel (hadi_kernel.cu).
#pragma once
#include <cstddef>
#include "hadi_phases.cuh"

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
  double* out_values;        // [n_items] price at (S0, V0) per item (slot item.out)
  double* out_U;             // optional [n_items][P] natural layout
  double* out_lam;           // optional [n_items][P]
};

// shared memory (bytes) the Douglas kernel needs for a grid
inline size_t hadi_smem_bytes(int m1, int m2, int ld, int n1, int n2) {
  (void)m1;
  size_t d = (size_t)2 * (size_t)(m2 + 1) * (size_t)ld + (size_t)TI_COUNT * (size_t)n1 + (size_t)TJ_COUNT * (size_t)n2;
  return d * sizeof(double) + (size_t)n1 * sizeof(int) + 16;
}
inline size_t hadi_scratch_doubles(int m1, int m2, int ld, int pj) {
  return (size_t)2 * (size_t)(m1 + 1) * (size_t)pj + (size_t)(m2 + 1) * (size_t)ld;
}

// defined in hadi_kernel.cu
int hadi_launch_douglas(const HadiLaunch& L, int grid_ctas, size_t smem_bytes, void* stream);
int hadi_douglas_config(int* threads, int* max_smem_optin, int* sm_count, int device);
