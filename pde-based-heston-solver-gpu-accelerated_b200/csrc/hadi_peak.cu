// hadi — FP64 issue-rate micro-benchmark (roofline denominator).
//
// MEASURED_PEAKS.json carries HBM and bf16 peaks only; the SMEM-resident Douglas kernel is bound by
// the FP64 pipe.  Parity forbids FMA contraction, so every algorithmic flop is one DADD or DMUL issue:
// the honest roof is the un-fused DADD/DMUL issue rate.  This file measures it (and the DFMA rate and
// the dependent-issue latency for context) with a register-only kernel, timed with CUDA events.
#include <cuda_runtime.h>

#include "../../include/hadi.h"

namespace {

// ILP independent chains per thread, un-fused mul/add pairs (mode 0), fma (mode 1), or one fully
// dependent add chain (mode 2, latency).
template <int MODE, int ILP>
__global__ void __launch_bounds__(256) peak_kernel(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = a + k + threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      if (MODE == 0) {
        x[k] = __dmul_rn(x[k], a);
        x[k] = __dadd_rn(x[k], b);
      } else if (MODE == 1) {
        x[k] = __fma_rn(x[k], a, b);
      } else {
        x[k] = __dadd_rn(x[k], b);
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k];
  if (s == 123.456) out[0] = s;  // keep the chains alive
}

template <int MODE, int ILP>
float run(int blocks, int threads, int iters, cudaStream_t st) {
  double* d = nullptr;
  cudaMalloc(&d, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, st);
    peak_kernel<MODE, ILP><<<blocks, threads, 0, st>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}

}  // namespace

extern "C" int hadi_measure_fp64(int device, double* unfused_tflops, double* fma_tflops, double* dep_latency_ns) {
  if (cudaSetDevice(device) != cudaSuccess) return HADI_ERR_CUDA;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return HADI_ERR_CUDA;
  const int threads = 256, blocks = sms * 8, iters = 4096;
  constexpr int ILP = 8;
  const float ms0 = run<0, ILP>(blocks, threads, iters, 0);
  const float ms1 = run<1, ILP>(blocks, threads, iters, 0);
  const float ms2 = run<2, 1>(1, 32, 1 << 16, 0);
  if (cudaGetLastError() != cudaSuccess) return HADI_ERR_CUDA;
  const double n = (double)blocks * threads * (double)iters * ILP;
  if (unfused_tflops) *unfused_tflops = 2.0 * n / (ms0 * 1e-3) / 1e12;   // one DMUL + one DADD = 2 flops, 2 issues
  if (fma_tflops) *fma_tflops = 2.0 * n / (ms1 * 1e-3) / 1e12;           // one DFMA = 2 flops, 1 issue
  if (dep_latency_ns) *dep_latency_ns = (ms2 * 1e6) / (double)(1 << 16);  // ns per dependent DADD
  return HADI_OK;
}
