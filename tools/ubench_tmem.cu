// Development aid: tensor memory (TMEM) as a per-lane scratchpad for FP64 streams on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench_tmem ubench_tmem.cu
// Checks (1) that two CTAs per SM can hold 256 columns each, (2) the lane mapping of the 32x32b shape
// (warp w reaches lanes 32*(w%4)..+31, also from another warp of the same quarter after a CTA barrier),
// (3) the latency of a dependent tcgen05.ld and (4) a Thomas back-substitution chain fed from TMEM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ void tm_alloc(uint32_t* slot, int cols) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(slot);
  if (cols == 256) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(a) : "memory");
  else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(a) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_free(uint32_t addr, int cols) {
  if (cols == 256) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(addr) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void tm_st2(uint32_t taddr, double v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__double2loint(v)), "r"(__double2hiint(v)) : "memory");
}
__device__ __forceinline__ double tm_ld2(uint32_t taddr) {
  int lo, hi;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(taddr) : "memory");
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ void tm_ld8(uint32_t taddr, double (&v)[4]) {
  int r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = __hiloint2double(r[2 * k + 1], r[2 * k]);
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ double pattern(int cta, int q, int lane, int col) {
  return 1.0 + cta * 1e-3 + q * 0.25 + lane * 0.001953125 + col * 7.62939453125e-06;
}

// (1) + (2): every CTA fills its 256 columns from warps 0-3, reads them back from warps 4-7 and 0-3.
__global__ void __launch_bounds__(320, 2) k_check(int* bad, int* smids) {
  extern __shared__ double dyn[];
  __shared__ uint32_t s_addr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tm_alloc(&s_addr, 256);
  tm_fence_before();
  __syncthreads();
  tm_fence_after();
  const uint32_t base = s_addr;
  if (tid == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    smids[blockIdx.x] = (int)smid | ((int)(base & 0xffff) << 16);
  }
  const int q = warp & 3;
  const uint32_t mine = base + ((uint32_t)(32 * q) << 16);
  if (warp < 4) {
    for (int c = 0; c < 128; ++c) tm_st2(mine + 2 * c, pattern(blockIdx.x, q, lane, c));
    tm_wait_st();
  }
  tm_fence_before();
  __syncthreads();
  tm_fence_after();
  int nb = 0;
  if (warp < 8) {
    for (int c = 0; c < 128; ++c) {
      const double v = tm_ld2(mine + 2 * c);
      tm_wait_ld();
      if (v != pattern(blockIdx.x, q, lane, c)) nb++;
    }
    // x8 loads see the same data
    for (int c = 0; c < 128; c += 4) {
      double v[4];
      tm_ld8(mine + 2 * c, v);
      tm_wait_ld();
      for (int k = 0; k < 4; ++k) if (v[k] != pattern(blockIdx.x, q, lane, c + k)) nb++;
    }
  }
  if (nb) atomicAdd(bad, nb);
  dyn[tid] = nb;
  tm_fence_before();
  __syncthreads();
  if (warp == 0) tm_free(base, 256);
}

// (3) dependent tcgen05.ld latency: the loaded value (0.0) feeds the next address
__global__ void k_lat(long long* cyc, int iters, int* sink) {
  __shared__ uint32_t s_addr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tm_alloc(&s_addr, 128);
  tm_fence_before();
  __syncthreads();
  tm_fence_after();
  const uint32_t base = s_addr + ((uint32_t)(32 * (warp & 3)) << 16);
  for (int c = 0; c < 64; ++c) tm_st2(base + 2 * c, 0.0);
  tm_wait_st();
  uint32_t off = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const double v = tm_ld2(base + off);
    tm_wait_ld();
    off = (off + 2 + (uint32_t)__double2loint(v)) & 63;
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; *sink = (int)off; }
  // independent loads, one wait per 8
  const long long t2 = clock64();
  double acc = 0.0;
  for (int it = 0; it < iters; it += 8) {
    double v[4], w[4];
    tm_ld8(base + 0, v);
    tm_ld8(base + 8, w);
    tm_wait_ld();
    acc += v[0] + w[3];
  }
  const long long t3 = clock64();
  if (threadIdx.x == 0) { cyc[1] = t3 - t2; sink[1] = (int)acc; }
  tm_fence_before();
  __syncthreads();
  if (warp == 0) tm_free(s_addr, 128);
}

// (4) back substitution x_i = (y_i - u_i x_{i+1}) / t_i with (t, r) from TMEM: lane l < 16 holds the pivots of
// its row, lane l + 16 the prepared reciprocals; chunks of KB nodes, next chunk requested before the chain.
// MODE 0: operands from registers only (chain floor), 1: TMEM + shuffle, 2: + y from shared memory, x stored back
template <int MODE, int KB>
__global__ void __launch_bounds__(320, 2) k_chain(long long* cyc, double* out, int reps) {
  extern __shared__ double Y[];   // [52][103]
  __shared__ uint32_t s_addr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int k = tid; k < 52 * 103; k += blockDim.x) Y[k] = 1.0 + 1e-6 * k;
  if (warp == 0) tm_alloc(&s_addr, 256);
  tm_fence_before();
  __syncthreads();
  tm_fence_after();
  const uint32_t mine = s_addr + ((uint32_t)(32 * (warp & 3)) << 16);
  if (warp < 4) {
    for (int c = 0; c < 100; ++c) tm_st2(mine + 2 * c, lane < 16 ? 1.25 + 1e-3 * c : 1.0 / (1.25 + 1e-3 * c));
    tm_wait_st();
  }
  __syncthreads();
  if (warp < 4) {
  const int row = 13 * warp + (lane & 15);
  double* y = Y + (row < 52 ? row : 51) * 103;
  double xn = 0.5;
  const long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    double cur[KB];
    if (MODE >= 1) {
      if (KB == 4) { double v[4]; tm_ld8(mine, v); for (int k = 0; k < KB; ++k) cur[k] = v[k]; }
      else for (int k = 0; k < KB; ++k) cur[k] = tm_ld2(mine + 2 * k);
      tm_wait_ld();
    }
#pragma unroll 1
    for (int c = 0; c < 100 / KB; ++c) {
      double nxt[KB];
      if (MODE >= 1) {
        const int cn = (c + 1 < 100 / KB) ? c + 1 : c;
        if (KB == 4) { double v[4]; tm_ld8(mine + 2 * KB * cn, v); for (int k = 0; k < KB; ++k) nxt[k] = v[k]; }
        else for (int k = 0; k < KB; ++k) nxt[k] = tm_ld2(mine + 2 * (KB * cn + k));
      }
      double tt[KB], rr[KB], yy[KB];
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        if (MODE >= 1) {
          tt[k] = cur[k];
          rr[k] = __shfl_sync(0xffffffffu, cur[k], (lane & 15) + 16);
        } else {
          tt[k] = 1.25 + 1e-3 * k; rr[k] = 0.8;
        }
        yy[k] = (MODE >= 2) ? y[100 - (c * KB + k)] : 1.0;
      }
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        const double a = yy[k] - 0.125 * xn;
        const double q0 = __dmul_rn(a, rr[k]);
        const double r = __fma_rn(-tt[k], q0, a);
        xn = __fma_rn(rr[k], r, q0);
        if (MODE >= 2) y[100 - (c * KB + k)] = xn;
      }
      if (MODE >= 1) {
        tm_wait_ld();
#pragma unroll
        for (int k = 0; k < KB; ++k) cur[k] = nxt[k];
      }
    }
  }
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  if (xn == 123.456) out[0] = xn;
  }
  tm_fence_before();
  __syncthreads();
  if (warp == 0) tm_free(s_addr, 256);
}

template <int MODE, int KB>
void run_chain(const char* name, int blocks) {
  long long* c; double* o;
  cudaMalloc(&c, 8 * blocks); cudaMalloc(&o, 8);
  const int reps = 50;
  const size_t smem = 100 * 1024;
  cudaFuncSetAttribute(k_chain<MODE, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_chain<MODE, KB><<<blocks, 320, smem>>>(c, o, reps);
  k_chain<MODE, KB><<<blocks, 320, smem>>>(c, o, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("%-44s blocks %3d: %.1f cycles per node (%s)\n", name, blocks, (double)h / (reps * 100.0), cudaGetErrorString(e));
  cudaFree(c); cudaFree(o);
}

int main() {
  int* bad; int* smids;
  const int nb = 296;
  cudaMalloc(&bad, 4); cudaMalloc(&smids, 4 * nb);
  cudaMemset(bad, 0, 4);
  const size_t smem = 100 * 1024;
  cudaFuncSetAttribute(k_check, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_check, 320, smem);
  k_check<<<nb, 320, smem>>>(bad, smids);
  cudaError_t e = cudaDeviceSynchronize();
  int hb = -1; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
  int hs[nb]; cudaMemcpy(hs, smids, 4 * nb, cudaMemcpyDeviceToHost);
  int per_sm[256] = {0}, two = 0, col0 = 0, col256 = 0;
  for (int k = 0; k < nb; ++k) {
    per_sm[hs[k] & 0xffff]++;
    if ((hs[k] >> 16) == 0) col0++;
    if ((hs[k] >> 16) == 256) col256++;
  }
  for (int k = 0; k < 256; ++k) if (per_sm[k] == 2) two++;
  printf("TMEM check: %s, occupancy %d CTAs/SM, mismatches %d, SMs holding two CTAs %d, base column 0: %d CTAs, 256: %d CTAs\n",
         cudaGetErrorString(e), occ, hb, two, col0, col256);

  long long* c; int* s;
  cudaMalloc(&c, 16); cudaMalloc(&s, 8);
  k_lat<<<1, 32>>>(c, 4096, s);
  k_lat<<<1, 32>>>(c, 4096, s);
  e = cudaDeviceSynchronize();
  long long hc[2]; cudaMemcpy(hc, c, 16, cudaMemcpyDeviceToHost);
  printf("tcgen05.ld.32x32b.x2 + wait::ld, dependent: %.1f cycles (%s); two x8 loads + one wait: %.1f cycles\n",
         (double)hc[0] / 4096, cudaGetErrorString(e), (double)hc[1] / 512);

  for (int blocks : {148, 296}) {
    run_chain<0, 4>("chain only (registers)", blocks);
    run_chain<1, 4>("TMEM x8 + shuffle, 4-node chunks", blocks);
    run_chain<2, 4>("TMEM x8 + shuffle + LDS y / STS x, 4-node", blocks);
    run_chain<2, 5>("TMEM x2 + shuffle + LDS y / STS x, 5-node", blocks);
  }
  return 0;
}
