// Development aid: dependent-issue latencies that bound the line solves (FP64 chains, shared and L2 loads).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench_lat ubench_lat.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE>
__global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
  double x = a + threadIdx.x * 1e-9, y = b;
  __shared__ double sm[64];
  if (threadIdx.x < 64) sm[threadIdx.x] = a;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (MODE == 0) x = __dadd_rn(x, b);
      if (MODE == 1) x = __dmul_rn(x, a);
      if (MODE == 2) { x = __dmul_rn(x, a); x = __dadd_rn(y, -x); }
      if (MODE == 3) x = __fma_rn(x, a, b);
      if (MODE == 4) { x = __dmul_rn(x, a); x = __dadd_rn(y, -x); sm[threadIdx.x & 63] = x; }   // + STS
      if (MODE == 5) { x = __dadd_rn(sm[(k + threadIdx.x) & 63], -__dmul_rn(x, a)); sm[threadIdx.x & 63] = x; }  // LDS off-chain
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (x == 123.456) out[0] = x;
}

__global__ void chase(const int* next, long long* cyc, int iters, int* sink) {
  int p = 0;
  for (int it = 0; it < 64; ++it) p = next[p];
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) p = __ldcg(&next[p]);
  const long long t1 = clock64();
  cyc[0] = t1 - t0;
  *sink = p;
}

template <int MODE>
void run(const char* name, int threads, int blocks, int ops_per_iter) {
  double* d; long long* c;
  cudaMalloc(&d, 8); cudaMalloc(&c, 8 * blocks);
  const int iters = 4096;
  chain<MODE><<<blocks, threads>>>(d, c, iters, 1.0000001, 1e-9);
  chain<MODE><<<blocks, threads>>>(d, c, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("%-28s threads %4d blocks %4d: %.2f cycles per dependent op\n", name, threads, blocks, (double)h / (iters * 8.0 * ops_per_iter));
  cudaFree(d); cudaFree(c);
}

int main() {
  for (int threads : {32, 128, 320, 640}) {
    run<0>("DADD chain", threads, 148, 1);
    run<1>("DMUL chain", threads, 148, 1);
    run<2>("DMUL->DADD chain", threads, 148, 2);
    run<3>("DFMA chain", threads, 148, 1);
    run<4>("DMUL->DADD + STS", threads, 148, 2);
    run<5>("LDS, DMUL->DADD, STS", threads, 148, 2);
  }
  // L2-hit pointer chase (stride 4 KB over 8 MB)
  const int n = 1 << 21;
  int* h = new int[n];
  for (int i = 0; i < n; ++i) h[i] = 0;
  const int stride = 1024, cnt = n / stride;
  for (int k = 0; k < cnt; ++k) h[k * stride] = ((k * 37 + 11) % cnt) * stride;
  int* d; long long* c; int* s;
  cudaMalloc(&d, n * 4); cudaMalloc(&c, 8); cudaMalloc(&s, 4);
  cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
  chase<<<1, 1>>>(d, c, 2000, s);
  chase<<<1, 1>>>(d, c, 2000, s);
  cudaDeviceSynchronize();
  long long hc; cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
  printf("L2 pointer chase: %.1f cycles per load\n", (double)hc / 2000);
  return 0;
}
