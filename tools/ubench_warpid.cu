// Development aid: which hardware warp slots (and so which of the four SM sub-partitions, slot % 4) the warps of two
// co-resident 320-thread CTAs get.  nvcc -gencode arch=compute_100a,code=sm_100a -o ubench_warpid ubench_warpid.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(320, 2) probe(int* out, int spin) {
  extern __shared__ double sm[];
  unsigned smid, wid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  if ((threadIdx.x & 31) == 0) {
    int* o = out + (blockIdx.x * 10 + (threadIdx.x >> 5)) * 2;
    o[0] = (int)smid;
    o[1] = (int)wid;
  }
  // stay resident so that the second wave cannot reuse the slots of the first
  long long t0 = clock64();
  while (clock64() - t0 < spin) {}
  if (sm[threadIdx.x] == 123.456) out[0] = 0;
}
int main() {
  int n = 296;
  int* d;
  cudaMalloc(&d, n * 20 * sizeof(int));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 104000);
  probe<<<n, 320, 104000>>>(d, 2000000);
  cudaDeviceSynchronize();
  static int h[296 * 20];
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  int hist[2][4] = {{0}};
  for (int b = 0; b < n; ++b) {
    if (b < 4 || (b >= 148 && b < 152)) {
      printf("cta %3d sm %3d warps:", b, h[b * 20]);
      for (int w = 0; w < 10; ++w) printf(" %2d", h[(b * 10 + w) * 2 + 1]);
      printf("\n");
    }
    for (int w = 0; w < 10; ++w) hist[b >= 148][h[(b * 10 + w) * 2 + 1] & 3]++;
  }
  printf("sub-partition histogram, first 148 CTAs: %d %d %d %d; last 148: %d %d %d %d\n", hist[0][0], hist[0][1], hist[0][2], hist[0][3],
         hist[1][0], hist[1][1], hist[1][2], hist[1][3]);
  return 0;
}
