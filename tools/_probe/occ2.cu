// probe: what limits the forward attention kernel to one CTA per SM? Occupancy of stand-in kernels with the same
// launch bounds / dynamic shared memory, with and without setmaxnreg, plus the device limits.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(384, 2) k384_plain(float* p) { extern __shared__ float s[]; s[threadIdx.x] = 1.f; __syncthreads(); if (p) p[0] = s[0]; }
__global__ void __launch_bounds__(384, 2) k384_setmax(float* p) {
  extern __shared__ float s[];
  if (threadIdx.x >= 256) { asm volatile("setmaxnreg.dec.sync.aligned.u32 32;"); }
  else { asm volatile("setmaxnreg.inc.sync.aligned.u32 104;"); }
  s[threadIdx.x] = 1.f; __syncthreads(); if (p) p[0] = s[0];
}
__global__ void __launch_bounds__(192, 2) k192_plain(float* p) { extern __shared__ float s[]; s[threadIdx.x] = 1.f; __syncthreads(); if (p) p[0] = s[0]; }
template <typename K> void probe(K kern, const char* name, int threads) {
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
  for (int smem : {115328, 114688, 113664, 112640, 110592, 100000}) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int carve : {-1, 100}) {
      cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      int n = -1; cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem);
      printf("%s regs=%d static=%zu threads=%d dyn_smem=%d carveout=%d -> %d CTAs/SM %s\n", name, fa.numRegs, fa.sharedSizeBytes, threads, smem, carve, n, e ? cudaGetErrorString(e) : "");
    }
  }
}
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("sharedMemPerMultiprocessor=%zu sharedMemPerBlockOptin=%zu reservedSharedMemPerBlock=%zu regsPerMultiprocessor=%d maxThreadsPerSM=%d maxBlocksPerSM=%d\n",
         pr.sharedMemPerMultiprocessor, pr.sharedMemPerBlockOptin, pr.reservedSharedMemPerBlock, pr.regsPerMultiprocessor, pr.maxThreadsPerMultiProcessor, pr.maxBlocksPerMultiProcessor);
  probe(k384_plain, "k384_plain", 384); probe(k384_setmax, "k384_setmax", 384); probe(k192_plain, "k192_plain", 192);
  return 0;
}
