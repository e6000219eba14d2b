// probe: how many clusters of a 1024-thread, large-shared-memory kernel can be co-resident on this GPU
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(1024, 1) k1024(float* p) { extern __shared__ float s[]; s[threadIdx.x] = 1.f; __syncthreads(); if (p) p[0] = s[0]; }
__global__ void __launch_bounds__(512, 1) k512(float* p) { extern __shared__ float s[]; s[threadIdx.x] = 1.f; __syncthreads(); if (p) p[0] = s[0]; }
template <typename K> void probe(K kern, const char* name, int threads) {
  for (int smem_kb : {32, 64, 100, 128, 200}) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
    for (int cs : {1, 2, 4, 8}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(148 / cs * cs); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem_kb * 1024;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
      printf("%s smem=%dKB cluster=%d -> max active clusters %d (%d CTAs) %s\n", name, smem_kb, cs, n, n * cs, e ? cudaGetErrorString(e) : "");
    }
  }
}
int main() { probe(k1024, "k1024", 1024); probe(k512, "k512", 512); return 0; }
