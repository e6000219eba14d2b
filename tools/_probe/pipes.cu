// probe: issue cost (cycles per warp instruction per SM sub-partition) of the instructions the attention softmax is
// made of, on this GPU: MUFU.EX2, F2FP.BF16 pack, FFMA2 / FADD2, FMNMX, IADD+PRMT. One CTA per SM, W warps per
// sub-partition, each warp runs N iterations of 8 independent chains; cycles measured with clock64 around the loop.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#define ITERS 4096
template <int OP>
__global__ void k(float* out, long long* cyc) {
  float a[8];
  uint32_t u[8];
  for (int i = 0; i < 8; ++i) { a[i] = 0.001f * (threadIdx.x + i); u[i] = threadIdx.x * 77u + i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7])); u[i] ^= r; }
      if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
      if (OP == 3) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 3) & 7]));
      if (OP == 4) { asm volatile("add.u32 %0, %0, 32768;" : "+r"(u[i])); }
      if (OP == 5) { asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(u[(i + 1) & 7])); }
      if (OP == 6) { unsigned long long x, y; asm volatile("mov.b64 %0, {%2, %3}; mov.b64 %1, {%3, %2}; fma.rn.f32x2 %0, %0, %1, %1; mov.b64 {%2, %3}, %0;" : "=l"(x), "=l"(y), "+f"(a[i]), "+f"(a[(i + 1) & 7])); }
    }
  }
  long long t1 = clock64();
  float s = 0; uint32_t v = 0;
  for (int i = 0; i < 8; ++i) { s += a[i]; v ^= u[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)v;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name, float* out, long long* cyc) {
  for (int warps_per_smsp : {1, 2, 4, 8}) {
    int threads = warps_per_smsp * 4 * 32;
    k<OP><<<148, threads>>>(out, cyc);
    cudaDeviceSynchronize();
    k<OP><<<148, threads>>>(out, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
    // per sub-partition: warps_per_smsp warps x ITERS x 8 instructions in `avg` cycles
    printf("%-22s warps/SMSP=%d  cycles per warp-instruction per SMSP = %.2f\n", name, warps_per_smsp, avg / ((double)ITERS * 8 * warps_per_smsp));
  }
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  run<0>("MUFU.EX2", out, cyc); run<1>("F2FP.BF16.PACK_AB", out, cyc); run<2>("FFMA", out, cyc); run<3>("FMNMX", out, cyc);
  run<4>("IADD", out, cyc); run<5>("PRMT", out, cyc); run<6>("FFMA2 (+movs)", out, cyc);
  cudaError_t e = cudaGetLastError(); printf("%s\n", cudaGetErrorString(e));
  return 0;
}
