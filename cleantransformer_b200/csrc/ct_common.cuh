// ct_common.cuh — shared host/device helpers for libct_b200.so (sm_100a only).
//
// Host side: thread-local error buffer + return-code convention:
//   0 = ok, >0 = cudaError_t passthrough, <0 = library code.
// Device side: thin inline-PTX wrappers for mbarrier / TMA / tcgen05 / TMEM.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>

#define CT_ERR_BAD_ARG      (-1)
#define CT_ERR_UNSUPPORTED  (-2)
#define CT_ERR_WORKSPACE    (-3)
#define CT_ERR_COMM         (-4)

namespace ct {

// Thread-local last-error message (re-entrant: main thread + autograd thread).
void set_error(const char* fmt, ...);
const char* get_error();

#define CT_CUDA_OK(expr)                                                        \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      ct::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,               \
                    cudaGetErrorString(_e));                                    \
      return (int)_e;                                                           \
    }                                                                           \
  } while (0)

#define CT_REQUIRE(cond, code, ...)                                             \
  do {                                                                          \
    if (!(cond)) {                                                              \
      ct::set_error(__VA_ARGS__);                                               \
      return (code);                                                            \
    }                                                                           \
  } while (0)

#define CT_LAUNCH_OK()                                                          \
  do {                                                                          \
    cudaError_t _e = cudaGetLastError();                                        \
    if (_e != cudaSuccess) {                                                    \
      ct::set_error("%s:%d: launch failed: %s", __FILE__, __LINE__,             \
                    cudaGetErrorString(_e));                                    \
      return (int)_e;                                                           \
    }                                                                           \
  } while (0)

int sm_count();  // cached multiprocessor count of the current device

// ---- programmatic dependent launch -----------------------------------------------------------------------------------
// A training step is ~620 kernels, a decode step ~170, each waiting for the previous grid to drain before its own launch
// latency even starts: 3.0 - 3.5 ms of a 38 ms step were idle gaps between graph nodes (profiles/r02t_bench.json).
// Kernels launched through launch_k() carry cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs may become
// resident while the previous kernel is still running (it allows that with pdl_launch_dependents() at its top) and run
// their prologue — barrier init, TMEM allocation, descriptor prefetch — until pdl_wait(), which returns once the previous
// grid has completed and its memory is visible. EVERY kernel launched this way calls pdl_wait() before its first access to
// global memory; without the attribute both instructions are no-ops.
// PDL option: 1 = on, 2 = off, 0 = auto: off, except inside the captured decode step (generation.py switches it on
// around the capture). Measured (profiles/r02u_*, r02v_*, r02x_*): decode +9 %; training step on one GPU +0.6 % (three
// alternating pairs of runs on a power-capped box); under DistributedDataParallel -12 % at two GPUs — the pre-launched CTAs of the compute
// chain take the SM slots that the bucket all-reduce kernels of the side stream used to slip into.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// Tuning knobs (ct_set_option / CT_<NAME> environment variables); 0 always means "default / auto".
enum : int { OPT_LN_BWD_IMPL = 0, OPT_GEMM_EPI_IMPL, OPT_GEMM_2CTA, OPT_CE_IMPL,
             OPT_GEMM_SPLITK, OPT_ATTN_BWD_IMPL, OPT_PDL, OPT_COUNT };
int option(int which);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = option(OPT_PDL) == 1 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// dtype enum shared with include/ct_b200.h
enum : int { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };
// activation enum shared with include/ct_b200.h
enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU_ERF = 2, ACT_GELU_TANH = 3, ACT_TANH = 4,
              // GEMM epilogues only: y = gelu_tanh(t) with gelu_tanh'(t) (not t) written to `preact`; and, as an
              // activation-gradient kind, "the source already holds act'(pre)": multiply by it as is
              ACT_GELU_TANH_SAVE_GRAD = 5, ACT_GRAD_PRECOMPUTED = 6 };

#ifdef __CUDACC__

// ---------------------------------------------------------------------------------------------
// generic device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// The same packing on the integer ALU pipe: add half an ulp of bf16 to the fp32 bit pattern, keep the upper halves
// (round to nearest, ties away from zero; Inf stays Inf). F2FP.BF16.PACK_AB is issued to the XU pipe — the one MUFU.EX2
// uses — and occupies it twice as long as an ex2 (two conversions per lane): in the attention kernels, where every
// exponential is followed by a conversion, that made the CONVERSIONS the bottleneck (r02g ncu: XU 65 % busy in the
// forward kernel with 2.4 M MUFU.EX2 + 1.3 M F2FP warp instructions). IADD + PRMT cost 1.5 ALU slots per element.
__device__ __forceinline__ uint32_t pack_bf16x2_alu(float lo, float hi) {
  const uint32_t a = __float_as_uint(lo) + 0x8000u, b = __float_as_uint(hi) + 0x8000u;
  return __byte_perm(a, b, 0x7632);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Activations. gelu_tanh follows modeling_bloom.py:335-345 (constant 0.79788456) which equals
// modeling_gpt.py:112-122 up to rounding of sqrt(2/pi).
// ---- counter-based dropout mask (include/ct_b200.h: "dropout") ----
struct DropKey { uint32_t key0, key1, thr; float rscale; };  // thr == 0: dropout off
__host__ __device__ __forceinline__ DropKey make_drop_key(float p, uint64_t seed, uint32_t stream) {
  DropKey k;
  k.key0 = (uint32_t)seed + stream * 0x632BE5ABu;
  k.key1 = (uint32_t)(seed >> 32) ^ (stream * 0x2545F491u);
  k.thr = p > 0.f ? (uint32_t)(p * 16777216.f + 0.5f) : 0u;
  k.rscale = p > 0.f ? 1.f / (1.f - p) : 1.f;
  return k;
}
__device__ __forceinline__ uint32_t drop_mix(uint32_t h) {
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  return h;
}
// `pre` = hi * 0x85EBCA77 + key1 (constant per (b, h) in attention); lo_term = lo * 0x9E3779B1 + key0
__device__ __forceinline__ bool drop_keep_pre(uint32_t lo_term, uint32_t pre, uint32_t thr) {
  return (drop_mix(lo_term ^ pre) >> 8) >= thr;
}
__device__ __forceinline__ bool drop_keep(const DropKey& k, uint32_t hi, uint32_t lo) {
  return drop_keep_pre(lo * 0x9E3779B1u + k.key0, hi * 0x85EBCA77u + k.key1, k.thr);
}

__device__ __forceinline__ float act_apply(float x, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(x, 0.f);
    case ACT_GELU_ERF: return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    case ACT_GELU_TANH:
    case ACT_GELU_TANH_SAVE_GRAD: {
      float u = 0.79788456f * x * (1.f + 0.044715f * x * x);
      float hx = 0.5f * x;
      return fmaf(hx, tanh_approx(u), hx);
    }
    case ACT_TANH: return tanhf(x);  // BERT pooler, modeling_bert.py:283-286
    default: return x;
  }
}
// d act / dx evaluated at the pre-activation x. gelu_tanh follows bloom_gelu_back
// (modeling_bloom.py:348-363).
__device__ __forceinline__ float act_grad(float x, int act) {
  switch (act) {
    case ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case ACT_GELU_ERF:
      return 0.5f * (1.f + erff(x * 0.70710678118654752f)) +
             x * 0.3989422804014327f * __expf(-0.5f * x * x);
    case ACT_GELU_TANH:
    case ACT_GELU_TANH_SAVE_GRAD: {
      float t = tanh_approx(0.79788456f * x * (1.f + 0.044715f * x * x));
      return 0.5f * x * ((1.f - t * t) * (0.79788456f + 0.1070322243f * x * x)) + 0.5f * (1.f + t);
    }
    case ACT_TANH: {
      float t = tanhf(x);
      return 1.f - t * t;
    }
    case ACT_GRAD_PRECOMPUTED: return x;  // x IS the saved derivative
    default: return 1.f;
  }
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  // make barrier inits visible to the async proxy (TMA / tcgen05.commit)
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Bounded wait: a protocol bug traps (sticky CUDA error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      printf("ct_b200: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — 2D/3D tile loads into shared memory, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (UMMA / TMA store)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16/fp16 inputs with fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All previously issued tcgen05.mma of this thread arrive on `bar` when complete
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base+i), cols [c, c+32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}


// ---- 2-CTA (cta_group::2) variants: a CTA pair in a cluster shares one 256-row MMA ----------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` inside CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma_f16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same smem offset in BOTH CTAs of the pair when the MMAs retire
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
// TMA tile load issued by either CTA of the pair; completion bytes are credited to `leader_bar`
// (a shared::cluster address inside CTA 0)
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const void* tmap, uint32_t leader_bar, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tmap), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}

// ---- UMMA descriptors (bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor / InstrDescriptor)
// Shared-memory matrix descriptor, SWIZZLE_128B, version 1 (Blackwell).
//   bits [0,14)  start address >> 4      bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4 bits [46,48) version = 1
//   bits [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16, fp32 accumulate.
//   [4,6) c_format=1(F32)  [7,10) a_format  [10,13) b_format (0=F16, 1=BF16)
//   [15] a_major (0=K,1=MN)  [16] b_major  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_f16(int fmt /*0 f16, 1 bf16*/, int a_mn_major,
                                                      int b_mn_major, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#endif  // __CUDACC__

// Host: build a CUtensorMap through the driver entry point (no link-time libcuda dependency).
// dims/strides are innermost-first; strides in bytes for dims 1..rank-1.
int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, int swizzle_128b);

}  // namespace ct
