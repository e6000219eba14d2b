// gemm.cu — the GEMM family behind every nn.Linear / Conv1D on the hot path, forward and backward.
//
//   C[M,N] = epilogue(alpha * sum_k A(m,k) * B(n,k))     (see include/ct_b200.h: ct_gemm_args)
//
// Main kernel: persistent, warp-specialised tcgen05 GEMM for sm_100a
//   warp 0      TMA producer  (cp.async.bulk.tensor, SWIZZLE_128B tiles, 4-6 stage mbarrier ring)
//   warp 1      MMA issuer    (one thread issues tcgen05.mma kind::f16, M=128 x N=BN x K=16,
//                              fp32 accumulators in TMEM, double-buffered across tiles)
//   warps 2..5  epilogue      (tcgen05.ld TMEM->registers, fused bias / activation / activation-grad
//                              / residual / accumulate, 128-bit global stores or red.add for split-K)
// Operands may be K-major (row = M|N index, K contiguous) or MN-major (row = K index, M|N
// contiguous): forward uses K-major x K-major, dgrad K-major x MN-major, wgrad MN-major x MN-major,
// Conv1D ([in,out] weight, modeling_gpt.py:32-46) K-major x MN-major. No transposes are
// materialised.
//
// Reference lines replaced: transformer.py:37,98-102; modeling_bloom.py:79,121-122,256,267-269;
// modeling_gpt.py:45,71,106,133-135; modeling_bert.py:238-247 and the dgrad/wgrad GEMMs autograd
// derives from them. Tensor-pipe bound: 2*M*N*K flop per call.
//
// Fallback kernel: a plain SIMT tile GEMM with the same epilogue, used for tiny or TMA-unaligned
// problems (e.g. the 28-way BERT classifier) and as an in-library cross-check (impl = 2).
#include "ct_common.cuh"
#include "../../include/ct_b200.h"
#include <cstdlib>
#include <cstring>

namespace ct {

#ifdef CT_DEBUG_TIMING
__device__ long long ct_dbg_gemm[4096];
#define CT_GDBG(thread, slot) do { if (blockIdx.x == 5 && threadIdx.x == (thread) && (slot) < 4096) ct_dbg_gemm[(slot)] = clock64(); } while (0)
#else
#define CT_GDBG(thread, slot) do {} while (0)
#endif

struct EpiParams {
  int M, N;
  void* C; int c_dtype; int64_t ldc;
  float alpha, beta;
  const float* bias;
  int act;
  void* preact; int preact_dtype; int64_t ldp;
  const void* actgrad_src; int actgrad_dtype; int actgrad_act; int64_t ldg;
  const void* residual; int res_dtype; int64_t ldr;
  int vec_ok;     // all pointers 16B aligned and leading dims multiples of 8 elements
  int atomic_out; // split-K: C += t via red.global.add.f32 (C f32, linear epilogue only)
  int res_coalesced;  // EPIK_RES_F32: residual fetched (and added) in the coalesced write-out pattern
  float* row_stats;   // EPIK_BF16_STATS: [2 * ceil(N/256)][M] float2 (max2, sum2) softmax statistics of the stored row
};

__device__ __forceinline__ float ld_elem(const void* p, int dt, int64_t i) {
  if (dt == DT_F32) return reinterpret_cast<const float*>(p)[i];
  if (dt == DT_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}
__device__ __forceinline__ void st_elem(void* p, int dt, int64_t i, float v) {
  if (dt == DT_F32) reinterpret_cast<float*>(p)[i] = v;
  else if (dt == DT_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
}

// Code size matters here: the tcgen05 kernel's epilogue is executed by 4 warps while the TMA and MMA
// warps run other code, and a fully inlined epilogue (32 unrolled elements x every activation x every
// dtype x scalar fallbacks) reached 47k SASS instructions (750 KB) and stalled on instruction fetch.
// Rare paths are therefore real function calls.
__device__ __noinline__ float act_apply_call(float x, int act) { return act_apply(x, act); }
__device__ __noinline__ float act_grad_call(float x, int act) { return act_grad(x, act); }

__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float u = 0.79788456f * x * (1.f + 0.044715f * x * x);
  const float hx = 0.5f * x;
  return fmaf(hx, tanh_approx(u), hx);
}
__device__ __forceinline__ float gelu_tanh_grad_fast(float x) {
  const float t = tanh_approx(0.79788456f * x * (1.f + 0.044715f * x * x));
  return 0.5f * x * ((1.f - t * t) * (0.79788456f + 0.1070322243f * x * x)) + 0.5f * (1.f + t);
}
__device__ __forceinline__ void act32(float (&t)[32], int act) {
  if (act == ACT_GELU_TANH || act == ACT_GELU_TANH_SAVE_GRAD) {
#pragma unroll
    for (int j = 0; j < 32; ++j) t[j] = gelu_tanh_fast(t[j]);
  } else if (act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) t[j] = fmaxf(t[j], 0.f);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) t[j] = act_apply_call(t[j], act);
  }
}
__device__ __forceinline__ void actgrad32(float (&t)[32], const float (&s)[32], int act) {
  if (act == ACT_GELU_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) t[j] *= gelu_tanh_grad_fast(s[j]);
  } else if (act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) t[j] = s[j] > 0.f ? t[j] : 0.f;
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) t[j] *= act_grad_call(s[j], act);
  }
}

// Scalar epilogue for one element (SIMT kernel and ragged edges of the tcgen05 kernel).
__device__ __noinline__ void epi_scalar(const EpiParams& e, int m, int n, float acc, bool add_bias) {
  float t = e.alpha * acc;
  if (e.bias && add_bias) t += e.bias[n];
  if (e.atomic_out) {
    atomicAdd(reinterpret_cast<float*>(e.C) + (int64_t)m * e.ldc + n, t);
    return;
  }
  if (e.preact)
    st_elem(e.preact, e.preact_dtype, (int64_t)m * e.ldp + n,
            e.act == ACT_GELU_TANH_SAVE_GRAD ? act_grad_call(t, ACT_GELU_TANH) : t);
  if (e.act != ACT_NONE) t = act_apply_call(t, e.act);
  if (e.actgrad_src)
    t *= act_grad_call(ld_elem(e.actgrad_src, e.actgrad_dtype, (int64_t)m * e.ldg + n), e.actgrad_act);
  if (e.residual) t += ld_elem(e.residual, e.res_dtype, (int64_t)m * e.ldr + n);
  if (e.beta != 0.f) t += e.beta * ld_elem(e.C, e.c_dtype, (int64_t)m * e.ldc + n);
  st_elem(e.C, e.c_dtype, (int64_t)m * e.ldc + n, t);
}

// 32 consecutive elements of one row, 128-bit accesses.
__device__ __forceinline__ void ld32(const void* p, int dt, int64_t off, float (&v)[32]) {
  if (dt == DT_F32) {
    const float4* q = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + off);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 a = q[i];
      v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
    }
  } else {
    const uint4* q = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p) + off);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 a = q[i];
      const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 f;
        if (dt == DT_BF16) f = unpack_bf16x2(w[k]);
        else { __half2 h = *reinterpret_cast<const __half2*>(&w[k]); f = __half22float2(h); }
        v[8 * i + 2 * k] = f.x; v[8 * i + 2 * k + 1] = f.y;
      }
    }
  }
}

// Direct (row-per-thread) epilogue for 32 consecutive columns [n0, n0+32) of row m: split-K
// red.add output, or the scalar fallback for ragged / unaligned problems.
__device__ __forceinline__ void epi_chunk32(const EpiParams& e, int m, int n0, float (&t)[32],
                                            bool add_bias) {
  if (m >= e.M || n0 >= e.N) return;
  const bool full = (n0 + 32 <= e.N) && e.vec_ok;
  if (full && e.atomic_out) {
    float* c = reinterpret_cast<float*>(e.C) + (int64_t)m * e.ldc + n0;
    const float4* b4 = reinterpret_cast<const float4*>(e.bias + n0);
    const bool bias = e.bias && add_bias;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 b = bias ? __ldg(b4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(c + 4 * i),
                   "f"(fmaf(e.alpha, t[4 * i], b.x)), "f"(fmaf(e.alpha, t[4 * i + 1], b.y)),
                   "f"(fmaf(e.alpha, t[4 * i + 2], b.z)), "f"(fmaf(e.alpha, t[4 * i + 3], b.w))
                   : "memory");
    }
    return;
  }
  // static indices only: a runtime-indexed t[] would be demoted to local memory for EVERY path
  const int lim = e.N - n0;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (j < lim) epi_scalar(e, m, n0 + j, t[j], add_bias);
}

// ---- coalesced epilogue stores --------------------------------------------------------------------
// A thread owns one accumulator ROW, so a direct store instruction of a warp touches 32 different
// rows (32 half-used sectors per request). Instead each warp transposes its 32x32 result block
// through a private smem buffer (row stride 144 B: conflict-free 16-byte row writes) and writes it
// out with lanes running along the row: 64 B (16-bit) or 128 B (fp32) contiguous per row, 8 or 4
// rows per instruction.
constexpr int EPI_STAGE_ROW = 64;
constexpr int EPI_STAGE_BYTES = 32 * EPI_STAGE_ROW;  // 2 KB per epilogue warp
constexpr int EPI_WARPS = 8;

// Staging rows are 64 bytes (32 x 16-bit, or HALF of a 32 x fp32 chunk: fp32 goes out in two
// rounds). 16-byte chunk i of row r lives at position i ^ ((r >> 1) & 3): with rows 2k / 2k+1 in the
// two halves of a 128-byte line, both the row-per-thread writes and the lanes-along-the-row reads
// are bank-conflict free without padding. (Keeping the buffers at 16 KB total leaves enough shared
// memory for an all-reduce CTA to be co-resident with a GEMM CTA during the DDP backward.)
__device__ __forceinline__ void stage_round(uint32_t stage, int lane, uint32_t w0, uint32_t w1, uint32_t w2,
                                            uint32_t w3, int i) {
  const uint32_t my = stage + lane * EPI_STAGE_ROW;
  const int key = (lane >> 1) & 3;
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my + 16 * (i ^ key)), "r"(w0), "r"(w1), "r"(w2),
               "r"(w3) : "memory");
}
// write out the staged 32 rows x 64 bytes: 8 rows x 64 B per instruction
__device__ __forceinline__ void stage_flush(uint32_t stage, int lane, uint8_t* gbase_bytes, int64_t ld_bytes,
                                            int m_base, int M) {
  const int piece = lane & 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = (lane >> 2) + 8 * k;
    uint4 w;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w)
                 : "r"(stage + r * EPI_STAGE_ROW + 16 * (piece ^ ((r >> 1) & 3))));
    const int m = m_base + r;
    if (m < M) *reinterpret_cast<uint4*>(gbase_bytes + (int64_t)m * ld_bytes + piece * 16) = w;
  }
}

// fp32 chunk + fp32 residual: the residual pieces were fetched in the SAME lane -> (row, 16-byte piece) pattern as the
// write-out (epi_fast_aux_load, coalesced form), so the add happens here, after the transpose through the stage
__device__ __forceinline__ void stage_store32_f32_res(uint32_t stage, int lane, const float (&v)[32], void* gbase,
                                                      int64_t ld, int m_base, int M, int n0, const uint4 (&x)[8]) {
  uint8_t* g = reinterpret_cast<uint8_t*>(gbase) + (int64_t)n0 * 4;
  const int piece = lane & 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      stage_round(stage, lane, __float_as_uint(v[16 * h + 4 * i]), __float_as_uint(v[16 * h + 4 * i + 1]),
                  __float_as_uint(v[16 * h + 4 * i + 2]), __float_as_uint(v[16 * h + 4 * i + 3]), i);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = (lane >> 2) + 8 * k;
      float4 w;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w)
                   : "r"(stage + r * EPI_STAGE_ROW + 16 * (piece ^ ((r >> 1) & 3))));
      const uint4 q = x[4 * h + k];
      w.x += __uint_as_float(q.x); w.y += __uint_as_float(q.y); w.z += __uint_as_float(q.z); w.w += __uint_as_float(q.w);
      const int m = m_base + r;
      if (m < M) *reinterpret_cast<float4*>(g + 64 * h + (int64_t)m * ld * 4 + piece * 16) = w;
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void stage_store32(uint32_t stage, int lane, const float (&v)[32], int dt,
                                              void* gbase, int64_t ld, int m_base, int M, int n0) {
  if (dt == DT_F32) {
    uint8_t* g = reinterpret_cast<uint8_t*>(gbase) + (int64_t)n0 * 4;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        stage_round(stage, lane, __float_as_uint(v[16 * h + 4 * i]), __float_as_uint(v[16 * h + 4 * i + 1]),
                    __float_as_uint(v[16 * h + 4 * i + 2]), __float_as_uint(v[16 * h + 4 * i + 3]), i);
      __syncwarp();
      stage_flush(stage, lane, g + 64 * h, ld * 4, m_base, M);
      __syncwarp();
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t w0, w1, w2, w3;
      if (dt == DT_BF16) {
        w0 = pack_bf16x2(v[8 * i], v[8 * i + 1]); w1 = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
        w2 = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]); w3 = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
      } else {
        __half2 h;
        h = __floats2half2_rn(v[8 * i], v[8 * i + 1]); w0 = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2half2_rn(v[8 * i + 2], v[8 * i + 3]); w1 = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2half2_rn(v[8 * i + 4], v[8 * i + 5]); w2 = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2half2_rn(v[8 * i + 6], v[8 * i + 7]); w3 = *reinterpret_cast<uint32_t*>(&h);
      }
      stage_round(stage, lane, w0, w1, w2, w3, i);
    }
    __syncwarp();
    stage_flush(stage, lane, reinterpret_cast<uint8_t*>(gbase) + (int64_t)n0 * 2, ld * 2, m_base, M);
    __syncwarp();
  }
}

// The epilogue is latency-bound: besides running 8 warps (two per TMEM lane quarter, alternating
// 32-column chunks) every global operand it consumes is fetched one of its chunks AHEAD. aux_kind names the single per-element operand that is prefetched:
//   1 = residual, 2 = activation-gradient source (saved pre-activation), 3 = old C (beta accumulate).
enum : int { AUX_NONE = 0, AUX_RES = 1, AUX_ACTGRAD = 2, AUX_C = 3 };

__device__ __forceinline__ int epi_aux_kind(const EpiParams& e) {
  if (!e.vec_ok || e.atomic_out) return AUX_NONE;
  if (e.residual) return AUX_RES;
  if (e.actgrad_src) return AUX_ACTGRAD;
  if (e.beta != 0.f) return AUX_C;
  return AUX_NONE;
}
__device__ __forceinline__ void epi_aux_load(const EpiParams& e, int kind, int m, int n0, float (&a)[32]) {
  if (kind == AUX_NONE || m >= e.M || n0 + 32 > e.N) return;
  if (kind == AUX_RES) ld32(e.residual, e.res_dtype, (int64_t)m * e.ldr + n0, a);
  else if (kind == AUX_ACTGRAD) ld32(e.actgrad_src, e.actgrad_dtype, (int64_t)m * e.ldg + n0, a);
  else ld32(e.C, e.c_dtype, (int64_t)m * e.ldc + n0, a);
}

// Warp-collective epilogue for a 32-row x 32-column block: lane l owns row m_base + l.
// bias_s: shared-memory address of this tile's bias slice (floats, indexed from the tile's n origin).
__device__ __forceinline__ void epi_block32(const EpiParams& e, int m_base, int lane, int n0, int n_in_tile,
                                            float (&t)[32], const float (&aux)[32], int aux_kind, bool add_bias,
                                            uint32_t stage, uint32_t bias_s) {
  if (n0 >= e.N) return;  // warp-uniform
  const int m = m_base + lane;
  const bool full = (n0 + 32 <= e.N) && e.vec_ok;  // warp-uniform
  if (!full || e.atomic_out) {
    epi_chunk32(e, m, n0, t, add_bias);
    return;
  }
  const bool row_ok = m < e.M;
#pragma unroll
  for (int j = 0; j < 32; ++j) t[j] *= e.alpha;
  if (e.bias && add_bias) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 b;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                   : "r"(bias_s + 4 * (n_in_tile + 4 * i)));
      t[4 * i] += b.x; t[4 * i + 1] += b.y; t[4 * i + 2] += b.z; t[4 * i + 3] += b.w;
    }
  }
  if (e.preact) {
    if (e.act == ACT_GELU_TANH_SAVE_GRAD) {
      float g[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) g[j] = gelu_tanh_grad_fast(t[j]);
      stage_store32(stage, lane, g, e.preact_dtype, e.preact, e.ldp, m_base, e.M, n0);
    } else {
      stage_store32(stage, lane, t, e.preact_dtype, e.preact, e.ldp, m_base, e.M, n0);
    }
  }
  if (e.act != ACT_NONE) act32(t, e.act);
  if (e.actgrad_src && row_ok) {
    if (aux_kind == AUX_ACTGRAD) {
      actgrad32(t, aux, e.actgrad_act);
    } else {
      float s[32];
      ld32(e.actgrad_src, e.actgrad_dtype, (int64_t)m * e.ldg + n0, s);
      actgrad32(t, s, e.actgrad_act);
    }
  }
  if (e.residual && row_ok) {
    if (aux_kind == AUX_RES) {
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] += aux[j];
    } else {
      float s[32];
      ld32(e.residual, e.res_dtype, (int64_t)m * e.ldr + n0, s);
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] += s[j];
    }
  }
  if (e.beta != 0.f && row_ok) {
    if (aux_kind == AUX_C) {
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = fmaf(e.beta, aux[j], t[j]);
    } else {
      float s[32];
      ld32(e.C, e.c_dtype, (int64_t)m * e.ldc + n0, s);
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = fmaf(e.beta, s[j], t[j]);
    }
  }
  stage_store32(stage, lane, t, e.c_dtype, e.C, e.ldc, m_base, e.M, n0);
}

// ---- specialised epilogues -------------------------------------------------------------------------
// The generic epilogue above decides dtype / activation / operand set per chunk at run time and
// measured ~23 warp instructions per output element on the two epilogue-bound layer GEMMs (FFN1 forward:
// bias + GELU + saved pre-activation; FFN2 dgrad: x GELU'(pre)). The hot Bloom / GPT-2 combinations get
// compile-time variants: packed f32x2 math, no register copies (the TMEM loads ping-pong between two
// register sets), the per-element operand (residual / saved pre-activation) prefetched one chunk ahead.
enum : int { EPIK_GENERIC = 0, EPIK_BF16 = 1, EPIK_GELU_PRE = 2, EPIK_RES_F32 = 3, EPIK_ACTGRAD = 4,
              EPIK_BF16_STATS = 5 /* LM head: bf16 logits + per (row, 128-column half tile) softmax statistics */ };

__device__ __forceinline__ float2 gelu_tanh2(float2 x) {
  const float2 x2 = __fmul2_rn(x, x);
  const float2 pp = __ffma2_rn(x2, make_float2(0.79788456f * 0.044715f, 0.79788456f * 0.044715f),
                               make_float2(0.79788456f, 0.79788456f));
  const float2 u = __fmul2_rn(x, pp);
  const float2 th = make_float2(tanh_approx(u.x), tanh_approx(u.y));
  const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  return __ffma2_rn(hx, th, hx);
}
// d gelu_tanh / dx at x (modeling_bloom.py:348-363)
__device__ __forceinline__ float2 gelu_tanh_grad2(float2 x) {
  const float2 x2 = __fmul2_rn(x, x);
  const float2 pp = __ffma2_rn(x2, make_float2(0.79788456f * 0.044715f, 0.79788456f * 0.044715f),
                               make_float2(0.79788456f, 0.79788456f));
  const float2 u = __fmul2_rn(x, pp);
  const float2 th = make_float2(tanh_approx(u.x), tanh_approx(u.y));
  const float2 om = __ffma2_rn(make_float2(-th.x, -th.y), th, make_float2(1.f, 1.f));            // 1 - t^2
  const float2 q = __ffma2_rn(x2, make_float2(0.1070322243f, 0.1070322243f), make_float2(0.79788456f, 0.79788456f));
  const float2 w = __fmul2_rn(__fmul2_rn(x, make_float2(0.5f, 0.5f)), om);                        // 0.5 x (1 - t^2)
  const float2 hh = __ffma2_rn(th, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));            // 0.5 (1 + t)
  return __ffma2_rn(w, q, hh);
}

// gelu_tanh(x) and its derivative from one tanh (forward epilogue that saves the derivative for the backward)
__device__ __forceinline__ void gelu_tanh_both2(float2 x, float2& y, float2& g) {
  const float2 x2 = __fmul2_rn(x, x);
  const float2 pp = __ffma2_rn(x2, make_float2(0.79788456f * 0.044715f, 0.79788456f * 0.044715f),
                               make_float2(0.79788456f, 0.79788456f));
  const float2 u = __fmul2_rn(x, pp);
  const float2 th = make_float2(tanh_approx(u.x), tanh_approx(u.y));
  const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  y = __ffma2_rn(hx, th, hx);
  const float2 om = __ffma2_rn(make_float2(-th.x, -th.y), th, make_float2(1.f, 1.f));
  const float2 q = __ffma2_rn(x2, make_float2(0.1070322243f, 0.1070322243f), make_float2(0.79788456f, 0.79788456f));
  const float2 hh = __ffma2_rn(th, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
  g = __ffma2_rn(__fmul2_rn(hx, om), q, hh);
}

// per-element operand of one chunk (row m, columns [n0, n0+32)): f32 residual = 8 x 16 B, bf16 saved
// pre-activation = 4 x 16 B
template <int EPIK>
__device__ __forceinline__ void epi_fast_aux_load(const EpiParams& e, int m, int n0, uint4 (&x)[8]) {
  if constexpr (EPIK == EPIK_RES_F32) {
    if (e.res_coalesced) {
      // lane -> (row (lane >> 2) + 8k, 16-byte piece lane & 3) of each 64-byte half h: 8 rows x 64 B per
      // instruction (8 L1 wavefronts) instead of 32 rows x 16 B (32 wavefronts) for the row-per-thread form
      const int lane = threadIdx.x & 31, m_base = m - lane, piece = lane & 3;
      const float* base = reinterpret_cast<const float*>(e.residual) + n0 + 4 * piece;
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int mm = m_base + (lane >> 2) + 8 * k;
          x[4 * h + k] = mm < e.M ? *reinterpret_cast<const uint4*>(base + (int64_t)mm * e.ldr + 16 * h)
                                  : make_uint4(0u, 0u, 0u, 0u);
        }
    } else if (m < e.M) {
      const uint4* q = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(e.residual) + (int64_t)m * e.ldr + n0);
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = q[i];
    }
  } else if constexpr (EPIK == EPIK_ACTGRAD) {
    if (e.res_coalesced) {
      // same lane -> (row, piece) pattern for the 64-byte bf16 rows of the saved pre-activation; epi_fast_chunk
      // brings each row back to its owner thread through the staging tile
      const int lane = threadIdx.x & 31, m_base = m - lane, piece = lane & 3;
      const uint16_t* base = reinterpret_cast<const uint16_t*>(e.actgrad_src) + n0 + 8 * piece;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int mm = m_base + (lane >> 2) + 8 * k;
        x[k] = mm < e.M ? *reinterpret_cast<const uint4*>(base + (int64_t)mm * e.ldg) : make_uint4(0u, 0u, 0u, 0u);
      }
    } else if (m < e.M) {
      const uint4* q =
          reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(e.actgrad_src) + (int64_t)m * e.ldg + n0);
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = q[i];
    }
  }
}

// Running softmax statistics (log2 domain: max2 = max * log2e, sum2 = sum 2^(v * log2e - max2)) of the bf16-ROUNDED
// values of one row chunk — exactly the numbers the cross-entropy kernel would read back from the stored logits.
__device__ __forceinline__ void row_stats_update(const float (&t)[32], float2& ms) {
  constexpr float L2E = 1.4426950408889634f;
  float v[32];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float2 f = unpack_bf16x2(pack_bf16x2(t[2 * j], t[2 * j + 1]));
    v[2 * j] = f.x; v[2 * j + 1] = f.y;
  }
  float mx = v[0];
#pragma unroll
  for (int j = 1; j < 32; ++j) mx = fmaxf(mx, v[j]);
  mx *= L2E;
  if (mx > ms.x) {
    float sc;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(sc) : "f"(ms.x - mx));
    ms.y *= sc;  // (-inf - mx) -> 0 on the first chunk
    ms.x = mx;
  }
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float ex;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fmaf(v[j], L2E, -ms.x)));
    acc += ex;
  }
  ms.y += acc;
}

template <int EPIK>
__device__ __forceinline__ void epi_fast_chunk(const EpiParams& e, int m_base, int lane, int n0, int n_in_tile,
                                               const uint32_t (&r)[32], const uint4 (&x)[8], uint32_t stage,
                                               uint32_t bias_s, float2* ms = nullptr) {
  float t[32];
  if (e.bias) {  // CTA-uniform
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 bv;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bv.x), "=f"(bv.y), "=f"(bv.z), "=f"(bv.w)
                   : "r"(bias_s + 4 * (n_in_tile + 4 * i)));
      const float2 a = __fadd2_rn(make_float2(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])),
                                  make_float2(bv.x, bv.y));
      const float2 c = __fadd2_rn(make_float2(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])),
                                  make_float2(bv.z, bv.w));
      t[4 * i] = a.x; t[4 * i + 1] = a.y; t[4 * i + 2] = c.x; t[4 * i + 3] = c.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) t[j] = __uint_as_float(r[j]);
  }
  if constexpr (EPIK == EPIK_BF16_STATS) row_stats_update(t, *ms);
  if constexpr (EPIK == EPIK_GELU_PRE) {
    if (e.act == ACT_GELU_TANH_SAVE_GRAD) {  // CTA-uniform: the backward wants gelu'(t), one tanh serves both
      float g[32];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float2 y, d;
        gelu_tanh_both2(make_float2(t[2 * j], t[2 * j + 1]), y, d);
        t[2 * j] = y.x; t[2 * j + 1] = y.y;
        g[2 * j] = d.x; g[2 * j + 1] = d.y;
      }
      stage_store32(stage, lane, g, DT_BF16, e.preact, e.ldp, m_base, e.M, n0);
    } else {
      stage_store32(stage, lane, t, DT_BF16, e.preact, e.ldp, m_base, e.M, n0);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float2 y = gelu_tanh2(make_float2(t[2 * j], t[2 * j + 1]));
        t[2 * j] = y.x; t[2 * j + 1] = y.y;
      }
    }
  }
  if constexpr (EPIK == EPIK_ACTGRAD) {
    uint4 xr[4];
    if (e.res_coalesced) {  // CTA-uniform: (row, piece) pieces -> staging tile -> this thread's own row
      const int piece = lane & 3;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = (lane >> 2) + 8 * k;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + r * EPI_STAGE_ROW + 16 * (piece ^ ((r >> 1) & 3))),
                     "r"(x[k].x), "r"(x[k].y), "r"(x[k].z), "r"(x[k].w) : "memory");
      }
      __syncwarp();
      const int key = (lane >> 1) & 3;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(xr[i].x), "=r"(xr[i].y), "=r"(xr[i].z), "=r"(xr[i].w)
                     : "r"(stage + lane * EPI_STAGE_ROW + 16 * (i ^ key)));
      __syncwarp();  // the output staging below reuses the tile
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) xr[i] = x[i];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t w[4] = {xr[i].x, xr[i].y, xr[i].z, xr[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 s2 = unpack_bf16x2(w[k]);
        const float2 g = e.actgrad_act == ACT_GRAD_PRECOMPUTED ? s2 : gelu_tanh_grad2(s2);  // CTA-uniform
        const float2 y = __fmul2_rn(make_float2(t[8 * i + 2 * k], t[8 * i + 2 * k + 1]), g);
        t[8 * i + 2 * k] = y.x; t[8 * i + 2 * k + 1] = y.y;
      }
    }
  }
  if constexpr (EPIK == EPIK_RES_F32) {
    if (e.res_coalesced) {  // CTA-uniform
      stage_store32_f32_res(stage, lane, t, e.C, e.ldc, m_base, e.M, n0, x);
      return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 a = __fadd2_rn(make_float2(t[4 * i], t[4 * i + 1]),
                                  make_float2(__uint_as_float(x[i].x), __uint_as_float(x[i].y)));
      const float2 c = __fadd2_rn(make_float2(t[4 * i + 2], t[4 * i + 3]),
                                  make_float2(__uint_as_float(x[i].z), __uint_as_float(x[i].w)));
      t[4 * i] = a.x; t[4 * i + 1] = a.y; t[4 * i + 2] = c.x; t[4 * i + 3] = c.y;
    }
  }
  stage_store32(stage, lane, t, EPIK == EPIK_RES_F32 ? DT_F32 : DT_BF16, e.C, e.ldc, m_base, e.M, n0);
}

// =================================================================================================
// tcgen05 kernel
// =================================================================================================
constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int GEMM_THREADS = 64 + 32 * 8;  // TMA warp, MMA warp, 8 epilogue warps
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB

template <int BN>
struct GemmCfg {  // (EPI_STAGE_BYTES is defined above)
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;  // double-buffered accumulator
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 /*barriers*/ +
                                    EPI_WARPS * EPI_STAGE_BYTES /*epilogue transposition buffers*/ +
                                    2 * 256 * 4 /*double-buffered bias slice of the tile*/;
};

struct TcParams {
  int M, N, K;
  int a_mn, b_mn;
  int fmt;       // 0 f16, 1 bf16
  int split_k;   // number of K splits (>=1)
  int group_m;   // rasterisation group (m-tiles per group)
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const TcParams p, const EpiParams e) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment (the budget has no slack for rounding up)
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0u) __trap();
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + STAGES * A_STAGE_BYTES;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t full_bar = bar_base;                     // STAGES x 8 B
  const uint32_t empty_bar = bar_base + 8 * STAGES;       // STAGES x 8 B
  const uint32_t tfull_bar = bar_base + 16 * STAGES;      // 2 x 8 B
  const uint32_t tempty_bar = tfull_bar + 16;             // 2 x 8 B
  const uint32_t tmem_slot = tempty_bar + 16;             // 4 B
  const uint32_t epi_stage = bar_base + 256;               // 4 x EPI_STAGE_BYTES
  const uint32_t bias_smem = epi_stage + EPI_WARPS * EPI_STAGE_BYTES;  // 2 x 256 floats
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int k_blocks_total = (p.K + BK - 1) / BK;
  const int kb_per_split = (k_blocks_total + p.split_k - 1) / p.split_k;
  const int num_work = m_tiles * n_tiles * p.split_k;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);
      mbar_init(tempty_bar + 8 * a, 32 * EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();               // programmatic dependent launch: see ct_common.cuh
  pdl_launch_dependents();

  // work item -> (m_blk, n_blk, split): m fastest inside groups of `group_m` m-tiles so that the
  // CTAs resident at one time share a handful of A row-panels and B column-panels in L2.
  auto decode = [&](int w, int& m_blk, int& n_blk, int& split) {
    split = w % p.split_k;
    int t = w / p.split_k;
    const int per_group = p.group_m * n_tiles;
    const int grp = t / per_group;
    const int in_grp = t - grp * per_group;
    const int m_first = grp * p.group_m;
    const int gm = min(p.group_m, m_tiles - m_first);
    m_blk = m_first + in_grp % gm;
    n_blk = in_grp / gm;
  };

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        int m_blk, n_blk, split;
        decode(w, m_blk, n_blk, split);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, k_blocks_total);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty_bar + 8 * s, ph ^ 1);
          mbar_expect_tx(full_bar + 8 * s, Cfg::STAGE_BYTES);
          const uint32_t a_dst = sA + s * A_STAGE_BYTES;
          const uint32_t b_dst = sB + s * Cfg::B_STAGE_BYTES;
          if (!p.a_mn) {
            tma_load_2d(a_dst, &tmA, full_bar + 8 * s, kb * BK, m_blk * BM);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_2d(a_dst + j * (64 * BK * 2), &tmA, full_bar + 8 * s, m_blk * BM + 64 * j,
                          kb * BK);
          }
          if (!p.b_mn) {
            tma_load_2d(b_dst, &tmB, full_bar + 8 * s, kb * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_2d(b_dst + j * (64 * BK * 2), &tmB, full_bar + 8 * s, n_blk * BN + 64 * j,
                          kb * BK);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ====================================== MMA issuer ======================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(p.fmt, p.a_mn, p.b_mn, BM, BN);
      // K-major : 8-row x 128 B swizzle atoms stacked along M/N (SBO = 1024 B); a K=16 step moves
      //           the start address by 32 B inside the atom.
      // MN-major: atoms are 64 (M/N) x 8 (K); SBO = 1024 B between 8-row K groups, LBO = BK*128 B
      //           between 64-wide M/N chunks; a K=16 step moves the start address by 2048 B.
      const uint32_t a_lbo = p.a_mn ? (BK * 128) : 0, b_lbo = p.b_mn ? (BK * 128) : 0;
      const uint32_t a_kstep = p.a_mn ? 2048 : 32, b_kstep = p.b_mn ? 2048 : 32;
      uint32_t it = 0;
      uint32_t local = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++local) {
        int m_blk, n_blk, split;
        decode(w, m_blk, n_blk, split);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, k_blocks_total);
        const uint32_t acc = local & 1;
        const uint32_t acc_ph = (local >> 1) & 1;
        CT_GDBG(32, 2048 + 8 * local + 0);
        mbar_wait(tempty_bar + 8 * acc, acc_ph ^ 1);
        CT_GDBG(32, 2048 + 8 * local + 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(full_bar + 8 * s, ph);
          tc_fence_after();
          const uint32_t a_addr = sA + s * A_STAGE_BYTES;
          const uint32_t b_addr = sB + s * Cfg::B_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = umma_smem_desc_sw128(a_addr + k * a_kstep, a_lbo, 1024);
            const uint64_t db = umma_smem_desc_sw128(b_addr + k * b_kstep, b_lbo, 1024);
            umma_f16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar + 8 * s);  // smem stage reusable once these MMAs retire
        }
        umma_commit(tfull_bar + 8 * acc);  // accumulator complete -> epilogue
        CT_GDBG(32, 2048 + 8 * local + 2);
      }
    }
  } else {
    // ======================================= epilogue =======================================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t local = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++local) {
      int m_blk, n_blk, split;
      decode(w, m_blk, n_blk, split);
      const int kb0 = split * kb_per_split;
      const int kb1 = min(kb0 + kb_per_split, k_blocks_total);
      const uint32_t acc = local & 1;
      const uint32_t acc_ph = (local >> 1) & 1;
      CT_GDBG(64, 8 * local + 0);
      mbar_wait(tfull_bar + 8 * acc, acc_ph);
      CT_GDBG(64, 8 * local + 1);
      tc_fence_after();
      const int m_base = m_blk * BM + q * 32;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      const int ew = warp - 2;                 // 0..7
      const int hf = ew >> 2;                  // even / odd 32-column chunks
      const uint32_t stage = epi_stage + (uint32_t)ew * EPI_STAGE_BYTES;
      const uint32_t bias_s = bias_smem + (local & 1) * (256 * 4);
      const int et = ew * 32 + lane;           // 0..255 among the epilogue threads
      if (e.bias) {
        // stage this tile's bias slice (double-buffered by tile parity: the single barrier below also
        // orders the previous-but-one tile's reads before this write)
        if (et < BN) {
          const int n = n_blk * BN + et;
          const float bv = n < e.N ? __ldg(e.bias + n) : 0.f;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + 4 * et), "f"(bv) : "memory");
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      if (kb1 > kb0) {
        // the TMEM load of this warp's next chunk is in flight while the current one is processed
        const float aux[32] = {};
        uint32_t r[32];
        tmem_ld_32x32(t_row + hf * 32, r);
#pragma unroll 1
        for (int c = hf; c < BN / 32; c += 2) {
          float t[32];
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = __uint_as_float(r[j]);
          if (c + 2 < BN / 32) tmem_ld_32x32(t_row + (c + 2) * 32, r);
          epi_block32(e, m_base, lane, n_blk * BN + c * 32, c * 32, t, aux, AUX_NONE, split == 0, stage, bias_s);
        }
      }
      CT_GDBG(64, 8 * local + 2);
      tc_fence_before();
      mbar_arrive(tempty_bar + 8 * acc);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}


// =================================================================================================
// 2-CTA variant: a cluster of two CTAs computes a 256 x 256 tile with tcgen05.mma.cta_group::2.
// Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 columns): 32 KB per
// K block per SM instead of 48 KB, i.e. one third less L2->SM traffic, which is what bounds the
// 1-CTA main loop (measured 645 cycles per K block against the 512-cycle tensor-pipe floor).
// CTA 0 (leader) issues the MMAs; TMA completions of both CTAs are credited to the leader's full
// barrier; tcgen05.commit multicasts to the barriers of both CTAs; both CTAs run the epilogue on
// their own 128 accumulator rows and release the accumulator on the leader's barrier.
// =================================================================================================
constexpr int BN2 = 256;
constexpr int STAGES2 = 6;
constexpr int STAGE2_BYTES = A_STAGE_BYTES + (BN2 / 2) * BK * 2;  // 32 KB per CTA
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 256 + EPI_WARPS * EPI_STAGE_BYTES + 2 * 256 * 4;

template <int EPIK>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
    gemm_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                             const TcParams p, const EpiParams e) {
  constexpr int BN = BN2;
  constexpr int STAGES = STAGES2;
  constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0u) __trap();
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + STAGES * A_STAGE_BYTES;
  const uint32_t bar_base = smem_base + STAGES * STAGE2_BYTES;
  const uint32_t full_bar = bar_base;
  const uint32_t empty_bar = bar_base + 8 * STAGES;
  const uint32_t tfull_bar = bar_base + 16 * STAGES;
  const uint32_t tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  const uint32_t epi_stage = bar_base + 256;
  const uint32_t bias_smem = epi_stage + EPI_WARPS * EPI_STAGE_BYTES;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  const int m_tiles = (p.M + 255) / 256;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int k_blocks_total = (p.K + BK - 1) / BK;
  const int kb_per_split = (k_blocks_total + p.split_k - 1) / p.split_k;
  const int num_work = m_tiles * n_tiles * p.split_k;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);
      mbar_init(tempty_bar + 8 * a, 2 * EPI_WARPS);  // one elected arrive per epilogue warp of both CTAs
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits + TMEM allocation visible to the peer before any remote access
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // everything above touched shared memory, TMEM and kernel parameters only: it may run under the previous kernel's tail
  pdl_wait();               // programmatic dependent launch: see ct_common.cuh
  pdl_launch_dependents();

  auto decode = [&](int w, int& m_blk, int& n_blk, int& split) {
    split = w % p.split_k;
    int t = w / p.split_k;
    const int gmx = max(1, p.group_m / 2);
    const int per_group = gmx * n_tiles;
    const int grp = t / per_group;
    const int in_grp = t - grp * per_group;
    const int m_first = grp * gmx;
    const int gm = min(gmx, m_tiles - m_first);
    m_blk = m_first + in_grp % gm;
    n_blk = in_grp / gm;
  };

  if (warp == 0) {
    // ============================== TMA producer (both CTAs) ==============================
    if (lane == 0) {
      const uint32_t leader_full = mapa_shared(full_bar, 0);
      uint32_t it = 0;
      for (int w = pair; w < num_work; w += num_pairs) {
        int m_blk, n_blk, split;
        decode(w, m_blk, n_blk, split);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, k_blocks_total);
        const int m0 = m_blk * 256 + (int)rank * 128;
        const int n0 = n_blk * BN + (int)rank * (BN / 2);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty_bar + 8 * s, ph ^ 1);
          if (leader) mbar_expect_tx(full_bar + 8 * s, 2 * STAGE2_BYTES);
          const uint32_t a_dst = sA + s * A_STAGE_BYTES;
          const uint32_t b_dst = sB + s * B_HALF_BYTES;
          const uint32_t bar = leader_full + 8 * s;
          if (!p.a_mn) {
            tma_load_2d_2cta(a_dst, &tmA, bar, kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d_2cta(a_dst + j * (64 * BK * 2), &tmA, bar, m0 + 64 * j, kb * BK);
          }
          if (!p.b_mn) {
            tma_load_2d_2cta(b_dst, &tmB, bar, kb * BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d_2cta(b_dst + j * (64 * BK * 2), &tmB, bar, n0 + 64 * j, kb * BK);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader only) ================================
    if (leader && lane == 0) {
      const uint32_t idesc = umma_idesc_f16(p.fmt, p.a_mn, p.b_mn, 256, BN);
      const uint32_t a_lbo = p.a_mn ? (BK * 128) : 0, b_lbo = p.b_mn ? (BK * 128) : 0;
      const uint32_t a_kstep = p.a_mn ? 2048 : 32, b_kstep = p.b_mn ? 2048 : 32;
      uint32_t it = 0;
      uint32_t local = 0;
      for (int w = pair; w < num_work; w += num_pairs, ++local) {
        int m_blk, n_blk, split;
        decode(w, m_blk, n_blk, split);
        const int kb0 = split * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, k_blocks_total);
        const uint32_t acc = local & 1;
        const uint32_t acc_ph = (local >> 1) & 1;
        mbar_wait(tempty_bar + 8 * acc, acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(full_bar + 8 * s, ph);
          tc_fence_after();
          const uint32_t a_addr = sA + s * A_STAGE_BYTES;
          const uint32_t b_addr = sB + s * B_HALF_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = umma_smem_desc_sw128(a_addr + k * a_kstep, a_lbo, 1024);
            const uint64_t db = umma_smem_desc_sw128(b_addr + k * b_kstep, b_lbo, 1024);
            umma_f16_2cta(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_2cta(empty_bar + 8 * s);  // frees the stage in both CTAs
        }
        umma_commit_2cta(tfull_bar + 8 * acc);  // accumulator complete -> both epilogues
      }
    }
  } else {
    // ================================ epilogue (both CTAs) ================================
    const int q = warp & 3;
    const uint32_t leader_tempty = mapa_shared(tempty_bar, 0);
    uint32_t local = 0;
    for (int w = pair; w < num_work; w += num_pairs, ++local) {
      int m_blk, n_blk, split;
      decode(w, m_blk, n_blk, split);
      const int kb0 = split * kb_per_split;
      const int kb1 = min(kb0 + kb_per_split, k_blocks_total);
      const uint32_t acc = local & 1;
      const uint32_t acc_ph = (local >> 1) & 1;
      mbar_wait(tfull_bar + 8 * acc, acc_ph);
      tc_fence_after();
      const int m_base = m_blk * 256 + (int)rank * 128 + q * 32;
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      const int ew = warp - 2;
      const int hf = ew >> 2;
      const uint32_t stage = epi_stage + (uint32_t)ew * EPI_STAGE_BYTES;
      const uint32_t bias_s = bias_smem + (local & 1) * (256 * 4);
      const int et = ew * 32 + lane;
      if (e.bias) {
        const int n = n_blk * BN + et;
        const float bv = n < e.N ? __ldg(e.bias + n) : 0.f;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_s + 4 * et), "f"(bv) : "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      if constexpr (EPIK == EPIK_GENERIC) {
        if (kb1 > kb0) {
          const float aux[32] = {};
          uint32_t r[32];
          tmem_ld_32x32(t_row + hf * 32, r);
#pragma unroll 1
          for (int c = hf; c < BN / 32; c += 2) {
            float t[32];
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = __uint_as_float(r[j]);
            if (c + 2 < BN / 32) tmem_ld_32x32(t_row + (c + 2) * 32, r);
            epi_block32(e, m_base, lane, n_blk * BN + c * 32, c * 32, t, aux, AUX_NONE, split == 0, stage, bias_s);
          }
        }
      } else {
        // specialised: 4 chunks per warp, TMEM loads and per-element operand loads one chunk ahead,
        // alternating between two register sets (fully unrolled: no copies)
        constexpr int NCH = BN / 64;
        uint32_t ra[32], rb[32];
        uint4 xa[8], xb[8];
        const int m = m_base + lane;
        const int nt0 = n_blk * BN;
        float2 ms = make_float2(-INFINITY, 0.f);  // EPIK_BF16_STATS: this thread's row over its four chunks
        tmem_ld_32x32(t_row + hf * 32, ra);
        if (nt0 + hf * 32 < e.N) epi_fast_aux_load<EPIK>(e, m, nt0 + hf * 32, xa);
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
          const int c = hf + 2 * k;
          tmem_ld_wait();
          if (k + 1 < NCH) {
            tmem_ld_32x32(t_row + (c + 2) * 32, (k & 1) ? ra : rb);
            if (nt0 + (c + 2) * 32 < e.N) epi_fast_aux_load<EPIK>(e, m, nt0 + (c + 2) * 32, (k & 1) ? xa : xb);
          }
          if (nt0 + c * 32 < e.N)  // warp-uniform (N % 32 == 0 on this path)
            epi_fast_chunk<EPIK>(e, m_base, lane, nt0 + c * 32, c * 32, (k & 1) ? rb : ra, (k & 1) ? xb : xa, stage,
                                 bias_s, &ms);
        }
        if constexpr (EPIK == EPIK_BF16_STATS) {
          // slot (2 * tile + hf) of row m; [slot][row] layout: a warp writes 32 consecutive rows (256 contiguous bytes).
          // A half tile entirely beyond N keeps (-inf, 0), which the merge in the loss kernel ignores.
          if (m < e.M)
            reinterpret_cast<float2*>(e.row_stats)[(int64_t)(2 * n_blk + hf) * e.M + m] = ms;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tempty + 8 * acc);
    }
  }

  tc_fence_before();
  cluster_sync_all();  // nobody exits (or frees TMEM) while the peer can still touch its smem / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

// =================================================================================================
// SIMT fallback (64x64 tile, 256 threads, 4x4 outputs per thread)
// =================================================================================================
__global__ void __launch_bounds__(256)
    gemm_simt_kernel(const void* __restrict__ A, int64_t lda, int a_mn, const void* __restrict__ B,
                     int64_t ldb, int b_mn, int ab_dtype, int K, const EpiParams e) {
  __shared__ float As[16][65];
  __shared__ float Bs[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      int mm, kk;
      if (a_mn) { mm = i & 63; kk = i >> 6; } else { kk = i & 15; mm = i >> 4; }
      float v = 0.f;
      if (m0 + mm < e.M && k0 + kk < K)
        v = ld_elem(A, ab_dtype, a_mn ? (int64_t)(k0 + kk) * lda + (m0 + mm)
                                      : (int64_t)(m0 + mm) * lda + (k0 + kk));
      As[kk][mm] = v;
      int nn;
      if (b_mn) { nn = i & 63; kk = i >> 6; } else { kk = i & 15; nn = i >> 4; }
      v = 0.f;
      if (n0 + nn < e.N && k0 + kk < K)
        v = ld_elem(B, ab_dtype, b_mn ? (int64_t)(k0 + kk) * ldb + (n0 + nn)
                                      : (int64_t)(n0 + nn) * ldb + (k0 + kk));
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < e.M && n < e.N) epi_scalar(e, m, n, acc[i][j], true);
    }
}

// =================================================================================================
// Skinny GEMM (M <= 32 rows): the q_len = 1 decode step of GenerationMixin._greedy_search (generation_util.py:57-119)
// multiplies 32 tokens by every weight matrix once per emitted token, so the weights are streamed from HBM and nothing
// else matters. A 128-row tcgen05 tile wastes the launch on a handful of CTAs (N / 256 of them); here a
// CTA owns 8*G output features, its 8 warps split K in interleaved 32-element chunks, every lane reads whole 16-byte
// pieces of a weight row (and of the token rows), and the products run on mma.sync m16n8k16 with the TOKENS as the
// 16-row operand. A lane's 16 bytes are 8 consecutive k; the fragment layout wants (2q, 2q+1) and (2q+8, 2q+9) per lane,
// which is only a permutation of k applied to both operands alike, so no shuffle or shared-memory transpose is needed.
// The 8 partial sums are added in warp order (deterministic), then the common scalar epilogue runs.
// A [M,K] and B [N,K] both K-major, K % 32 == 0, 16-byte aligned rows.
// =================================================================================================
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1, bool bf16) {
  if (bf16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// SK_WARPS warps split K: 8, or 16 for the narrowest tile (G = 1: few CTAs, so each one should keep more loads in flight)
// BF16 / LEAN are compile-time so that a launch fetches only the code it runs (a decode step is ~100 of these kernels,
// each a few microseconds long: instruction-cache misses at kernel start showed as stall_no_instruction ~ 1 per issue).
// LEAN: epilogue = alpha, bias, activation, residual, store; otherwise the common scalar epilogue (epi_scalar).
template <int G, int SK_WARPS, bool BF16, bool LEAN>
__global__ void __launch_bounds__(SK_WARPS * 32, SK_WARPS == 8 ? 2 : 1)
    gemm_skinny_kernel(const uint16_t* __restrict__ A, int64_t lda, const uint16_t* __restrict__ B, int64_t ldb,
                       int K, const EpiParams e) {
  constexpr int FT = 8 * G;  // output features per CTA
  __shared__ float red[SK_WARPS][G * 8][32];
  pdl_wait();               // programmatic dependent launch: see ct_common.cuh
  pdl_launch_dependents();
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int n0 = blockIdx.x * FT;
  const uint16_t* wrow[G];
#pragma unroll
  for (int j = 0; j < G; ++j) wrow[j] = B + (int64_t)min(n0 + 8 * j + g, e.N - 1) * ldb + q * 8;
  const uint16_t* xrow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) xrow[i] = A + (int64_t)min(g + 8 * i, e.M - 1) * lda + q * 8;
  // the epilogue's operands are requested before the weight stream so that they do not add a second memory round trip
  constexpr int EPT = (32 * FT + SK_WARPS * 32 - 1) / (SK_WARPS * 32);  // output elements per thread
  constexpr bool lean = LEAN;
  float ep_bias[EPT], ep_res[EPT];
#pragma unroll
  for (int it = 0; it < EPT; ++it) {
    const int idx = threadIdx.x + it * SK_WARPS * 32;
    const int tok = idx / FT, n = n0 + idx % FT;
    const bool ok = idx < 32 * FT && tok < e.M && n < e.N;
    ep_bias[it] = (lean && ok && e.bias) ? __ldg(e.bias + n) : 0.f;
    ep_res[it] = (lean && ok && e.residual) ? ld_elem(e.residual, e.res_dtype, (int64_t)tok * e.ldr + n) : 0.f;
  }
  float acc[G][2][4];
#pragma unroll
  for (int j = 0; j < G; ++j)
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[j][t][c] = 0.f;
  const int chunks = K >> 5;
  constexpr int U = G >= 2 ? 2 : 4;  // chunks in flight per warp: (G + 4) * U 16-byte loads per lane
#pragma unroll
  for (int j = 0; j < G; ++j) wrow[j] += wib * 32;  // pointers walk with the loop: constant offsets inside an iteration
#pragma unroll
  for (int i = 0; i < 4; ++i) xrow[i] += wib * 32;
  for (int c0 = wib; c0 < chunks; c0 += SK_WARPS * U) {
    uint4 w[U][G], x[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (c0 + u * SK_WARPS < chunks) {
#pragma unroll
        for (int j = 0; j < G; ++j) w[u][j] = ld_stream16(wrow[j] + u * SK_WARPS * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) x[u][i] = __ldg(reinterpret_cast<const uint4*>(xrow[i] + u * SK_WARPS * 32));
      }
    }
#pragma unroll
    for (int j = 0; j < G; ++j) wrow[j] += SK_WARPS * U * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) xrow[i] += SK_WARPS * U * 32;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (c0 + u * SK_WARPS < chunks) {
#pragma unroll
        for (int j = 0; j < G; ++j) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            mma16816(acc[j][t], x[u][2 * t].x, x[u][2 * t + 1].x, x[u][2 * t].y, x[u][2 * t + 1].y, w[u][j].x, w[u][j].y,
                     BF16);
            mma16816(acc[j][t], x[u][2 * t].z, x[u][2 * t + 1].z, x[u][2 * t].w, x[u][2 * t + 1].w, w[u][j].z, w[u][j].w,
                     BF16);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < G; ++j)
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int c = 0; c < 4; ++c) red[wib][(j * 2 + t) * 4 + c][lane] = acc[j][t][c];
  __syncthreads();
  // element (token, feature) of the CTA tile lives in fragment slot (j, t, c) of lane (gg, qq)
#pragma unroll
  for (int it = 0; it < EPT; ++it) {
    const int idx = threadIdx.x + it * SK_WARPS * 32;
    const int tok = idx / FT, f = idx % FT;
    if (idx >= 32 * FT || tok >= e.M || n0 + f >= e.N) continue;
    const int j = f >> 3, qq = (f & 7) >> 1, t = tok >> 4, gg = tok & 7;
    const int c = (((tok >> 3) & 1) << 1) | (f & 1);
    const int slot = (j * 2 + t) * 4 + c, ln = gg * 4 + qq;
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < SK_WARPS; ++wv) v += red[wv][slot][ln];
    if constexpr (LEAN) {  // bias, activation, residual: the decode step's epilogues
      v = fmaf(e.alpha, v, ep_bias[it]);
      if (e.act != ACT_NONE) v = act_apply_call(v, e.act);
      st_elem(e.C, e.c_dtype, (int64_t)tok * e.ldc + n0 + f, v + ep_res[it]);
    } else {
      epi_scalar(e, tok, n0 + f, v, true);
    }
  }
}

static bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

template <int BN>
static int launch_tc(const ct_gemm_args& a, const EpiParams& e, int split_k, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2], strides[2];
    uint32_t box[2];
    if (!a.a_mn_major) { dims[0] = a.K; dims[1] = a.M; box[0] = BK; box[1] = BM; }
    else               { dims[0] = a.M; dims[1] = a.K; box[0] = 64; box[1] = BK; }
    strides[0] = 2; strides[1] = (uint64_t)a.lda * 2;
    int r = make_tmap(&tmA, a.A, 2, 2, dims, strides, box, 1);
    if (r) return r;
    if (!a.b_mn_major) { dims[0] = a.K; dims[1] = a.N; box[0] = BK; box[1] = BN; }
    else               { dims[0] = a.N; dims[1] = a.K; box[0] = 64; box[1] = BK; }
    strides[1] = (uint64_t)a.ldb * 2;
    r = make_tmap(&tmB, a.B, 2, 2, dims, strides, box, 1);
    if (r) return r;
  }
  TcParams p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.a_mn = a.a_mn_major; p.b_mn = a.b_mn_major;
  p.fmt = a.ab_dtype == DT_BF16 ? 1 : 0;
  p.split_k = split_k;
  p.group_m = 16;
  static bool attr_set = false;  // benign race: idempotent
  if (!attr_set) {
    CT_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int m_tiles = (a.M + BM - 1) / BM, n_tiles = (a.N + BN - 1) / BN;
  int64_t work = (int64_t)m_tiles * n_tiles * split_k;
  int grid = sm_count();
  if (work < grid) grid = (int)work;
  CT_CUDA_OK(launch_k(gemm_tcgen05_kernel<BN>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, st, tmA, tmB, p, e));
  CT_LAUNCH_OK();
  return 0;
}


template <int EPIK>
static int launch_2cta_kind(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p, const EpiParams& e,
                            int pairs, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    CT_CUDA_OK(cudaFuncSetAttribute(gemm_tcgen05_2cta_kernel<EPIK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    SMEM2_BYTES));
    attr_set = true;
  }
  CT_CUDA_OK(launch_k(gemm_tcgen05_2cta_kernel<EPIK>, dim3(2 * pairs), dim3(GEMM_THREADS), SMEM2_BYTES, st, tmA, tmB, p, e));
  CT_LAUNCH_OK();
  return 0;
}

// which compile-time epilogue covers this call (EPIK_GENERIC when none does)
static int pick_epi_kind(const ct_gemm_args& a, const EpiParams& e) {
  if (option(OPT_GEMM_EPI_IMPL) == 1) return EPIK_GENERIC;
  if (!e.vec_ok || e.atomic_out || a.alpha != 1.f || a.beta != 0.f || (a.N % 32) != 0) return EPIK_GENERIC;
  const bool plain = a.act == ACT_NONE && !a.preact && !a.actgrad_src && !a.residual;
  if (plain && a.c_dtype == DT_BF16 && a.ab_dtype == DT_BF16) return e.row_stats ? EPIK_BF16_STATS : EPIK_BF16;
  if ((a.act == ACT_GELU_TANH || a.act == ACT_GELU_TANH_SAVE_GRAD) && a.preact && a.preact_dtype == DT_BF16 &&
      a.c_dtype == DT_BF16 && !a.actgrad_src &&
      !a.residual)
    return EPIK_GELU_PRE;
  if (a.act == ACT_NONE && !a.preact && !a.actgrad_src && a.residual && a.res_dtype == DT_F32 && a.c_dtype == DT_F32)
    return EPIK_RES_F32;
  if (a.act == ACT_NONE && !a.preact && a.actgrad_src && a.actgrad_dtype == DT_BF16 &&
      (a.actgrad_act == ACT_GELU_TANH || a.actgrad_act == ACT_GRAD_PRECOMPUTED) &&
      !a.residual && a.c_dtype == DT_BF16)
    return EPIK_ACTGRAD;
  return EPIK_GENERIC;
}

static int launch_tc_2cta(const ct_gemm_args& a, const EpiParams& e, int split_k, cudaStream_t st) {
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2], strides[2];
    uint32_t box[2];
    if (!a.a_mn_major) { dims[0] = a.K; dims[1] = a.M; box[0] = BK; box[1] = 128; }
    else               { dims[0] = a.M; dims[1] = a.K; box[0] = 64; box[1] = BK; }
    strides[0] = 2; strides[1] = (uint64_t)a.lda * 2;
    int r = make_tmap(&tmA, a.A, 2, 2, dims, strides, box, 1);
    if (r) return r;
    if (!a.b_mn_major) { dims[0] = a.K; dims[1] = a.N; box[0] = BK; box[1] = BN2 / 2; }
    else               { dims[0] = a.N; dims[1] = a.K; box[0] = 64; box[1] = BK; }
    strides[1] = (uint64_t)a.ldb * 2;
    r = make_tmap(&tmB, a.B, 2, 2, dims, strides, box, 1);
    if (r) return r;
  }
  TcParams p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.a_mn = a.a_mn_major; p.b_mn = a.b_mn_major;
  p.fmt = a.ab_dtype == DT_BF16 ? 1 : 0;
  p.split_k = split_k;
  p.group_m = 16;
  const int m_tiles = (a.M + 255) / 256, n_tiles = (a.N + BN2 - 1) / BN2;
  int64_t work = (int64_t)m_tiles * n_tiles * split_k;
  int pairs = sm_count() / 2;
  if (work < pairs) pairs = (int)work;
  switch (split_k == 1 ? pick_epi_kind(a, e) : EPIK_GENERIC) {
    case EPIK_BF16: return launch_2cta_kind<EPIK_BF16>(tmA, tmB, p, e, pairs, st);
    case EPIK_BF16_STATS: return launch_2cta_kind<EPIK_BF16_STATS>(tmA, tmB, p, e, pairs, st);
    case EPIK_GELU_PRE: return launch_2cta_kind<EPIK_GELU_PRE>(tmA, tmB, p, e, pairs, st);
    case EPIK_RES_F32: return launch_2cta_kind<EPIK_RES_F32>(tmA, tmB, p, e, pairs, st);
    case EPIK_ACTGRAD: return launch_2cta_kind<EPIK_ACTGRAD>(tmA, tmB, p, e, pairs, st);
    default: return launch_2cta_kind<EPIK_GENERIC>(tmA, tmB, p, e, pairs, st);
  }
}

}  // namespace ct

using namespace ct;

extern "C" int ct_gemm(const ct_gemm_args* args, void* stream) {
  CT_REQUIRE(args != nullptr, CT_ERR_BAD_ARG, "ct_gemm: null args");
  const ct_gemm_args& a = *args;
  CT_REQUIRE(a.A && a.B && a.C, CT_ERR_BAD_ARG, "ct_gemm: null operand");
  CT_REQUIRE(a.M >= 0 && a.N >= 0 && a.K >= 0, CT_ERR_BAD_ARG, "ct_gemm: negative dim");
  CT_REQUIRE(a.ab_dtype == DT_BF16 || a.ab_dtype == DT_F16, CT_ERR_UNSUPPORTED,
             "ct_gemm: A/B must be bf16 or f16");
  CT_REQUIRE(a.c_dtype >= 0 && a.c_dtype <= 2, CT_ERR_UNSUPPORTED, "ct_gemm: bad c_dtype");
  CT_REQUIRE(a.beta == 0.f || a.beta == 1.f, CT_ERR_BAD_ARG, "ct_gemm: beta must be 0 or 1");
  CT_REQUIRE(a.beta == 0.f || a.c_dtype == DT_F32, CT_ERR_UNSUPPORTED,
             "ct_gemm: beta=1 requires f32 C");
  CT_REQUIRE(a.act >= 0 && a.act <= ACT_GELU_TANH_SAVE_GRAD && a.actgrad_act >= 0 &&
                 a.actgrad_act <= ACT_GRAD_PRECOMPUTED && a.actgrad_act != ACT_GELU_TANH_SAVE_GRAD &&
                 (a.act != ACT_GELU_TANH_SAVE_GRAD || a.preact != nullptr),
             CT_ERR_BAD_ARG,
             "ct_gemm: bad activation enum");
  CT_REQUIRE(a.lda >= (a.a_mn_major ? a.M : a.K) && a.ldb >= (a.b_mn_major ? a.N : a.K) &&
                 a.ldc >= a.N,
             CT_ERR_BAD_ARG, "ct_gemm: leading dimension too small");
  if (a.M == 0 || a.N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;

  EpiParams e;
  e.M = a.M; e.N = a.N;
  e.C = a.C; e.c_dtype = a.c_dtype; e.ldc = a.ldc;
  e.alpha = a.alpha; e.beta = a.beta;
  e.bias = a.bias; e.act = a.act;
  e.preact = a.preact; e.preact_dtype = a.preact_dtype; e.ldp = a.ldp;
  e.actgrad_src = a.actgrad_src; e.actgrad_dtype = a.actgrad_dtype; e.actgrad_act = a.actgrad_act;
  e.ldg = a.ldg;
  e.residual = a.residual; e.res_dtype = a.res_dtype; e.ldr = a.ldr;
  e.atomic_out = 0;
  e.res_coalesced = option(OPT_GEMM_EPI_IMPL) != 2;  // GEMM_EPI_IMPL 2 = row-per-thread residual loads (first form)
  e.row_stats = a.row_stats;
  e.vec_ok = al16(a.C) && (a.ldc % 8 == 0) && (!a.bias || al16(a.bias)) &&
             (!a.preact || (al16(a.preact) && a.ldp % 8 == 0)) &&
             (!a.actgrad_src || (al16(a.actgrad_src) && a.ldg % 8 == 0)) &&
             (!a.residual || (al16(a.residual) && a.ldr % 8 == 0));

  const bool tma_ok = al16(a.A) && al16(a.B) && (a.lda % 8 == 0) && (a.ldb % 8 == 0) && a.K > 0;
  bool use_tc;
  if (a.impl == 1) {
    CT_REQUIRE(tma_ok, CT_ERR_UNSUPPORTED,
               "ct_gemm: tcgen05 path needs 16-byte aligned A/B and lda/ldb multiples of 8");
    use_tc = true;
  } else if (a.impl == 3) {
    CT_REQUIRE(tma_ok, CT_ERR_UNSUPPORTED, "ct_gemm: 2-CTA path needs 16-byte aligned A/B and lda/ldb multiples of 8");
    use_tc = true;
  } else if (a.impl == 2) {
    use_tc = false;
  } else {
    use_tc = tma_ok && ((int64_t)a.M * a.N * a.K >= (1 << 18));
  }

  // M <= 32 (one decode step): weight-streaming kernel on mma.sync; impl 4 forces it, impl 0 picks it when the layout fits
  const bool skinny_ok = a.M <= 32 && !a.a_mn_major && !a.b_mn_major && tma_ok && (a.K % 32 == 0) && !a.row_stats;
  CT_REQUIRE(a.impl != 4 || skinny_ok, CT_ERR_UNSUPPORTED,
             "ct_gemm: the skinny kernel needs M <= 32, K-major A and B, K %% 32 == 0, 16-byte aligned rows");
  // (a very wide, 16-byte addressable output — Bloom's 250 880-column LM head — is still served better by the 128-row
  // tcgen05 kernel: 108 vs 229 us at M = 32, profiles/r02o_skinny.json; GPT-2's 50 257 columns are not addressable that
  // way and take the skinny kernel: 57 vs 138 us)
  const bool very_wide = a.N >= 131072 && (a.N % 8 == 0) && (a.ldc % 8 == 0) && al16(a.C);
  if (skinny_ok && (a.impl == 4 || (a.impl == 0 && !very_wide && (int64_t)a.N * a.K >= (1 << 16)))) {
    const uint16_t* A16 = (const uint16_t*)a.A;
    const uint16_t* B16 = (const uint16_t*)a.B;
    const bool bf = a.ab_dtype == DT_BF16;
    const bool lean = !a.preact && !a.actgrad_src && a.beta == 0.f;
    const int sms2 = sm_count();
    // features per CTA: wide tiles amortise the token rows (read once per CTA) when there are CTAs to spare
    const int g = ((a.N + 31) / 32 >= 2 * sms2) ? 4 : ((a.N + 15) / 16 >= sms2) ? 2 : 1;
#define CT_SK_GO(G, W, BF, LEAN) \
  CT_CUDA_OK(launch_k(gemm_skinny_kernel<G, W, BF, LEAN>, dim3((unsigned)((a.N + 8 * G - 1) / (8 * G))), dim3(W * 32), 0, st, \
                      A16, a.lda, B16, a.ldb, a.K, e))
#define CT_SK_PICK(G, W)                                  \
  do {                                                    \
    if (bf && lean) CT_SK_GO(G, W, true, true);           \
    else if (bf) CT_SK_GO(G, W, true, false);             \
    else if (lean) CT_SK_GO(G, W, false, true);           \
    else CT_SK_GO(G, W, false, false);                    \
  } while (0)
    if (g == 4) CT_SK_PICK(4, 8);
    else if (g == 2) CT_SK_PICK(2, 8);
    else CT_SK_PICK(1, 16);
#undef CT_SK_PICK
#undef CT_SK_GO
    CT_LAUNCH_OK();
    return 0;
  }
  CT_REQUIRE(use_tc || !a.row_stats, CT_ERR_UNSUPPORTED, "ct_gemm: row_stats needs the tcgen05 2-CTA kernel");
  if (!use_tc) {
    if (a.K == 0) {
      // degenerate: C = epilogue(0)
    }
    dim3 grid((a.N + 63) / 64, (a.M + 63) / 64);
    gemm_simt_kernel<<<grid, 256, 0, st>>>(a.A, a.lda, a.a_mn_major, a.B, a.ldb, a.b_mn_major,
                                           a.ab_dtype, a.K, e);
    CT_LAUNCH_OK();
    return 0;
  }

  const int bn = (a.N > 128) ? 256 : 128;
  const int m_tiles = (a.M + BM - 1) / BM, n_tiles = (a.N + bn - 1) / bn;
  const int k_blocks = (a.K + BK - 1) / BK;
  // split-K only for linear f32-accumulate epilogues (wgrad): partial sums go out via red.add
  int split_k = 1;
  const bool linear = a.act == ACT_NONE && !a.preact && !a.actgrad_src && !a.residual &&
                      a.c_dtype == DT_F32;
  const int tiles = m_tiles * n_tiles;
  const int sms = sm_count();
  if (linear && tiles < sms && k_blocks >= 16) {
    split_k = (2 * sms + tiles - 1) / tiles;
    const int max_split = k_blocks / 8;  // at least 8 k-blocks (K=512) per split
    if (split_k > max_split) split_k = max_split;
    if (split_k < 1) split_k = 1;
    // make sure no split is empty
    const int per = (k_blocks + split_k - 1) / split_k;
    split_k = (k_blocks + per - 1) / per;
  }
  const int auto_2cta = option(OPT_GEMM_2CTA);
  const bool use_2cta = a.impl == 3 || (a.impl == 0 && auto_2cta && a.M >= 512 && a.N >= 256);
  // row statistics exist only in one compile-time epilogue of the 2-CTA kernel: refuse rather than skip them silently
  CT_REQUIRE(!a.row_stats || (use_2cta && pick_epi_kind(a, e) == EPIK_BF16_STATS), CT_ERR_UNSUPPORTED,
             "ct_gemm: row_stats needs the 2-CTA kernel's plain bf16 epilogue (M >= 512, N >= 256, N %% 32 == 0, "
             "16-byte aligned, no bias-free restriction, alpha 1, beta 0)");
  if (use_2cta) {
    // recompute the split for 256 x 256 tiles
    const int t2 = ((a.M + 255) / 256) * ((a.N + 255) / 256);
    int sk = 1;
    if (linear && t2 < sms / 2 && k_blocks >= 16) {
      if (option(OPT_GEMM_SPLITK) == 1) {  // first-generation rule: about two work units per CTA pair
        sk = (sms + t2 - 1) / t2;
        const int max_split = k_blocks / 8;
        if (sk > max_split) sk = max_split;
        if (sk < 1) sk = 1;
        const int per = (k_blocks + sk - 1) / sk;
        sk = (k_blocks + per - 1) / per;
      } else {
        // Wave quantisation decides: t2 * sk work units run in ceil(units / pairs) rounds of ceil(k_blocks / sk)
        // K blocks each, plus an epilogue per round (~2 K-block times for plain stores, ~4 when the partial sums
        // leave through red.global.add, which also costs the memset of C). Bloom-560M wgrads on 74 pairs:
        // 4h<->h (64 tiles) 3 -> 1 split (no atomics at all), QKV (48 tiles) 4 -> 3, h->h (16 tiles) 10 -> 4.
        const int pairs = sms / 2;
        const int max_split = k_blocks / 8 < 32 ? k_blocks / 8 : 32;
        long best_cost = -1;
        for (int c = 1; c <= (max_split < 1 ? 1 : max_split); ++c) {
          const int per = (k_blocks + c - 1) / c;
          if ((k_blocks + per - 1) / per != c) continue;  // same partition as a smaller candidate
          const long rounds = ((long)t2 * c + pairs - 1) / pairs;
          const long cost = rounds * (per + (c > 1 ? 4 : 2));
          if (best_cost < 0 || cost < best_cost) { best_cost = cost; sk = c; }
        }
      }
    }
    if (sk > 1) {
      e.atomic_out = 1;
      if (a.beta == 0.f)
        CT_CUDA_OK(cudaMemset2DAsync(a.C, (size_t)a.ldc * 4, 0, (size_t)a.N * 4, (size_t)a.M, st));
    }
    return launch_tc_2cta(a, e, sk, st);
  }
  if (split_k > 1) {
    e.atomic_out = 1;
    if (a.beta == 0.f)
      CT_CUDA_OK(cudaMemset2DAsync(a.C, (size_t)a.ldc * 4, 0, (size_t)a.N * 4, (size_t)a.M, st));
  }
  return bn == 256 ? launch_tc<256>(a, e, split_k, st) : launch_tc<128>(a, e, split_k, st);
}

static void init_args(ct_gemm_args& g) {
  memset(&g, 0, sizeof(g));
  g.alpha = 1.f;
}

extern "C" int ct_gemm_bias_act(const void* x, const void* w, int w_in_out, const float* bias,
                                const void* residual, int res_dtype, void* y, int y_dtype,
                                void* preact, int act, int64_t M, int64_t N, int64_t K,
                                int ab_dtype, void* stream) {
  ct_gemm_args g;
  init_args(g);
  g.M = (int)M; g.N = (int)N; g.K = (int)K;
  g.ab_dtype = ab_dtype;
  g.A = x; g.lda = K; g.a_mn_major = 0;
  g.B = w; g.b_mn_major = w_in_out ? 1 : 0; g.ldb = w_in_out ? N : K;
  g.C = y; g.c_dtype = y_dtype; g.ldc = N;
  g.bias = bias; g.act = act;
  g.preact = preact; g.preact_dtype = ab_dtype; g.ldp = N;
  g.residual = residual; g.res_dtype = res_dtype; g.ldr = N;
  return ct_gemm(&g, stream);
}

extern "C" int ct_gemm_dgrad(const void* dy, const void* w, int w_in_out, void* dx, int dx_dtype,
                             const void* actgrad_src, int actgrad_act, int64_t M, int64_t N,
                             int64_t K, int ab_dtype, void* stream) {
  // dx[M,K] = dy[M,N] @ W  -> GEMM with (M, N'=K, K'=N); B(n'=k, k'=n) = W[n,k]
  ct_gemm_args g;
  init_args(g);
  g.M = (int)M; g.N = (int)K; g.K = (int)N;
  g.ab_dtype = ab_dtype;
  g.A = dy; g.lda = N; g.a_mn_major = 0;
  // nn.Linear W [N,K]: row = reduction index n, contiguous = output index k -> MN-major.
  // Conv1D  W [K,N]: row = output index k, contiguous = reduction n      -> K-major.
  g.B = w; g.b_mn_major = w_in_out ? 0 : 1; g.ldb = w_in_out ? N : K;
  g.C = dx; g.c_dtype = dx_dtype; g.ldc = K;
  g.actgrad_src = actgrad_src; g.actgrad_dtype = ab_dtype; g.actgrad_act = actgrad_act; g.ldg = K;
  return ct_gemm(&g, stream);
}

extern "C" int ct_gemm_wgrad_bias(const void* dy, const void* x, int w_in_out, float* dw,
                                  float* db, int accumulate, int64_t M, int64_t N, int64_t K,
                                  int ab_dtype, void* stream) {
  ct_gemm_args g;
  init_args(g);
  g.ab_dtype = ab_dtype;
  g.c_dtype = DT_F32;
  g.beta = accumulate ? 1.f : 0.f;
  if (!w_in_out) {
    // dW[N,K] = dy^T[N,M] @ x[M,K]: A(n,m) = dy[m,n] (MN-major), B(k,m) = x[m,k] (MN-major)
    g.M = (int)N; g.N = (int)K; g.K = (int)M;
    g.A = dy; g.lda = N; g.a_mn_major = 1;
    g.B = x; g.ldb = K; g.b_mn_major = 1;
    g.C = dw; g.ldc = K;
  } else {
    // Conv1D: dW[K,N] = x^T[K,M] @ dy[M,N]
    g.M = (int)K; g.N = (int)N; g.K = (int)M;
    g.A = x; g.lda = K; g.a_mn_major = 1;
    g.B = dy; g.ldb = N; g.b_mn_major = 1;
    g.C = dw; g.ldc = N;
  }
  int r = ct_gemm(&g, stream);
  if (r) return r;
  if (db) return ct_colsum(dy, ab_dtype, N, db, accumulate, M, N, stream);
  return 0;
}

#ifdef CT_DEBUG_TIMING
extern "C" int ct_debug_timing_gemm(long long* out, int n) {
  if (n > 4096) n = 4096;
  return (int)cudaMemcpyFromSymbol(out, ct::ct_dbg_gemm, sizeof(long long) * n);
}
#endif
