// attention.cu — fused QK^T -> (+ALiBi / masks) -> online softmax -> PV, forward and backward.
//
// Replaces the materialised [B,H,Sq,Sk] score pipeline of
//   CleanTransformer/models/modeling_bloom.py:99-116, modeling_gpt.py:83-103, transformer.py:41-57
// (see include/ct_b200.h for the exact score definition shared by all three variants).
//
// tcgen05 forward  (head_dim 64): CTA = 128 query rows of one (b,h); warp 0 TMA producer (Q once, K/V
//   double-buffered 128-key tiles), warp 1 single-thread MMA issuer (S = Q K^T into one of two TMEM
//   score buffers; O_tile = P V into the score buffer that was just drained), warps 2..5 softmax: one
//   thread per query row (no shuffles), two TMEM passes (max, then exp2 + P -> swizzled smem),
//   running O kept in registers. 2 CTAs/SM (112 KB smem, 256 TMEM columns each) so one CTA's softmax
//   overlaps the other's MMAs.
// tcgen05 backward (head_dim 64): CTA = 128 keys of one (b,h), loops over query tiles:
//   S^T = K Q^T, dP^T = V dO^T (TMEM) -> P^T, dS^T (bf16, swizzled smem) -> dV += P^T dO,
//   dK += dS^T Q (TMEM accumulators), dQ_tile = dS K -> red.global.add.f32 into an fp32 dQ workspace.
//   The Q/dO tiles are read through two descriptor views (K-major for the first pair of MMAs,
//   MN-major for the second), dS^T likewise (K-major for dK, MN-major for dQ): no transposes.
// SIMT kernels: any head_dim <= 128 and tiny shapes (golden-vector tests, q_len = 1 decode).
//
// FLOPs: forward 4*Sq*Sk*D per (b,h) dense (half that under the causal mask); backward 2.5x.
#include "ct_common.cuh"
#include "../../include/ct_b200.h"
#include <cfloat>
#include <cstring>
#include <type_traits>

namespace ct {

constexpr float LOG2E = 1.4426950408889634f;

#ifdef CT_DEBUG_TIMING
__device__ long long ct_dbg_clk[4096];
#define CT_DBG_STAMP(slot) do { if (blockIdx.x == CT_DBG_BLOCK && threadIdx.x == CT_DBG_THREAD && (slot) < 4096) ct_dbg_clk[(slot)] = clock64(); } while (0)
// whole-CTA timeline of every 64th block (slots 1024 + 8 * (block / 64) + k): clock64 for k < 6, globaltimer (ns) at
// entry / exit in slots 6 / 7
__device__ __forceinline__ long long ct_dbg_gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define CT_DBG_CTA(k) do { if ((blockIdx.x & 63) == 60 && threadIdx.x == CT_DBG_THREAD && blockIdx.x < 64 * 128) { \
    ct_dbg_clk[1024 + 8 * (blockIdx.x >> 6) + (k)] = clock64(); \
    if ((k) == 0) ct_dbg_clk[1024 + 8 * (blockIdx.x >> 6) + 6] = ct_dbg_gtime(); \
    if ((k) == 5) ct_dbg_clk[1024 + 8 * (blockIdx.x >> 6) + 7] = ct_dbg_gtime(); } } while (0)
#else
#define CT_DBG_STAMP(slot) do {} while (0)
#define CT_DBG_CTA(k) do {} while (0)
#endif
#ifndef CT_DBG_BLOCK
#define CT_DBG_BLOCK 700
#define CT_DBG_THREAD 64
#endif

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bar_sync_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct AttnP {
  int B, H, Sq, Sk;
  int fmt;  // 0 f16, 1 bf16
  float sl2;           // scale * log2e
  float scale;
  int causal;
  float causal_fill2;  // causal_fill * log2e (may be -inf)
  int off;             // Sk - Sq
  const float* kbias2; int64_t kb_sb, kb_sh;
  const int32_t* first_valid;
  void* o; int64_t o_sb, o_sh, o_ss;
  float* lse2;
};

// score in the log2 domain for element (query i, key j) given the raw dot product
__device__ __forceinline__ float score2(float acc, float sl2, float kb, bool future, float cf2,
                                        bool oob) {
  float v = future ? (cf2 + kb) : fmaf(acc, sl2, kb);
  v = fmaxf(v, -FLT_MAX);
  return oob ? -INFINITY : v;
}

// =================================================================================================
// tcgen05 forward
// =================================================================================================
constexpr int FA_THREADS = 192;
constexpr int FA_TILE = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16, SWIZZLE_128B
constexpr int FA_SMEM = FA_TILE /*Q*/ + 2 * FA_TILE /*K*/ + 2 * FA_TILE /*V*/ + 2 * FA_TILE /*P*/ + 128 /*barriers*/ +
                        512 /*per-key bias of the current tile*/;

__global__ void __launch_bounds__(FA_THREADS, 2)
    attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const AttnP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sQ = base;
  const uint32_t sK = base + FA_TILE;
  const uint32_t sV = base + 3 * FA_TILE;
  const uint32_t sP = base + 5 * FA_TILE;
  const uint32_t bars = base + 7 * FA_TILE;
  const uint32_t q_full = bars, k_full = bars + 8, v_full = bars + 24, kv_empty = bars + 40,
                 s_full = bars + 56, p_ready = bars + 72, o_full = bars + 80, tmem_slot = bars + 88,
                 kb_s = bars + 128;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem + 7 * FA_TILE + 88);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_q_tiles = (p.Sq + 127) / 128;
  // heavy (late) query tiles first: better tail under the causal mask
  const int q_tile = n_q_tiles - 1 - (int)(blockIdx.x % n_q_tiles);
  const int bh = blockIdx.x / n_q_tiles;
  const int h = bh % p.H, b = bh / p.H;
  const int q0 = q_tile * 128;

  int n_kv = (p.Sk + 127) / 128;
  if (p.causal) {
    const bool full_sweep = p.first_valid && (q0 + p.off < p.first_valid[b]);
    if (!full_sweep) {
      const int last_key = min(p.Sk - 1, q0 + 127 + p.off);
      n_kv = last_key < 0 ? 0 : last_key / 128 + 1;
    }
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full + 8 * s, 1); mbar_init(v_full + 8 * s, 1); mbar_init(kv_empty + 8 * s, 1);
      mbar_init(s_full + 8 * s, 1);
    }
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0 && n_kv > 0) {
      mbar_expect_tx(q_full, FA_TILE);
      tma_load_4d(sQ, &tmQ, q_full, 0, q0, h, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        mbar_wait(kv_empty + 8 * s, ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(k_full + 8 * s, FA_TILE);
        tma_load_4d(sK + s * FA_TILE, &tmK, k_full + 8 * s, 0, j * 128, h, b);
        mbar_expect_tx(v_full + 8 * s, FA_TILE);
        tma_load_4d(sV + s * FA_TILE, &tmV, v_full + 8 * s, 0, j * 128, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_kv > 0) {
      const uint32_t idesc_s = umma_idesc_f16(p.fmt, 0, 0, 128, 128);
      const uint32_t idesc_o = umma_idesc_f16(p.fmt, 0, 1, 128, 64);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(k_full + 8 * s, (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem + (j & 1) * 128, umma_smem_desc_sw128(sQ + k * 32, 0, 1024),
                   umma_smem_desc_sw128(sK + s * FA_TILE + k * 32, 0, 1024), idesc_s, k > 0);
        umma_commit(s_full + 8 * (j & 1));
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        mbar_wait(p_ready, j & 1);
        mbar_wait(v_full + 8 * s, (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tmem + (j & 1) * 128,
                   umma_smem_desc_sw128(sP + (k >> 2) * FA_TILE + (k & 3) * 32, 0, 1024),
                   umma_smem_desc_sw128(sV + s * FA_TILE + k * 2048, 64 * 128, 1024), idesc_o, k > 0);
        umma_commit(kv_empty + 8 * s);
        umma_commit(o_full);
        if (j + 1 < n_kv) issue_s(j + 1);
      }
    }
  } else {
    // ------------------------------ softmax: one thread per query row ------------------------------
    const int qr = (warp & 3) * 32 + lane;  // row inside the tile == TMEM lane
    const int i = q0 + qr;
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const float* kb_row = p.kbias2 ? p.kbias2 + (int64_t)b * p.kb_sb + (int64_t)h * p.kb_sh : nullptr;
    const bool has_kb = kb_row != nullptr;
    float o_acc[64];
#pragma unroll
    for (int d = 0; d < 64; ++d) o_acc[d] = 0.f;
    float m = -INFINITY, l = 0.f;
    const uint32_t p_row = sP + qr * 128;
    const int sw = qr & 7;
    // per-key bias of the current tile, staged through smem: one global load per thread per tile,
    // prefetched one tile ahead (a global load per element exposed ~L2 latency 64x per tile)
    float kb_next = (has_kb && qr < p.Sk) ? __ldg(kb_row + qr) : 0.f;

    auto kb4 = [&](int col) -> float4 {  // 4 consecutive staged bias values (broadcast LDS.128)
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "r"(kb_s + 4 * col));
      return v;
    };

    for (int j = 0; j < n_kv; ++j) {
      const int kv0 = j * 128;
      const uint32_t t_s = t_lane + (j & 1) * 128;
      CT_DBG_STAMP(2048 + 16 * j + 0);
      mbar_wait(s_full + 8 * (j & 1), (j >> 1) & 1);
      CT_DBG_STAMP(2048 + 16 * j + 1);
      tc_fence_after();
      // a tile needs per-element masking if it touches the causal diagonal or the ragged key edge
      const bool slow = (p.causal && (kv0 + 127 > q0 + p.off)) || (kv0 + 128 > p.Sk);
      if (has_kb) {
        // every thread is past o_full(j-1), i.e. all 128 have finished reading the previous tile's bias
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(kb_s + 4 * qr), "f"(kb_next) : "memory");
        bar_sync_named(1, 128);
        const int nj = kv0 + 128 + qr;
        kb_next = (j + 1 < n_kv && nj < p.Sk) ? __ldg(kb_row + nj) : 0.f;
      }
      // Both passes run as ROLLED loops over the four 32-column chunks (one body each for the
      // masked and the mask-free tile kinds): the first fully unrolled version was 7k SASS
      // instructions and spent most of its time stalled on instruction fetch. The TMEM load of the
      // next chunk is issued before the math of the current one (register copy = 32 MOVs).
      uint32_t r[32];
      float cur[32];
      float mt = -INFINITY;
      auto val = [&](float a, float kb, int jg, auto slow_tag) -> float {
        if constexpr (decltype(slow_tag)::value)
          return score2(a, p.sl2, kb, p.causal && (jg > i + p.off), p.causal_fill2, jg >= p.Sk);
        else
          return fmaf(a, p.sl2, kb);
      };
      auto pass1 = [&](auto slow_tag) {
        tmem_ld_32x32(t_s, r);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) cur[t] = __uint_as_float(r[t]);
          tmem_ld_32x32(t_s + ((c + 1) & 3) * 32, r);  // c == 3: first chunk of pass 2
#pragma unroll
          for (int t4 = 0; t4 < 8; ++t4) {
            float4 k4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_kb) k4 = kb4(c * 32 + t4 * 4);
            const int jg = kv0 + c * 32 + t4 * 4;
            mt = fmaxf(mt, val(cur[t4 * 4 + 0], k4.x, jg + 0, slow_tag));
            mt = fmaxf(mt, val(cur[t4 * 4 + 1], k4.y, jg + 1, slow_tag));
            mt = fmaxf(mt, val(cur[t4 * 4 + 2], k4.z, jg + 2, slow_tag));
            mt = fmaxf(mt, val(cur[t4 * 4 + 3], k4.w, jg + 3, slow_tag));
          }
        }
      };
      if (slow) pass1(std::true_type{}); else pass1(std::false_type{});
      CT_DBG_STAMP(2048 + 16 * j + 2);
      mt = fmaxf(mt, -FLT_MAX);  // clamp once: max(clamp(x)) == clamp(max(x))
      const float m_new = fmaxf(m, mt);
      const float alpha = ex2(m - m_new);
      float lt = 0.f;
      // ---- pass 2: p = 2^(s - m), write bf16 P into the K-major SWIZZLE_128B tile ----
      auto pass2 = [&](auto slow_tag) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) cur[t] = __uint_as_float(r[t]);
          if (c < 3) tmem_ld_32x32(t_s + (c + 1) * 32, r);
#pragma unroll
          for (int g = 0; g < 4; ++g) {  // 4 x 16-byte chunks (8 values each); panel = c / 2
            float pv[8];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int col = c * 32 + g * 8 + hh * 4;
              float4 k4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (has_kb) k4 = kb4(col);
              const float kb[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                float v = val(cur[g * 8 + hh * 4 + u], kb[u], kv0 + col + u, slow_tag);
                if constexpr (!decltype(slow_tag)::value) v = fmaxf(v, -FLT_MAX);
                const float e = ex2(v - m_new);
                lt += e;
                pv[hh * 4 + u] = e;
              }
            }
            uint32_t w0, w1, w2, w3;
            if (p.fmt == 1) {
              w0 = pack_bf16x2(pv[0], pv[1]); w1 = pack_bf16x2(pv[2], pv[3]);
              w2 = pack_bf16x2(pv[4], pv[5]); w3 = pack_bf16x2(pv[6], pv[7]);
            } else {
              __half2 h0 = __floats2half2_rn(pv[0], pv[1]), h1 = __floats2half2_rn(pv[2], pv[3]);
              __half2 h2 = __floats2half2_rn(pv[4], pv[5]), h3 = __floats2half2_rn(pv[6], pv[7]);
              w0 = *reinterpret_cast<uint32_t*>(&h0); w1 = *reinterpret_cast<uint32_t*>(&h1);
              w2 = *reinterpret_cast<uint32_t*>(&h2); w3 = *reinterpret_cast<uint32_t*>(&h3);
            }
            const int chunk = (c & 1) * 4 + g;
            const uint32_t addr = p_row + (c >> 1) * FA_TILE + ((chunk ^ sw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1),
                         "r"(w2), "r"(w3)
                         : "memory");
          }
        }
      };
      if (slow) pass2(std::true_type{}); else pass2(std::false_type{});
      CT_DBG_STAMP(2048 + 16 * j + 3);
      l = l * alpha + lt;
      m = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
      // ---- O = O * alpha + P V ----
      mbar_wait(o_full, j & 1);
      CT_DBG_STAMP(2048 + 16 * j + 4);
      tc_fence_after();
      {
        uint32_t r2[32];
        tmem_ld_32x32(t_s, r);
        tmem_ld_32x32(t_s + 32, r2);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          o_acc[t] = fmaf(o_acc[t], alpha, __uint_as_float(r[t]));
          o_acc[32 + t] = fmaf(o_acc[32 + t], alpha, __uint_as_float(r2[t]));
        }
      }
      tc_fence_before();
    }
    // ---- epilogue: normalise, write O (merged-head layout) and lse2 ----
    if (i < p.Sq) {
      const float inv = (n_kv > 0) ? 1.f / l : 0.f;
      uint8_t* orow = reinterpret_cast<uint8_t*>(p.o) +
                      2 * ((int64_t)b * p.o_sb + (int64_t)h * p.o_sh + (int64_t)i * p.o_ss);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 w;
        if (p.fmt == 1) {
          w.x = pack_bf16x2(o_acc[8 * g] * inv, o_acc[8 * g + 1] * inv);
          w.y = pack_bf16x2(o_acc[8 * g + 2] * inv, o_acc[8 * g + 3] * inv);
          w.z = pack_bf16x2(o_acc[8 * g + 4] * inv, o_acc[8 * g + 5] * inv);
          w.w = pack_bf16x2(o_acc[8 * g + 6] * inv, o_acc[8 * g + 7] * inv);
        } else {
          __half2 h0 = __floats2half2_rn(o_acc[8 * g] * inv, o_acc[8 * g + 1] * inv);
          __half2 h1 = __floats2half2_rn(o_acc[8 * g + 2] * inv, o_acc[8 * g + 3] * inv);
          __half2 h2 = __floats2half2_rn(o_acc[8 * g + 4] * inv, o_acc[8 * g + 5] * inv);
          __half2 h3 = __floats2half2_rn(o_acc[8 * g + 6] * inv, o_acc[8 * g + 7] * inv);
          w.x = *reinterpret_cast<uint32_t*>(&h0); w.y = *reinterpret_cast<uint32_t*>(&h1);
          w.z = *reinterpret_cast<uint32_t*>(&h2); w.w = *reinterpret_cast<uint32_t*>(&h3);
        }
        *reinterpret_cast<uint4*>(orow + 16 * g) = w;
      }
      if (p.lse2) p.lse2[((int64_t)b * p.H + h) * p.Sq + i] = (n_kv > 0) ? m + log2f(l) : -INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

// =================================================================================================
// tcgen05 forward, v2
// =================================================================================================
// Same CTA shape as above (128 query rows of one (b,h), TMA warp + MMA warp + 4 softmax warps, two CTAs
// per SM) but the softmax side was rebuilt around what the ncu capture of v1 showed (19.6 warp
// instructions per score element, issue-bound, S read twice from TMEM):
//   * interior tiles (no causal diagonal, no ragged edge, no masked key) take ONE pass: the whole
//     128-wide score row is loaded into registers, the score buffer is released to the MMA warp at
//     once (so S(j+1) = Q K(j+1)^T runs under the softmax of tile j: one TMEM score buffer is enough),
//     and scale / bias / max / exp2 / sum run on register pairs (fma.f32x2 / add.f32x2);
//   * the running output O lives in TMEM: P(j) V(j) accumulates into it on the tensor core and the
//     softmax threads rescale it in place (tcgen05.ld -> mul -> tcgen05.st) only when a row maximum
//     moved, instead of reading every P V product back into 64 registers per thread;
//   * tiles on the causal diagonal keep the generic two-pass code but skip the 32-column chunks that
//     lie entirely in the future of the warp's 32 rows (they are exactly -FLT_MAX for the Bloom fill).
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// CTA-subset barrier that also ORs a predicate over the participating threads
__device__ __forceinline__ bool bar_red_or(int id, int nthreads, bool pred) {
  uint32_t out;
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.u32 q, %3, 0;\n"
      "bar.red.or.pred p, %1, %2, q;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(out)
      : "r"(id), "r"(nthreads), "r"((uint32_t)pred)
      : "memory");
  return out != 0;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// Running max over one 32-column chunk held in registers. SCALE: r <- r * sl2 (+ kb) in place (always with a
// per-key bias; without one the raw scores are kept and scaled inside the exp2 instead).
template <bool HAS_KB, bool SCALE>
__device__ __forceinline__ float fa2_scale_max(uint32_t (&r)[32], int col0, float sl2, uint32_t kb_s, float mt) {
  static_assert(SCALE || !HAS_KB, "a per-key bias is folded in by scaling in place");
  const float2 s2 = make_float2(sl2, sl2);
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    float2 a = make_float2(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]));
    float2 b = make_float2(__uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
    if constexpr (SCALE) {
      if constexpr (HAS_KB) {
        const float4 k4 = lds128f(kb_s + 4 * (col0 + 4 * g));
        a = __ffma2_rn(a, s2, make_float2(k4.x, k4.y));
        b = __ffma2_rn(b, s2, make_float2(k4.z, k4.w));
      } else {
        a = __fmul2_rn(a, s2);
        b = __fmul2_rn(b, s2);
      }
      r[4 * g] = __float_as_uint(a.x); r[4 * g + 1] = __float_as_uint(a.y);
      r[4 * g + 2] = __float_as_uint(b.x); r[4 * g + 3] = __float_as_uint(b.y);
    }
    mt = fmaxf(mt, fmaxf(a.x, a.y));
    mt = fmaxf(mt, fmaxf(b.x, b.y));
  }
  return mt;
}

// p = 2^(t - m_new) for one 32-column chunk held in registers (SCALED: t = r, else t = r * sl2), bf16/f16 P
// into the swizzled K-major tile
template <bool SCALED, bool BF16>
__device__ __forceinline__ void fa2_exp_store(const uint32_t (&r)[32], int c, float m_new, float sl2, uint32_t p_row,
                                              int sw, float2& acc0, float2& acc1) {
  const float2 nm = make_float2(-m_new, -m_new);
  const float2 s2 = make_float2(sl2, sl2);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float2 a = make_float2(__uint_as_float(r[8 * g + 2 * u]), __uint_as_float(r[8 * g + 2 * u + 1]));
      if constexpr (SCALED) a = __fadd2_rn(a, nm);
      else a = __ffma2_rn(a, s2, nm);
      v[u] = make_float2(ex2(a.x), ex2(a.y));
    }
    acc0 = __fadd2_rn(acc0, v[0]); acc1 = __fadd2_rn(acc1, v[1]);
    acc0 = __fadd2_rn(acc0, v[2]); acc1 = __fadd2_rn(acc1, v[3]);
    uint32_t w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if constexpr (BF16) {
        w[u] = pack_bf16x2(v[u].x, v[u].y);
      } else {
        __half2 h = __floats2half2_rn(v[u].x, v[u].y);
        w[u] = *reinterpret_cast<uint32_t*>(&h);
      }
    }
    const int chunk = (c & 1) * 4 + g;
    const uint32_t addr = p_row + (c >> 1) * FA_TILE + ((chunk ^ sw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                 : "memory");
  }
}

// Tile on an ALIGNED causal diagonal (key tile origin == first visible key of query row 0 of the tile, fill
// -FLT_MAX, no masked / out-of-range key) for the warp owning rows 32*WQ..32*WQ+31: chunks < WQ are fully
// visible, chunk WQ is lower-triangular (column t visible iff t <= lane), chunks > WQ lie entirely in the
// future: every score there is exactly -FLT_MAX, no TMEM read, P = 2^(-FLT_MAX - m) (0 unless the row has
// seen nothing but masked keys). One pass over register-resident rows like the interior tiles.
template <bool HAS_KB, bool BF16, int WQ>
__device__ __forceinline__ void fa2_diag_tile(uint32_t t_s, float sl2, uint32_t kb_s, uint32_t p_row, int sw, int lane,
                                              float m, int j, uint32_t s_free, uint32_t o_full, float& m_new,
                                              float& alpha, float& lt) {
  uint32_t r[WQ + 1][32];
#pragma unroll
  for (int c = 0; c <= WQ; ++c) tmem_ld_32x32(t_s + c * 32, r[c]);
  tmem_ld_wait();
  float mt = -INFINITY;
#pragma unroll
  for (int c = 0; c < WQ; ++c) mt = fa2_scale_max<HAS_KB, true>(r[c], c * 32, sl2, kb_s, mt);
  (void)fa2_scale_max<HAS_KB, true>(r[WQ], WQ * 32, sl2, kb_s, mt);  // scale in place; max taken after masking
  tc_fence_before();
  mbar_arrive(s_free);  // after the last read of the staged bias (see the interior-tile path)
#pragma unroll
  for (int t = 1; t < 32; ++t)
    if (t > lane) r[WQ][t] = __float_as_uint(-FLT_MAX);  // future keys of the diagonal chunk
#pragma unroll
  for (int t = 0; t < 32; t += 2)
    mt = fmaxf(mt, fmaxf(__uint_as_float(r[WQ][t]), __uint_as_float(r[WQ][t + 1])));
  mt = fmaxf(mt, -FLT_MAX);
  m_new = fmaxf(m, mt);
  alpha = ex2(m - m_new);
  if (j > 0) {
    mbar_wait(o_full, (j - 1) & 1);
    tc_fence_after();
  }
  float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0;
#pragma unroll
  for (int c = 0; c <= WQ; ++c) fa2_exp_store<true, BF16>(r[c], c, m_new, sl2, p_row, sw, acc0, acc1);
  lt = (acc0.x + acc0.y) + (acc1.x + acc1.y);
  if constexpr (WQ < 3) {
    const float em = ex2(-FLT_MAX - m_new);
    uint32_t w;
    if constexpr (BF16) {
      w = pack_bf16x2(em, em);
    } else {
      __half2 hx = __floats2half2_rn(em, em);
      w = *reinterpret_cast<uint32_t*>(&hx);
    }
#pragma unroll
    for (int c = WQ + 1; c < 4; ++c) {
      lt += 32.f * em;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int chunk = (c & 1) * 4 + g;
        const uint32_t addr = p_row + (c >> 1) * FA_TILE + ((chunk ^ sw) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(w) : "memory");
      }
    }
  }
}

// LAZY (forward v3, ATTN_FWD_IMPL=2; written after the round's GPU budget was spent — compiled, not yet run):
// the r01f stamps show the slowest softmax warp of every tile waiting ~1.2 k cycles for P(j-1).V(j-1) before it
// may overwrite the single P tile and rescale O. Here (interior tiles)
//   * the running maximum is a REFERENCE that only moves when the tile maximum exceeds it by more than 8 (2^8:
//     P stays within bf16 range, sums stay fp32), so O is rescaled — and P.V waited for — only in those tiles;
//   * P is handed over in its two 64-key panels: panel 0 is rewritten as soon as the first four MMAs of the previous
//     P.V have retired (p0_free) and its MMAs start while the softmax warps still work on panel 1.
// Masked / diagonal tiles keep the exact per-tile maximum and arrive on both panel barriers at the end.
template <bool HAS_KB, bool BF16, bool LAZY = false>
__global__ void __launch_bounds__(FA_THREADS, 2)
    attn_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const AttnP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sQ = base;
  const uint32_t sK = base + FA_TILE;
  const uint32_t sV = base + 3 * FA_TILE;
  const uint32_t sP = base + 5 * FA_TILE;
  const uint32_t bars = base + 7 * FA_TILE;
  const uint32_t q_full = bars, k_full = bars + 8, v_full = bars + 24, kv_empty = bars + 40, s_full = bars + 56,
                 s_free = bars + 64, p_ready = bars + 72, o_full = bars + 80, tmem_slot = bars + 88,
                 p_half1 = bars + 96 /* LAZY: second P panel written */, p0_free = bars + 104 /* LAZY: first P panel read */,
                 kb_s = bars + 128;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem + 7 * FA_TILE + 88);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_q_tiles = (p.Sq + 127) / 128;
  const int q_tile = n_q_tiles - 1 - (int)(blockIdx.x % n_q_tiles);  // heavy (late) query tiles first
  const int bh = blockIdx.x / n_q_tiles;
  const int h = bh % p.H, b = bh / p.H;
  const int q0 = q_tile * 128;

  int n_kv = (p.Sk + 127) / 128;
  if (p.causal) {
    const bool full_sweep = p.first_valid && (q0 + p.off < p.first_valid[b]);
    if (!full_sweep) {
      const int last_key = min(p.Sk - 1, q0 + 127 + p.off);
      n_kv = last_key < 0 ? 0 : last_key / 128 + 1;
    }
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full + 8 * s, 1); mbar_init(v_full + 8 * s, 1); mbar_init(kv_empty + 8 * s, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    mbar_init(p_half1, 128);
    mbar_init(p0_free, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;  // columns [0,128): scores S ; [128,192): running output O

  if (warp == 0) {
    if (lane == 0 && n_kv > 0) {
      mbar_expect_tx(q_full, FA_TILE);
      tma_load_4d(sQ, &tmQ, q_full, 0, q0, h, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        mbar_wait(kv_empty + 8 * s, ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(k_full + 8 * s, FA_TILE);
        tma_load_4d(sK + s * FA_TILE, &tmK, k_full + 8 * s, 0, j * 128, h, b);
        mbar_expect_tx(v_full + 8 * s, FA_TILE);
        tma_load_4d(sV + s * FA_TILE, &tmV, v_full + 8 * s, 0, j * 128, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_kv > 0) {
      const uint32_t idesc_s = umma_idesc_f16(BF16 ? 1 : 0, 0, 0, 128, 128);
      const uint32_t idesc_o = umma_idesc_f16(BF16 ? 1 : 0, 0, 1, 128, 64);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(k_full + 8 * s, (j >> 1) & 1);
        if (j > 0) mbar_wait(s_free, (j - 1) & 1);  // every softmax thread holds S(j-1) in registers
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem, umma_smem_desc_sw128(sQ + k * 32, 0, 1024),
                   umma_smem_desc_sw128(sK + s * FA_TILE + k * 32, 0, 1024), idesc_s, k > 0);
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        if (j + 1 < n_kv) issue_s(j + 1);  // runs under the softmax of tile j
        mbar_wait(p_ready, j & 1);  // LAZY: first panel (keys 0..63 of the tile) only
        mbar_wait(v_full + 8 * s, (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if constexpr (LAZY) {
            if (k == 4) {
              umma_commit(p0_free);
              mbar_wait(p_half1, j & 1);
              tc_fence_after();
            }
          }
          umma_f16(tmem + 128, umma_smem_desc_sw128(sP + (k >> 2) * FA_TILE + (k & 3) * 32, 0, 1024),
                   umma_smem_desc_sw128(sV + s * FA_TILE + k * 2048, 64 * 128, 1024), idesc_o, (j > 0 || k > 0));
        }
        umma_commit(kv_empty + 8 * s);
        umma_commit(o_full);
      }
    }
  } else {
    // ------------------------------ softmax: one thread per query row ------------------------------
    const int wq = warp & 3;
    const int qr = wq * 32 + lane;  // row inside the tile == TMEM lane
    const int i = q0 + qr;
    const uint32_t t_s = tmem + ((uint32_t)(wq * 32) << 16);
    const uint32_t t_o = t_s + 128;
    const float* kb_row = HAS_KB ? p.kbias2 + (int64_t)b * p.kb_sb + (int64_t)h * p.kb_sh : nullptr;
    float m = -INFINITY, l = 0.f;
    const uint32_t p_row = sP + qr * 128;
    const int sw = qr & 7;
    const bool fill_is_ninf = p.causal_fill2 == -INFINITY;
    float kb_next = 0.f;
    if constexpr (HAS_KB) kb_next = (qr < p.Sk) ? __ldg(kb_row + qr) : 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const int kv0 = j * 128;
      CT_DBG_STAMP(2048 + 16 * j + 0);
      mbar_wait(s_full, j & 1);
      CT_DBG_STAMP(2048 + 16 * j + 1);
      tc_fence_after();
      // CTA-uniform tile kind
      const bool touches_diag = p.causal && (kv0 + 127 > q0 + p.off);
      bool irregular = kv0 + 128 > p.Sk;  // ragged key edge, or (below) a masked key: generic path
      if constexpr (HAS_KB) {
        // S(j) needs s_free(j-1), which every thread arrives on after its last read of the previous tile's bias
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(kb_s + 4 * qr), "f"(kb_next) : "memory");
        irregular |= bar_red_or(1, 128, kb_next < -1e30f);
        const int nj = kv0 + 128 + qr;
        kb_next = (j + 1 < n_kv && nj < p.Sk) ? __ldg(kb_row + nj) : 0.f;
      }
      CT_DBG_STAMP(2048 + 16 * j + 2);
      const bool slow = touches_diag || irregular;
      const bool diag_fast = touches_diag && !irregular && fill_is_ninf && kv0 == q0 + p.off;
      float m_new, alpha, lt;
      if (diag_fast) {
        switch (wq) {
          case 0: fa2_diag_tile<HAS_KB, BF16, 0>(t_s, p.sl2, kb_s, p_row, sw, lane, m, j, s_free, o_full, m_new, alpha, lt); break;
          case 1: fa2_diag_tile<HAS_KB, BF16, 1>(t_s, p.sl2, kb_s, p_row, sw, lane, m, j, s_free, o_full, m_new, alpha, lt); break;
          case 2: fa2_diag_tile<HAS_KB, BF16, 2>(t_s, p.sl2, kb_s, p_row, sw, lane, m, j, s_free, o_full, m_new, alpha, lt); break;
          default: fa2_diag_tile<HAS_KB, BF16, 3>(t_s, p.sl2, kb_s, p_row, sw, lane, m, j, s_free, o_full, m_new, alpha, lt); break;
        }
      } else if (!slow) {
        // ---------------- interior tile: one pass over a register-resident score row ----------------
        uint32_t r0[32], r1[32], r2[32], r3[32];
        tmem_ld_32x32(t_s, r0);
        tmem_ld_32x32(t_s + 32, r1);
        tmem_ld_32x32(t_s + 64, r2);
        tmem_ld_32x32(t_s + 96, r3);
        tmem_ld_wait();
        CT_DBG_STAMP(2048 + 16 * j + 3);
        float mt = -INFINITY;
        mt = fa2_scale_max<HAS_KB, HAS_KB>(r0, 0, p.sl2, kb_s, mt);
        mt = fa2_scale_max<HAS_KB, HAS_KB>(r1, 32, p.sl2, kb_s, mt);
        mt = fa2_scale_max<HAS_KB, HAS_KB>(r2, 64, p.sl2, kb_s, mt);
        mt = fa2_scale_max<HAS_KB, HAS_KB>(r3, 96, p.sl2, kb_s, mt);
        // Release the score buffer only after the last read of this tile's staged bias: S(j+1) and with it
        // the next tile's bias staging (single smem buffer) cannot start before every thread got here.
        tc_fence_before();
        mbar_arrive(s_free);
        CT_DBG_STAMP(2048 + 16 * j + 4);
        if constexpr (!HAS_KB) mt *= p.sl2;  // sl2 > 0 (checked on the host)
        mt = fmaxf(mt, -FLT_MAX);
        if constexpr (LAZY) {
          // reference maximum: moves (and O is rescaled, which needs P(j-1) V(j-1)) only when this tile exceeds it by
          // more than 2^8; warp-uniform branch because the TMEM accesses are warp-collective
          const bool need = mt > m + 8.f;
          if (j == 0) {
            m = mt;
          } else if (__any_sync(0xffffffffu, need)) {
            mbar_wait(o_full, (j - 1) & 1);
            tc_fence_after();
            const float m_up = need ? mt : m;
            const float a_up = ex2(m - m_up);
            const float2 a2 = make_float2(a_up, a_up);
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {  // 32 columns at a time: the score row stays in registers beside it
              uint32_t o0[32];
              tmem_ld_32x32(t_o + 32 * hh, o0);
              tmem_ld_wait();
#pragma unroll
              for (int t = 0; t < 16; ++t) {
                float2 x = __fmul2_rn(make_float2(__uint_as_float(o0[2 * t]), __uint_as_float(o0[2 * t + 1])), a2);
                o0[2 * t] = __float_as_uint(x.x); o0[2 * t + 1] = __float_as_uint(x.y);
              }
              tmem_st_32x32(t_o + 32 * hh, o0);
            }
            tmem_st_wait();
            l *= a_up;
            m = m_up;
          }
          m_new = m;
          alpha = 1.f;  // the common tail below: l = l * alpha + lt, no second rescale
          if (j > 0) {  // first four MMAs of P(j-1) V(j-1) retired: panel 0 of the P tile may be overwritten
            mbar_wait(p0_free, (j - 1) & 1);
            tc_fence_after();
          }
          CT_DBG_STAMP(2048 + 16 * j + 5);  // (debug build) 4 -> 5: lazy rescale + wait for panel 0
          float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0;
          fa2_exp_store<HAS_KB, BF16>(r0, 0, m_new, p.sl2, p_row, sw, acc0, acc1);
          fa2_exp_store<HAS_KB, BF16>(r1, 1, m_new, p.sl2, p_row, sw, acc0, acc1);
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(p_ready);  // panel 0: its MMAs run under the rest of this tile
          CT_DBG_STAMP(2048 + 16 * j + 6);  // 5 -> 6: exp + store of panel 0
          if (j > 0) {  // all of P(j-1) V(j-1) retired: panel 1 may be overwritten
            mbar_wait(o_full, (j - 1) & 1);
            tc_fence_after();
          }
          CT_DBG_STAMP(2048 + 16 * j + 7);  // 6 -> 7: wait for panel 1
          fa2_exp_store<HAS_KB, BF16>(r2, 2, m_new, p.sl2, p_row, sw, acc0, acc1);
          fa2_exp_store<HAS_KB, BF16>(r3, 3, m_new, p.sl2, p_row, sw, acc0, acc1);
          lt = (acc0.x + acc0.y) + (acc1.x + acc1.y);
          l += lt;
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(p_half1);
          CT_DBG_STAMP(2048 + 16 * j + 8);  // 7 -> 8: exp + store of panel 1
          continue;  // (the tail below belongs to the exact-maximum paths)
        }
        m_new = fmaxf(m, mt);
        alpha = ex2(m - m_new);
        if (j > 0) {  // P(j-1) V(j-1) retired: the P tile may be overwritten, O may be rescaled
          mbar_wait(o_full, (j - 1) & 1);
          tc_fence_after();
        }
        CT_DBG_STAMP(2048 + 16 * j + 5);
        float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0;
        fa2_exp_store<HAS_KB, BF16>(r0, 0, m_new, p.sl2, p_row, sw, acc0, acc1);
        fa2_exp_store<HAS_KB, BF16>(r1, 1, m_new, p.sl2, p_row, sw, acc0, acc1);
        fa2_exp_store<HAS_KB, BF16>(r2, 2, m_new, p.sl2, p_row, sw, acc0, acc1);
        fa2_exp_store<HAS_KB, BF16>(r3, 3, m_new, p.sl2, p_row, sw, acc0, acc1);
        lt = (acc0.x + acc0.y) + (acc1.x + acc1.y);
      } else {
        // ---------------- masked tile: two passes over TMEM, per-element masking ----------------
        int c_hi = 3;  // last 32-column chunk with a visible element for this warp's rows
        if (p.causal && fill_is_ninf && kv0 == q0 + p.off && kv0 + 128 <= p.Sk) c_hi = wq;
        uint32_t r[32];
        float cur[32];
        float mt = -INFINITY;
        auto val = [&](float a, float kb, int jg) -> float {
          return score2(a, p.sl2, kb, p.causal && (jg > i + p.off), p.causal_fill2, jg >= p.Sk);
        };
        tmem_ld_32x32(t_s, r);
#pragma unroll 1
        for (int c = 0; c <= c_hi; ++c) {
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) cur[t] = __uint_as_float(r[t]);
          tmem_ld_32x32(t_s + (c == c_hi ? 0 : c + 1) * 32, r);  // last: first chunk of pass 2
#pragma unroll
          for (int t4 = 0; t4 < 8; ++t4) {
            float4 k4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (HAS_KB) k4 = lds128f(kb_s + 4 * (c * 32 + t4 * 4));
            const int jg = kv0 + c * 32 + t4 * 4;
            mt = fmaxf(mt, val(cur[t4 * 4 + 0], k4.x, jg + 0));
            mt = fmaxf(mt, val(cur[t4 * 4 + 1], k4.y, jg + 1));
            mt = fmaxf(mt, val(cur[t4 * 4 + 2], k4.z, jg + 2));
            mt = fmaxf(mt, val(cur[t4 * 4 + 3], k4.w, jg + 3));
          }
        }
        mt = fmaxf(mt, -FLT_MAX);  // (skipped future chunks are exactly -FLT_MAX)
        m_new = fmaxf(m, mt);
        alpha = ex2(m - m_new);
        if (j > 0) {
          mbar_wait(o_full, (j - 1) & 1);
          tc_fence_after();
        }
        lt = 0.f;
#pragma unroll 1
        for (int c = 0; c <= c_hi; ++c) {
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t) cur[t] = __uint_as_float(r[t]);
          if (c < c_hi) tmem_ld_32x32(t_s + (c + 1) * 32, r);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float pv[8];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int col = c * 32 + g * 8 + hh * 4;
              float4 k4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if constexpr (HAS_KB) k4 = lds128f(kb_s + 4 * col);
              const float kb[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float e = ex2(val(cur[g * 8 + hh * 4 + u], kb[u], kv0 + col + u) - m_new);
                lt += e;
                pv[hh * 4 + u] = e;
              }
            }
            uint32_t w0, w1, w2, w3;
            if constexpr (BF16) {
              w0 = pack_bf16x2(pv[0], pv[1]); w1 = pack_bf16x2(pv[2], pv[3]);
              w2 = pack_bf16x2(pv[4], pv[5]); w3 = pack_bf16x2(pv[6], pv[7]);
            } else {
              __half2 h0 = __floats2half2_rn(pv[0], pv[1]), h1 = __floats2half2_rn(pv[2], pv[3]);
              __half2 h2 = __floats2half2_rn(pv[4], pv[5]), h3 = __floats2half2_rn(pv[6], pv[7]);
              w0 = *reinterpret_cast<uint32_t*>(&h0); w1 = *reinterpret_cast<uint32_t*>(&h1);
              w2 = *reinterpret_cast<uint32_t*>(&h2); w3 = *reinterpret_cast<uint32_t*>(&h3);
            }
            const int chunk = (c & 1) * 4 + g;
            const uint32_t addr = p_row + (c >> 1) * FA_TILE + ((chunk ^ sw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                         : "memory");
          }
        }
        tc_fence_before();
        mbar_arrive(s_free);  // every TMEM load of this tile has completed
        if (c_hi < 3) {
          // chunks entirely in the future of this warp's rows: every score is exactly -FLT_MAX, so
          // p = 2^(-FLT_MAX - m) is 0 unless the row has seen nothing but masked keys so far (then 1)
          const float em = ex2(-FLT_MAX - m_new);
          uint32_t w;
          if constexpr (BF16) {
            w = pack_bf16x2(em, em);
          } else {
            __half2 hx = __floats2half2_rn(em, em);
            w = *reinterpret_cast<uint32_t*>(&hx);
          }
          for (int c = c_hi + 1; c < 4; ++c) {
            lt += 32.f * em;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int chunk = (c & 1) * 4 + g;
              const uint32_t addr = p_row + (c >> 1) * FA_TILE + ((chunk ^ sw) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(w) : "memory");
            }
          }
        }
      }
      CT_DBG_STAMP(2048 + 16 * j + 6);
      l = l * alpha + lt;
      m = m_new;
      fence_proxy_async_smem();
      // ---- O *= alpha (in TMEM), skipped when no row maximum of this warp moved ----
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(t_o, o0);
        tmem_ld_32x32(t_o + 32, o1);
        tmem_ld_wait();
        const float2 a2 = make_float2(alpha, alpha);
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          float2 x = __fmul2_rn(make_float2(__uint_as_float(o0[2 * t]), __uint_as_float(o0[2 * t + 1])), a2);
          float2 y = __fmul2_rn(make_float2(__uint_as_float(o1[2 * t]), __uint_as_float(o1[2 * t + 1])), a2);
          o0[2 * t] = __float_as_uint(x.x); o0[2 * t + 1] = __float_as_uint(x.y);
          o1[2 * t] = __float_as_uint(y.x); o1[2 * t + 1] = __float_as_uint(y.y);
        }
        tmem_st_32x32(t_o, o0);
        tmem_st_32x32(t_o + 32, o1);
        tmem_st_wait();
      }
      CT_DBG_STAMP(2048 + 16 * j + 7);
      tc_fence_before();
      mbar_arrive(p_ready);
      if constexpr (LAZY) mbar_arrive(p_half1);  // exact-maximum paths hand both panels over together
      CT_DBG_STAMP(2048 + 16 * j + 8);
    }
    // ---- epilogue: O / l -> merged-head layout, lse2 ----
    float o_acc[64];
    if (n_kv > 0) {
      mbar_wait(o_full, (n_kv - 1) & 1);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(t_o, o0);
      tmem_ld_32x32(t_o + 32, o1);
      tmem_ld_wait();
#pragma unroll
      for (int t = 0; t < 32; ++t) { o_acc[t] = __uint_as_float(o0[t]); o_acc[32 + t] = __uint_as_float(o1[t]); }
    } else {
#pragma unroll
      for (int t = 0; t < 64; ++t) o_acc[t] = 0.f;
    }
    if (i < p.Sq) {
      const float inv = (n_kv > 0) ? 1.f / l : 0.f;
      uint8_t* orow = reinterpret_cast<uint8_t*>(p.o) +
                      2 * ((int64_t)b * p.o_sb + (int64_t)h * p.o_sh + (int64_t)i * p.o_ss);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 w;
        if constexpr (BF16) {
          w.x = pack_bf16x2(o_acc[8 * g] * inv, o_acc[8 * g + 1] * inv);
          w.y = pack_bf16x2(o_acc[8 * g + 2] * inv, o_acc[8 * g + 3] * inv);
          w.z = pack_bf16x2(o_acc[8 * g + 4] * inv, o_acc[8 * g + 5] * inv);
          w.w = pack_bf16x2(o_acc[8 * g + 6] * inv, o_acc[8 * g + 7] * inv);
        } else {
          __half2 h0 = __floats2half2_rn(o_acc[8 * g] * inv, o_acc[8 * g + 1] * inv);
          __half2 h1 = __floats2half2_rn(o_acc[8 * g + 2] * inv, o_acc[8 * g + 3] * inv);
          __half2 h2 = __floats2half2_rn(o_acc[8 * g + 4] * inv, o_acc[8 * g + 5] * inv);
          __half2 h3 = __floats2half2_rn(o_acc[8 * g + 6] * inv, o_acc[8 * g + 7] * inv);
          w.x = *reinterpret_cast<uint32_t*>(&h0); w.y = *reinterpret_cast<uint32_t*>(&h1);
          w.z = *reinterpret_cast<uint32_t*>(&h2); w.w = *reinterpret_cast<uint32_t*>(&h3);
        }
        *reinterpret_cast<uint4*>(orow + 16 * g) = w;
      }
      if (p.lse2) p.lse2[((int64_t)b * p.H + h) * p.Sq + i] = (n_kv > 0) ? m + log2f(l) : -INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

// =================================================================================================
// tcgen05 forward, generation 4 (default). r01z (ncu, profiles/r01z_ncu_full_attention_summary.csv) showed the
// generation-2 kernel above at 3 % tensor-pipe / 37 % MUFU activity with 16 % of the warp slots occupied: one thread
// owns a whole 128-key score row, four softmax warps per CTA, so each scheduler holds two warps that take turns
// waiting for TMEM, for MUFU results and for each other. Here
//   * a query row is shared by TWO threads (warps w and w+4 of the same TMEM lane quarter own key columns [0,64) and
//     [64,128) of every tile): 8 softmax warps per CTA, 16 per SM, half the registers and half the dependent chain
//     per thread. The only thing the two halves must agree on per tile is the row maximum: it crosses through a spare
//     TMEM column (tcgen05.st / 64-thread named barrier / tcgen05.ld) — shared memory is full (2 CTAs x 113 KB);
//     the row sums stay private and are added once in the epilogue;
//   * the running maximum is a REFERENCE that only moves when a tile exceeds it by more than 2^8 (P stays inside
//     bf16/f16 range, sums are fp32), so O is rescaled in TMEM — a warp-collective load/store — only in the first
//     tiles of a row instead of whenever any of 32 rows moves;
//   * one score path for regular tiles and one for tiles that need a per-element test (causal diagonal, ragged key
//     edge) instead of five specialised ones: the instruction stream shrinks (r01z: 22 % of the forward samples were
//     `no_instructions` — instruction-cache misses in a fully unrolled 168-register kernel);
//   * the per-key bias of a tile is staged by an otherwise idle warp under an mbarrier pair, not by the softmax
//     threads behind a CTA-wide barrier;
//   * heavy (late) query tiles of ALL heads are scheduled first (longest-processing-time order over the whole grid).
// Warp roles (384 threads): 0-3 / 4-7 softmax halves (setmaxnreg 104), 8 TMA producer, 9 MMA issuer, 10 bias staging,
// 11 idle (setmaxnreg 32). TMEM (256 columns per CTA, 2 CTAs/SM): [0,128) S, [128,192) O, [192,196) row-maximum
// mailbox.
// =================================================================================================
constexpr int F4_THREADS = 384;
// Dynamic shared memory = the seven 16 KB tiles, nothing else: 2 x (114,688 + 1 KB static (barriers, bias staging; a
// 1024-aligned dynamic window costs that much in any case) + 1 KB reserved by the system) = 233,472 bytes = exactly
// what an SM has. One more byte and a second CTA no longer fits (r02e: the occupancy API reported 1 CTA/SM with the
// barriers and the bias row appended to the dynamic window, 57 us; see profiles/r02_attention_notes.md).
constexpr int F4_SMEM = 7 * FA_TILE;

template <uint32_t N>
__device__ __forceinline__ void f4_setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <uint32_t N>
__device__ __forceinline__ void f4_setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
// 64-thread barrier of the two warps that share TMEM lane quarter `wq`; the barrier ids are compile-time constants so
// that ptxas reserves 5 named barriers for the kernel, not all 16 (the barrier file of an SM is shared by its CTAs)
__device__ __forceinline__ void f4_pair_sync(int wq) {
  switch (wq) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}
__device__ __forceinline__ void tmem_ld_32x1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_32x1(uint32_t taddr, uint32_t r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}

// One 32-column chunk of a REGULAR tile: t = s * sl2 + kb clamped at -FLT_MAX (in place) when there is a per-key
// bias, raw scores otherwise (scaled inside the exp2); returns the running maximum.
template <bool HAS_KB, bool CLAMP>
__device__ __forceinline__ float f4_scale_max(uint32_t (&r)[32], uint32_t kb_addr, float sl2, float mt) {
  if constexpr (HAS_KB) {
    const float2 s2 = make_float2(sl2, sl2);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float4 k4 = lds128f(kb_addr + 16 * g);
      float2 a = __ffma2_rn(make_float2(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1])), s2,
                            make_float2(k4.x, k4.y));
      float2 c = __ffma2_rn(make_float2(__uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3])), s2,
                            make_float2(k4.z, k4.w));
      if constexpr (CLAMP) {  // only tiles with a masked key (bias -inf): it scores finfo.min, not -inf
        a.x = fmaxf(a.x, -FLT_MAX); a.y = fmaxf(a.y, -FLT_MAX);
        c.x = fmaxf(c.x, -FLT_MAX); c.y = fmaxf(c.y, -FLT_MAX);
      }
      r[4 * g] = __float_as_uint(a.x); r[4 * g + 1] = __float_as_uint(a.y);
      r[4 * g + 2] = __float_as_uint(c.x); r[4 * g + 3] = __float_as_uint(c.y);
      mt = fmaxf(mt, fmaxf(a.x, a.y));
      mt = fmaxf(mt, fmaxf(c.x, c.y));
    }
  } else {
#pragma unroll
    for (int g = 0; g < 16; ++g)
      mt = fmaxf(mt, fmaxf(__uint_as_float(r[2 * g]), __uint_as_float(r[2 * g + 1])));
  }
  return mt;
}

// The same for a tile that needs a per-element test: causally masked entries REPLACED by the fill (+ bias), clamp,
// keys beyond Sk excluded (-inf). `lim` = last visible column of this chunk for this thread's row (>= 32: all).
template <bool HAS_KB>
__device__ __forceinline__ float f4_mask_max(uint32_t (&r)[32], uint32_t kb_addr, float sl2, float cf2, int lim, int n_ok,
                                            float mt) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    float4 k4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (HAS_KB) k4 = lds128f(kb_addr + 16 * g);
    const float kb[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = 4 * g + u;
      float v = (t > lim) ? (cf2 + kb[u]) : fmaf(__uint_as_float(r[t]), sl2, kb[u]);
      v = fmaxf(v, -FLT_MAX);
      if (t >= n_ok) v = -INFINITY;
      r[t] = __float_as_uint(v);
      mt = fmaxf(mt, v);
    }
  }
  return mt;
}

// p = 2^(t - m) (SCALED) or 2^(s * sl2 - m) for one 32-column chunk -> four 16-byte pieces of the swizzled P panel
template <bool SCALED, bool BF16>
__device__ __forceinline__ void f4_exp_store(const uint32_t (&r)[32], int c, float m, float sl2, uint32_t p_row, int sw,
                                             float2& acc0, float2& acc1) {
  const float2 nm = make_float2(-m, -m);
  const float2 s2 = make_float2(sl2, sl2);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float2 a = make_float2(__uint_as_float(r[8 * g + 2 * u]), __uint_as_float(r[8 * g + 2 * u + 1]));
      if constexpr (SCALED) a = __fadd2_rn(a, nm);
      else a = __ffma2_rn(a, s2, nm);
      v[u] = make_float2(ex2(a.x), ex2(a.y));
    }
    acc0 = __fadd2_rn(acc0, v[0]); acc1 = __fadd2_rn(acc1, v[1]);
    acc0 = __fadd2_rn(acc0, v[2]); acc1 = __fadd2_rn(acc1, v[3]);
    uint32_t w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if constexpr (BF16) {
        w[u] = pack_bf16x2_alu(v[u].x, v[u].y);  // not F2FP: see ct_common.cuh
      } else {
        __half2 h = __floats2half2_rn(v[u].x, v[u].y);
        w[u] = *reinterpret_cast<uint32_t*>(&h);
      }
    }
    const uint32_t addr = p_row + (((4 * c + g) ^ sw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                 : "memory");
  }
}

template <bool HAS_KB, bool BF16>
__global__ void __launch_bounds__(F4_THREADS, 2)
    attn_fwd_tc4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const AttnP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sQ = base;
  const uint32_t sK = base + FA_TILE;
  const uint32_t sV = base + 3 * FA_TILE;
  const uint32_t sP = base + 5 * FA_TILE;  // two 64-key panels, one per softmax half
  __shared__ __align__(16) uint8_t f4_aux[656];  // mbarriers + TMEM slot (128 B), per-key bias of the tile (512 B), masked-key flag
  const uint32_t bars = smem_u32(f4_aux);
  const uint32_t q_full = bars, k_full = bars + 8, v_full = bars + 24, kv_empty = bars + 40, s_full = bars + 56,
                 s_free = bars + 64, p_ready = bars + 72, o_full = bars + 80, tmem_slot = bars + 88,
                 kb_full = bars + 96, kb_free = bars + 104, kb_s = bars + 128;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(f4_aux + 88);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_q_tiles = (p.Sq + 127) / 128;
  const int n_bh = p.B * p.H;
  const int q_tile = n_q_tiles - 1 - (int)(blockIdx.x / n_bh);  // all heads' heavy (late) query tiles first
  const int bh = blockIdx.x % n_bh;
  const int h = bh % p.H, b = bh / p.H;
  const int q0 = q_tile * 128;

  int n_kv = (p.Sk + 127) / 128;
  if (p.causal) {
    // rows that are fully masked (left padding) weight every key uniformly in the reference: visit them all
    const bool full_sweep = p.first_valid && (q0 + p.off < p.first_valid[b]);
    if (!full_sweep) {
      const int last_key = min(p.Sk - 1, q0 + 127 + p.off);
      n_kv = last_key < 0 ? 0 : last_key / 128 + 1;
    }
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(k_full + 8 * s, 1); mbar_init(v_full + 8 * s, 1); mbar_init(kv_empty + 8 * s, 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 256);
    mbar_init(p_ready, 256);
    mbar_init(o_full, 1);
    mbar_init(kb_full, 32);
    mbar_init(kb_free, 256);
    mbar_fence_init();
  }
  if (warp == 9) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp >= 8) {
    f4_setmaxnreg_dec<32>();
    if (warp == 8) {
      // ------------------------------ TMA producer ------------------------------
      if (lane == 0 && n_kv > 0) {
        mbar_expect_tx(q_full, FA_TILE);
        tma_load_4d(sQ, &tmQ, q_full, 0, q0, h, b);
        for (int j = 0; j < n_kv; ++j) {
          const int s = j & 1;
          mbar_wait(kv_empty + 8 * s, ((j >> 1) & 1) ^ 1);
          mbar_expect_tx(k_full + 8 * s, FA_TILE);
          tma_load_4d(sK + s * FA_TILE, &tmK, k_full + 8 * s, 0, j * 128, h, b);
          mbar_expect_tx(v_full + 8 * s, FA_TILE);
          tma_load_4d(sV + s * FA_TILE, &tmV, v_full + 8 * s, 0, j * 128, h, b);
        }
      }
    } else if (warp == 9) {
      // ------------------------------ MMA issuer ------------------------------
      if (lane == 0 && n_kv > 0) {
        const uint32_t idesc_s = umma_idesc_f16(BF16 ? 1 : 0, 0, 0, 128, 128);
        const uint32_t idesc_o = umma_idesc_f16(BF16 ? 1 : 0, 0, 1, 128, 64);
        mbar_wait(q_full, 0);
        for (int j = 0; j <= n_kv; ++j) {
          if (j < n_kv) {  // S(j) = Q K(j)^T; for j > 0 it runs under the softmax of tile j-1
            const int s = j & 1;
            mbar_wait(k_full + 8 * s, (j >> 1) & 1);
            if (j > 0) mbar_wait(s_free, (j - 1) & 1);  // every softmax thread holds S(j-1) in registers
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(tmem, umma_smem_desc_sw128(sQ + k * 32, 0, 1024),
                       umma_smem_desc_sw128(sK + s * FA_TILE + k * 32, 0, 1024), idesc_s, k > 0);
            umma_commit(s_full);
          }
          if (j > 0) {  // O += P(j-1) V(j-1)
            const int jj = j - 1, s = jj & 1;
            mbar_wait(p_ready, jj & 1);
            mbar_wait(v_full + 8 * s, (jj >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_f16(tmem + 128, umma_smem_desc_sw128(sP + (k >> 2) * FA_TILE + (k & 3) * 32, 0, 1024),
                       umma_smem_desc_sw128(sV + s * FA_TILE + k * 2048, 64 * 128, 1024), idesc_o, (jj > 0 || k > 0));
            umma_commit(kv_empty + 8 * s);
            umma_commit(o_full);
          }
        }
      }
    } else if (warp == 10) {
      // ------------------------------ per-key bias staging ------------------------------
      if constexpr (HAS_KB) {
        const float* kb_row = p.kbias2 + (int64_t)b * p.kb_sb + (int64_t)h * p.kb_sh;
        for (int j = 0; j < n_kv; ++j) {
          float v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int col = j * 128 + lane * 4 + u;
            v[u] = col < p.Sk ? __ldg(kb_row + col) : 0.f;
          }
          const bool any_masked = __any_sync(0xffffffffu, v[0] < -1e30f || v[1] < -1e30f || v[2] < -1e30f || v[3] < -1e30f);
          if (j > 0) mbar_wait(kb_free, (j - 1) & 1);  // every softmax thread is done with the previous tile's bias
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(kb_s + 16 * lane), "f"(v[0]), "f"(v[1]),
                       "f"(v[2]), "f"(v[3])
                       : "memory");
          if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(kb_s + 512), "r"(any_masked ? 1u : 0u) : "memory");
          mbar_arrive(kb_full);
        }
      }
    }
  } else {
    f4_setmaxnreg_inc<104>();
    // ------------------------------ softmax: two threads per query row ------------------------------
    const int half = warp >> 2;  // key columns [64 * half, 64 * half + 64) of every tile
    const int wq = warp & 3;
    const int qr = wq * 32 + lane;  // row inside the tile == TMEM lane
    const int i = q0 + qr;
    const uint32_t t_lane = tmem + ((uint32_t)(wq * 32) << 16);
    const uint32_t t_s = t_lane + 64 * half;
    const uint32_t t_o = t_lane + 128 + 32 * half;  // this thread's 32 of the 64 output columns
    const uint32_t t_mail = t_lane + 192;
    const uint32_t p_row = sP + half * FA_TILE + qr * 128;
    const uint32_t kb_half = kb_s + 256 * half;
    const int sw = qr & 7;
    const int vis = i + p.off;  // last key this row may see under the causal mask
    float m_ref = -INFINITY, l = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const int col0 = j * 128 + 64 * half;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t ra[32], rb[32];
      tmem_ld_32x32(t_s, ra);
      tmem_ld_32x32(t_s + 32, rb);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);  // S(j) is in registers: S(j+1) may be issued
      // CTA-uniform: does any element of this tile need a test (causal boundary inside the tile, ragged key edge)?
      const bool masked_tile = (p.causal && (j * 128 + 127 > q0 + p.off)) || (j * 128 + 128 > p.Sk);
      if constexpr (HAS_KB) mbar_wait(kb_full, j & 1);
      float mt = -INFINITY;
      bool scaled = HAS_KB;
      bool key_masked = false;  // CTA-uniform: does this tile hold a masked key (staged by the bias warp)?
      if constexpr (HAS_KB) key_masked = lds32(kb_s + 512) != 0u;
      if (!masked_tile && !key_masked) {
        mt = f4_scale_max<HAS_KB, false>(ra, kb_half, p.sl2, mt);
        mt = f4_scale_max<HAS_KB, false>(rb, kb_half + 128, p.sl2, mt);
        if constexpr (!HAS_KB) mt *= p.sl2;  // sl2 > 0 (checked on the host)
      } else if (!masked_tile) {
        mt = f4_scale_max<HAS_KB, true>(ra, kb_half, p.sl2, mt);
        mt = f4_scale_max<HAS_KB, true>(rb, kb_half + 128, p.sl2, mt);
      } else {
        const int lim = p.causal ? vis - col0 : 1 << 20;
        const int n_ok = p.Sk - col0;
        mt = f4_mask_max<HAS_KB>(ra, kb_half, p.sl2, p.causal_fill2, lim, n_ok, mt);
        mt = f4_mask_max<HAS_KB>(rb, kb_half + 128, p.sl2, p.causal_fill2, lim - 32, n_ok - 32, mt);
        scaled = true;
      }
      if constexpr (HAS_KB) mbar_arrive(kb_free);
      // ---- row maximum of the tile: through the TMEM mailbox to the thread that owns the other 64 columns ----
      tmem_st_32x1(t_mail + 2 * (j & 1) + half, __float_as_uint(mt));
      tmem_st_wait();
      tc_fence_before();
      f4_pair_sync(wq);
      tc_fence_after();
      uint32_t other;
      tmem_ld_32x1(t_mail + 2 * (j & 1) + (half ^ 1), other);
      tmem_ld_wait();
      mt = fmaxf(fmaxf(mt, __uint_as_float(other)), -FLT_MAX);
      // ---- reference maximum: moves only when this tile exceeds it by more than 2^8 ----
      float alpha = 1.f;
      const bool need = mt > m_ref + 8.f;  // (first tile: m_ref = -inf)
      if (need) {
        alpha = ex2(m_ref - mt);
        l *= alpha;
        m_ref = mt;
      }
      const bool rescale = j > 0 && __any_sync(0xffffffffu, need);
      if (j > 0) {  // P(j-1) V(j-1) retired: the P panel may be overwritten, O may be rescaled
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
      }
      float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0;
      if (scaled) {
        f4_exp_store<true, BF16>(ra, 0, m_ref, p.sl2, p_row, sw, acc0, acc1);
        f4_exp_store<true, BF16>(rb, 1, m_ref, p.sl2, p_row, sw, acc0, acc1);
      } else {
        f4_exp_store<false, BF16>(ra, 0, m_ref, p.sl2, p_row, sw, acc0, acc1);
        f4_exp_store<false, BF16>(rb, 1, m_ref, p.sl2, p_row, sw, acc0, acc1);
      }
      l += (acc0.x + acc0.y) + (acc1.x + acc1.y);
      fence_proxy_async_smem();
      if (rescale) {  // warp-uniform: the TMEM accesses are warp-collective
        uint32_t o0[32];
        tmem_ld_32x32(t_o, o0);
        tmem_ld_wait();
        const float2 a2 = make_float2(alpha, alpha);
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const float2 x = __fmul2_rn(make_float2(__uint_as_float(o0[2 * t]), __uint_as_float(o0[2 * t + 1])), a2);
          o0[2 * t] = __float_as_uint(x.x); o0[2 * t + 1] = __float_as_uint(x.y);
        }
        tmem_st_32x32(t_o, o0);
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    // ---- epilogue: l = l(half 0) + l(half 1); O / l -> merged-head layout (32 of the 64 columns each), lse2 ----
    if (n_kv > 0) {
      tmem_st_32x1(t_mail + 2 * (n_kv & 1) + half, __float_as_uint(l));
      tmem_st_wait();
      tc_fence_before();
      f4_pair_sync(wq);
      tc_fence_after();
      uint32_t other;
      tmem_ld_32x1(t_mail + 2 * (n_kv & 1) + (half ^ 1), other);
      tmem_ld_wait();
      l += __uint_as_float(other);
      mbar_wait(o_full, (n_kv - 1) & 1);
      tc_fence_after();
    }
    uint32_t o0[32];
    if (n_kv > 0) {
      tmem_ld_32x32(t_o, o0);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int t = 0; t < 32; ++t) o0[t] = 0u;
    }
    if (i < p.Sq) {
      const float inv = (n_kv > 0) ? 1.f / l : 0.f;
      uint8_t* orow = reinterpret_cast<uint8_t*>(p.o) +
                      2 * ((int64_t)b * p.o_sb + (int64_t)h * p.o_sh + (int64_t)i * p.o_ss + 32 * half);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float f[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(o0[8 * g + u]) * inv;
        uint4 w;
        if constexpr (BF16) {
          w.x = pack_bf16x2(f[0], f[1]); w.y = pack_bf16x2(f[2], f[3]);
          w.z = pack_bf16x2(f[4], f[5]); w.w = pack_bf16x2(f[6], f[7]);
        } else {
          __half2 h0 = __floats2half2_rn(f[0], f[1]), h1 = __floats2half2_rn(f[2], f[3]);
          __half2 h2 = __floats2half2_rn(f[4], f[5]), h3 = __floats2half2_rn(f[6], f[7]);
          w.x = *reinterpret_cast<uint32_t*>(&h0); w.y = *reinterpret_cast<uint32_t*>(&h1);
          w.z = *reinterpret_cast<uint32_t*>(&h2); w.w = *reinterpret_cast<uint32_t*>(&h3);
        }
        *reinterpret_cast<uint4*>(orow + 16 * g) = w;
      }
      if (half == 0 && p.lse2) p.lse2[((int64_t)b * p.H + h) * p.Sq + i] = (n_kv > 0) ? m_ref + log2f(l) : -INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

// =================================================================================================
// tcgen05 backward
// =================================================================================================
struct AttnBwdP {
  AttnP f;
  const float* delta;
  float* dq_accum;  // [B, Sq, H, 64] f32
  void* dk; int64_t dk_sb, dk_sh, dk_ss;
  void* dv; int64_t dv_sb, dv_sh, dv_ss;
};
constexpr int FB_THREADS = 320;  // TMA warp + MMA warp + 8 compute warps (2 per TMEM lane quarter)
constexpr int FB_SMEM = 2 * FA_TILE /*K,V*/ + 4 * FA_TILE /*2 x (Q,dO)*/ + 2 * FA_TILE /*P^T*/ +
                        2 * FA_TILE /*dS^T*/ + 128 /*barriers*/ + 1024 /*lse2, delta of the query tile*/;
// v3 (pipelined): second P^T / dS^T buffer pair and a second statistics buffer
constexpr int FB_SMEM_PIPE = FB_SMEM + 4 * FA_TILE + 1024;
// dQ workspace, tiled layout: [b][h][query tile][d/4 (16)][row (128)][4] f32 — a warp's red.global.add.v4 then covers
// 512 contiguous bytes (32 rows x 16 B) instead of 32 different 128-byte lines
constexpr int FB_DQ_TILE = 128 * 64;

__device__ __forceinline__ void st_row64(void* basep, int64_t elem_off, const float (&v)[64], int fmt) {
  uint8_t* row = reinterpret_cast<uint8_t*>(basep) + 2 * elem_off;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    uint4 w;
    if (fmt == 1) {
      w.x = pack_bf16x2(v[8 * g], v[8 * g + 1]); w.y = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
      w.z = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]); w.w = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
    } else {
      __half2 h0 = __floats2half2_rn(v[8 * g], v[8 * g + 1]), h1 = __floats2half2_rn(v[8 * g + 2], v[8 * g + 3]);
      __half2 h2 = __floats2half2_rn(v[8 * g + 4], v[8 * g + 5]), h3 = __floats2half2_rn(v[8 * g + 6], v[8 * g + 7]);
      w.x = *reinterpret_cast<uint32_t*>(&h0); w.y = *reinterpret_cast<uint32_t*>(&h1);
      w.z = *reinterpret_cast<uint32_t*>(&h2); w.w = *reinterpret_cast<uint32_t*>(&h3);
    }
    *reinterpret_cast<uint4*>(row + 16 * g) = w;
  }
}

__global__ void __launch_bounds__(FB_THREADS, 1)
    attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                       const AttnBwdP bp) {
  const AttnP& p = bp.f;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sK = base, sV = base + FA_TILE;
  const uint32_t sQ = base + 2 * FA_TILE;   // 2 stages, each Q then dO
  const uint32_t sPT = base + 6 * FA_TILE;  // 2 panels
  const uint32_t sDS = base + 8 * FA_TILE;  // 2 panels
  const uint32_t bars = base + 10 * FA_TILE;
  const uint32_t kv_full = bars, qdo_full = bars + 8, qdo_empty = bars + 24, sdp_full = bars + 40,
                 pds_ready = bars + 48, dq_full = bars + 56, dkv_full = bars + 64, tmem_slot = bars + 72,
                 lse_s = bars + 128, del_s = bars + 640;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem + 10 * FA_TILE + 72);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv_tiles = (p.Sk + 127) / 128;
  const int kv_tile = blockIdx.x % n_kv_tiles;
  const int bh = blockIdx.x / n_kv_tiles;
  const int h = bh % p.H, b = bh / p.H;
  const int kv0 = kv_tile * 128;
  const int n_q_tiles = (p.Sq + 127) / 128;
  int i_start = 0;
  if (p.causal) {
    const bool full_sweep = p.first_valid && (p.first_valid[b] > p.off);
    if (!full_sweep) i_start = max(0, (kv0 - p.off) / 128);
  }
  const int n_it = max(0, n_q_tiles - i_start);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(qdo_full + 8 * s, 1); mbar_init(qdo_empty + 8 * s, 1); }
    mbar_init(sdp_full, 1);
    mbar_init(pds_ready, 256);
    mbar_init(dq_full, 1);
    mbar_init(dkv_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t T_ST = tmem, T_DPT = tmem + 128, T_DV = tmem + 256, T_DK = tmem + 320, T_DQ = tmem + 384;

  if (warp == 0) {
    if (lane == 0 && n_it > 0) {
      mbar_expect_tx(kv_full, 2 * FA_TILE);
      tma_load_4d(sK, &tmK, kv_full, 0, kv0, h, b);
      tma_load_4d(sV, &tmV, kv_full, 0, kv0, h, b);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1, q0 = (i_start + it) * 128;
        mbar_wait(qdo_empty + 8 * s, ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(qdo_full + 8 * s, 2 * FA_TILE);
        tma_load_4d(sQ + s * 2 * FA_TILE, &tmQ, qdo_full + 8 * s, 0, q0, h, b);
        tma_load_4d(sQ + s * 2 * FA_TILE + FA_TILE, &tmDO, qdo_full + 8 * s, 0, q0, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_it > 0) {
      const uint32_t idesc_kk = umma_idesc_f16(p.fmt, 0, 0, 128, 128);  // S^T, dP^T
      const uint32_t idesc_km = umma_idesc_f16(p.fmt, 0, 1, 128, 64);   // dV, dK
      const uint32_t idesc_mm = umma_idesc_f16(p.fmt, 1, 1, 128, 64);   // dQ
      auto issue_sdp = [&](int it) {
        const int s = it & 1;
        const uint32_t q = sQ + s * 2 * FA_TILE, d_o = q + FA_TILE;
        mbar_wait(qdo_full + 8 * s, (it >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)  // S^T[kv, q] = K[kv, d] . Q[q, d]
          umma_f16(T_ST, umma_smem_desc_sw128(sK + k * 32, 0, 1024), umma_smem_desc_sw128(q + k * 32, 0, 1024),
                   idesc_kk, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dP^T[kv, q] = V[kv, d] . dO[q, d]
          umma_f16(T_DPT, umma_smem_desc_sw128(sV + k * 32, 0, 1024),
                   umma_smem_desc_sw128(d_o + k * 32, 0, 1024), idesc_kk, k > 0);
        umma_commit(sdp_full);
      };
      mbar_wait(kv_full, 0);
      issue_sdp(0);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1;
        const uint32_t q = sQ + s * 2 * FA_TILE, d_o = q + FA_TILE;
        mbar_wait(pds_ready, it & 1);
        tc_fence_after();
        // dQ first: the compute warps turn it into red.global.add traffic (slow) while dV / dK and
        // the next tile's S^T / dP^T run on the tensor pipe
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dQ[q, d] = dS[q, kv] . K[kv, d]  (A MN-major view of dS^T)
          umma_f16(T_DQ, umma_smem_desc_sw128(sDS + k * 2048, FA_TILE, 1024),
                   umma_smem_desc_sw128(sK + k * 2048, 64 * 128, 1024), idesc_mm, k > 0);
        umma_commit(dq_full);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dV[kv, d] += P^T[kv, q] . dO[q, d]   (B MN-major: rows = q)
          umma_f16(T_DV, umma_smem_desc_sw128(sPT + (k >> 2) * FA_TILE + (k & 3) * 32, 0, 1024),
                   umma_smem_desc_sw128(d_o + k * 2048, 64 * 128, 1024), idesc_km, (it > 0 || k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dK[kv, d] += dS^T[kv, q] . Q[q, d]
          umma_f16(T_DK, umma_smem_desc_sw128(sDS + (k >> 2) * FA_TILE + (k & 3) * 32, 0, 1024),
                   umma_smem_desc_sw128(q + k * 2048, 64 * 128, 1024), idesc_km, (it > 0 || k > 0));
        umma_commit(qdo_empty + 8 * s);
        if (it + 1 < n_it) issue_sdp(it + 1);
      }
      umma_commit(dkv_full);
    }
  } else {
    // 8 compute warps: a single warp per scheduler cannot hide its own ALU/MUFU dependency latency
    // (measured: ~20% issue utilisation with 4 warps), so every TMEM lane quarter is served by two
    // warps that split the 128 query columns (the backward math is purely elementwise per (key,query)).
    const int rr = (warp & 3) * 32 + lane;  // key row inside the tile (S^T) / query row (dQ)
    const int hf = (warp - 2) >> 2;         // which half of the columns this warp owns
    const int jg = kv0 + rr;
    const uint32_t t_lane = (uint32_t)((warp & 3) * 32) << 16;
    const float kb = (p.kbias2 && jg < p.Sk)
                         ? __ldg(p.kbias2 + (int64_t)b * p.kb_sb + (int64_t)h * p.kb_sh + jg) : 0.f;
    const bool key_oob = jg >= p.Sk;
    const float* lse_bh = p.lse2 + ((int64_t)b * p.H + h) * p.Sq;
    const float* del_bh = bp.delta + ((int64_t)b * p.H + h) * p.Sq;
    const int sw = rr & 7;
    // per-query statistics of the current query tile, staged through smem (prefetched one tile ahead)
    float lse_next = INFINITY, del_next = 0.f;
    if (hf == 0 && n_it > 0 && i_start * 128 + rr < p.Sq) {
      lse_next = __ldg(lse_bh + i_start * 128 + rr);
      del_next = __ldg(del_bh + i_start * 128 + rr);
    }
    for (int it = 0; it < n_it; ++it) {
      const int q0 = (i_start + it) * 128;
      CT_DBG_STAMP(16 * it + 0);
      mbar_wait(sdp_full, it & 1);
      CT_DBG_STAMP(16 * it + 1);
      tc_fence_after();
      // all 256 threads are past dq_full(it-1): nobody still reads the previous tile's statistics
      if (hf == 0) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(lse_s + 4 * rr), "f"(lse_next) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(del_s + 4 * rr), "f"(del_next) : "memory");
      }
      bar_sync_named(1, 256);
      if (hf == 0) {
        const int nq = q0 + 128 + rr;
        const bool ok = (it + 1 < n_it) && nq < p.Sq;
        lse_next = ok ? __ldg(lse_bh + nq) : INFINITY;
        del_next = ok ? __ldg(del_bh + nq) : 0.f;
      }
      uint32_t rs[32], rd[32];
      tmem_ld_32x32(T_ST + t_lane + hf * 64, rs);
      tmem_ld_32x32(T_DPT + t_lane + hf * 64, rd);
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = hf * 2 + cc;  // 32-column chunk index inside the 128-query tile
        float cs[32], cdp[32];
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t) { cs[t] = __uint_as_float(rs[t]); cdp[t] = __uint_as_float(rd[t]); }
        if (cc == 0) {
          tmem_ld_32x32(T_ST + t_lane + (c + 1) * 32, rs);
          tmem_ld_32x32(T_DPT + t_lane + (c + 1) * 32, rd);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // 16-byte output chunks of 8 query columns
          float pt[8], ds[8];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int col = c * 32 + g * 8 + hh * 4;
            const int qg = q0 + col;
            float4 a, d;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(lse_s + 4 * col));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "r"(del_s + 4 * col));
            const float ls[4] = {a.x, a.y, a.z, a.w};
            const float dl[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int e = g * 8 + hh * 4 + u;
              const bool fut = p.causal && (jg > qg + u + p.off);
              const float v = score2(cs[e], p.sl2, kb, fut, p.causal_fill2, false);
              float pe = ex2(v - ls[u]);
              if (key_oob) pe = 0.f;
              pt[hh * 4 + u] = pe;
              // a causally masked score is a constant in the reference (modeling_gpt.py:89 `w*b`,
              // modeling_bloom.py:108 masked_fill): P still feeds dV, but no gradient reaches q.k
              ds[hh * 4 + u] = fut ? 0.f : pe * (cdp[e] - dl[u]) * p.scale;
            }
          }
          const int chunk = (c & 1) * 4 + g;
          const uint32_t off = rr * 128 + (c >> 1) * FA_TILE + ((chunk ^ sw) << 4);
          uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
          if (p.fmt == 1) {
            a0 = pack_bf16x2(pt[0], pt[1]); a1 = pack_bf16x2(pt[2], pt[3]);
            a2 = pack_bf16x2(pt[4], pt[5]); a3 = pack_bf16x2(pt[6], pt[7]);
            b0 = pack_bf16x2(ds[0], ds[1]); b1 = pack_bf16x2(ds[2], ds[3]);
            b2 = pack_bf16x2(ds[4], ds[5]); b3 = pack_bf16x2(ds[6], ds[7]);
          } else {
            __half2 x;
            x = __floats2half2_rn(pt[0], pt[1]); a0 = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(pt[2], pt[3]); a1 = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(pt[4], pt[5]); a2 = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(pt[6], pt[7]); a3 = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(ds[0], ds[1]); b0 = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(ds[2], ds[3]); b1 = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(ds[4], ds[5]); b2 = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(ds[6], ds[7]); b3 = *reinterpret_cast<uint32_t*>(&x);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sPT + off), "r"(a0), "r"(a1),
                       "r"(a2), "r"(a3) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sDS + off), "r"(b0), "r"(b1),
                       "r"(b2), "r"(b3) : "memory");
        }
      }
      CT_DBG_STAMP(16 * it + 2);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_ready);
      // ---- dQ tile: this thread owns query row (q0 + rr), columns [32*hf, 32*hf + 32) of d ----
      CT_DBG_STAMP(16 * it + 3);
      mbar_wait(dq_full, it & 1);
      CT_DBG_STAMP(16 * it + 4);
      tc_fence_after();
      const int qi = q0 + rr;
      {
        uint32_t r[32];
        tmem_ld_32x32(T_DQ + t_lane + hf * 32, r);
        tmem_ld_wait();
        if (qi < p.Sq) {
          float* dst = bp.dq_accum + (((int64_t)b * p.Sq + qi) * p.H + h) * 64 + hf * 32;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * g),
                         "f"(__uint_as_float(r[4 * g])), "f"(__uint_as_float(r[4 * g + 1])),
                         "f"(__uint_as_float(r[4 * g + 2])), "f"(__uint_as_float(r[4 * g + 3]))
                         : "memory");
        }
      }
      CT_DBG_STAMP(16 * it + 5);
      tc_fence_before();
    }
    // ---- dK / dV for this key row: 32 of the 64 head-dim columns per thread ----
    if (n_it > 0) {
      mbar_wait(dkv_full, 0);
      tc_fence_after();
    }
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      uint32_t r[32];
      if (n_it > 0) {
        tmem_ld_32x32((which == 0 ? T_DV : T_DK) + t_lane + hf * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int t = 0; t < 32; ++t) r[t] = 0u;
      }
      if (!key_oob) {
        void* basep = which == 0 ? bp.dv : bp.dk;
        const int64_t eo = which == 0
            ? (int64_t)b * bp.dv_sb + (int64_t)h * bp.dv_sh + (int64_t)jg * bp.dv_ss
            : (int64_t)b * bp.dk_sb + (int64_t)h * bp.dk_sh + (int64_t)jg * bp.dk_ss;
        uint8_t* row = reinterpret_cast<uint8_t*>(basep) + 2 * (eo + hf * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          const float f0 = __uint_as_float(r[8 * g]), f1 = __uint_as_float(r[8 * g + 1]),
                      f2 = __uint_as_float(r[8 * g + 2]), f3 = __uint_as_float(r[8 * g + 3]),
                      f4 = __uint_as_float(r[8 * g + 4]), f5 = __uint_as_float(r[8 * g + 5]),
                      f6 = __uint_as_float(r[8 * g + 6]), f7 = __uint_as_float(r[8 * g + 7]);
          if (p.fmt == 1) {
            w.x = pack_bf16x2(f0, f1); w.y = pack_bf16x2(f2, f3); w.z = pack_bf16x2(f4, f5); w.w = pack_bf16x2(f6, f7);
          } else {
            __half2 x;
            x = __floats2half2_rn(f0, f1); w.x = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(f2, f3); w.y = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(f4, f5); w.z = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(f6, f7); w.w = *reinterpret_cast<uint32_t*>(&x);
          }
          *reinterpret_cast<uint4*>(row + 16 * g) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// =================================================================================================
// tcgen05 backward, v2
// =================================================================================================
// ncu on v1 (same pipeline, 8 compute warps): 23 warp instructions per score element, and every query
// tile serialises  [S^T, dP^T MMAs] -> [element math] -> [dQ, dV, dK MMAs] -> [dQ red.add]. v2:
//   * each compute warp loads its two 32-query chunks of S^T and dP^T into registers up front and
//     immediately releases the two TMEM buffers (sdp_free): the MMA warp issues S^T / dP^T of the NEXT
//     query tile before the dQ / dV / dK MMAs of the current one, so they run under the element math;
//   * dQ of tile it-1 is drained (TMEM -> red.global.add) at the start of tile it, while the S^T / dP^T
//     loads are in flight, instead of stalling on the dQ MMA right after issuing its operands;
//   * element math on register pairs: statistics staged as -lse2 and -delta*scale, so
//     P^T = 2^(s*sl2 + (kb - lse2)), dS^T = P^T * (dP*scale - delta*scale): 2 fma.f32x2 + 1 mul.f32x2
//     + 1 add.f32x2 per element pair; 32-query chunks are classified per warp (visible / entirely in the
//     future of the warp's 32 keys on an aligned causal diagonal: no TMEM read, dS = 0 / generic = v1
//     arithmetic) with the classification hoisted out of the element loops.
struct FbCtx {
  float kb, sl2, scale, cf2;
  int jg, off, causal, key_oob;
  uint32_t lse_s, del_s, sPT, sDS;
  int rr, sw;
};

template <bool BF16>
__device__ __forceinline__ void fb2_store_pair(const FbCtx& cx, int c, int g, const float (&pt)[8], const float (&ds)[8]) {
  uint32_t a[4], d[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if constexpr (BF16) {
      a[u] = pack_bf16x2(pt[2 * u], pt[2 * u + 1]);
      d[u] = pack_bf16x2(ds[2 * u], ds[2 * u + 1]);
    } else {
      __half2 x = __floats2half2_rn(pt[2 * u], pt[2 * u + 1]);
      a[u] = *reinterpret_cast<uint32_t*>(&x);
      x = __floats2half2_rn(ds[2 * u], ds[2 * u + 1]);
      d[u] = *reinterpret_cast<uint32_t*>(&x);
    }
  }
  const int ch = (c & 1) * 4 + g;
  const uint32_t off = cx.rr * 128 + (c >> 1) * FA_TILE + ((ch ^ cx.sw) << 4);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cx.sPT + off), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3])
               : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cx.sDS + off), "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3])
               : "memory");
}

// one 32-query chunk c of this thread's key row. KIND 0 = visible, 1 = entirely future, 2 = generic
template <int KIND, bool BF16>
__device__ __forceinline__ void fb2_chunk(const FbCtx& cx, const uint32_t (&rs)[32], const uint32_t (&rd)[32], int c,
                                          int q0) {
  const float2 sl2v = make_float2(cx.sl2, cx.sl2), scv = make_float2(cx.scale, cx.scale), kbv = make_float2(cx.kb, cx.kb);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float pt[8], ds[8];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int col = c * 32 + g * 8 + hh * 4;
      const float4 nl = lds128f(cx.lse_s + 4 * col);
      const float nls[4] = {nl.x, nl.y, nl.z, nl.w};
      if constexpr (KIND == 0) {
        const float4 nd = lds128f(cx.del_s + 4 * col);
        const float nds[4] = {nd.x, nd.y, nd.z, nd.w};
#pragma unroll
        for (int u2 = 0; u2 < 2; ++u2) {
          const int e = g * 8 + hh * 4 + 2 * u2;
          const float2 add = __fadd2_rn(kbv, make_float2(nls[2 * u2], nls[2 * u2 + 1]));
          const float2 t = __ffma2_rn(make_float2(__uint_as_float(rs[e]), __uint_as_float(rs[e + 1])), sl2v, add);
          const float2 pe = make_float2(ex2(t.x), ex2(t.y));
          const float2 w = __ffma2_rn(make_float2(__uint_as_float(rd[e]), __uint_as_float(rd[e + 1])), scv,
                                      make_float2(nds[2 * u2], nds[2 * u2 + 1]));
          const float2 d2 = __fmul2_rn(pe, w);
          pt[hh * 4 + 2 * u2] = pe.x; pt[hh * 4 + 2 * u2 + 1] = pe.y;
          ds[hh * 4 + 2 * u2] = d2.x; ds[hh * 4 + 2 * u2 + 1] = d2.y;
        }
      } else if constexpr (KIND == 1) {
        // causally masked for every key of this warp: the score is the (clamped) fill, a constant
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          pt[hh * 4 + u] = ex2(-FLT_MAX + nls[u]);
          ds[hh * 4 + u] = 0.f;
        }
      } else {
        const float4 nd = lds128f(cx.del_s + 4 * col);
        const float nds[4] = {nd.x, nd.y, nd.z, nd.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = g * 8 + hh * 4 + u;
          const int qg = q0 + col + u;
          const bool fut = cx.causal && (cx.jg > qg + cx.off);
          const float v = score2(__uint_as_float(rs[e]), cx.sl2, cx.kb, fut, cx.cf2, false);
          float pe = ex2(v + nls[u]);
          if (cx.key_oob) pe = 0.f;
          pt[hh * 4 + u] = pe;
          // a causally masked score is a constant in the reference (modeling_gpt.py:89 `w*b`,
          // modeling_bloom.py:108 masked_fill): P still feeds dV, but no gradient reaches q.k
          ds[hh * 4 + u] = fut ? 0.f : pe * fmaf(__uint_as_float(rd[e]), cx.scale, nds[u]);
        }
      }
    }
    fb2_store_pair<BF16>(cx, c, g, pt, ds);
  }
}

template <bool BF16>
__device__ __forceinline__ void fb7_store_pair(const FbCtx& cx, int c, int g, const float (&pt)[8], const float (&ds)[8]) {
  uint32_t a[4], d[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if constexpr (BF16) {
      a[u] = pack_bf16x2_alu(pt[2 * u], pt[2 * u + 1]);
      d[u] = pack_bf16x2_alu(ds[2 * u], ds[2 * u + 1]);
    } else {
      __half2 x = __floats2half2_rn(pt[2 * u], pt[2 * u + 1]);
      a[u] = *reinterpret_cast<uint32_t*>(&x);
      x = __floats2half2_rn(ds[2 * u], ds[2 * u + 1]);
      d[u] = *reinterpret_cast<uint32_t*>(&x);
    }
  }
  const int ch = (c & 1) * 4 + g;
  const uint32_t off = cx.rr * 128 + (c >> 1) * FA_TILE + ((ch ^ cx.sw) << 4);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cx.sPT + off), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3])
               : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cx.sDS + off), "r"(d[0]), "r"(d[1]), "r"(d[2]), "r"(d[3])
               : "memory");
}

// (v7) the same chunk with the bf16 packing on the ALU pipe instead of F2FP (XU). One 32-query chunk c of this thread's key row. KIND 0 = visible, 1 = entirely future, 2 = generic
template <int KIND, bool BF16>
__device__ __forceinline__ void fb7_chunk(const FbCtx& cx, const uint32_t (&rs)[32], const uint32_t (&rd)[32], int c,
                                          int q0) {
  const float2 sl2v = make_float2(cx.sl2, cx.sl2), scv = make_float2(cx.scale, cx.scale), kbv = make_float2(cx.kb, cx.kb);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float pt[8], ds[8];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int col = c * 32 + g * 8 + hh * 4;
      const float4 nl = lds128f(cx.lse_s + 4 * col);
      const float nls[4] = {nl.x, nl.y, nl.z, nl.w};
      if constexpr (KIND == 0) {
        const float4 nd = lds128f(cx.del_s + 4 * col);
        const float nds[4] = {nd.x, nd.y, nd.z, nd.w};
#pragma unroll
        for (int u2 = 0; u2 < 2; ++u2) {
          const int e = g * 8 + hh * 4 + 2 * u2;
          const float2 add = __fadd2_rn(kbv, make_float2(nls[2 * u2], nls[2 * u2 + 1]));
          const float2 t = __ffma2_rn(make_float2(__uint_as_float(rs[e]), __uint_as_float(rs[e + 1])), sl2v, add);
          const float2 pe = make_float2(ex2(t.x), ex2(t.y));
          const float2 w = __ffma2_rn(make_float2(__uint_as_float(rd[e]), __uint_as_float(rd[e + 1])), scv,
                                      make_float2(nds[2 * u2], nds[2 * u2 + 1]));
          const float2 d2 = __fmul2_rn(pe, w);
          pt[hh * 4 + 2 * u2] = pe.x; pt[hh * 4 + 2 * u2 + 1] = pe.y;
          ds[hh * 4 + 2 * u2] = d2.x; ds[hh * 4 + 2 * u2 + 1] = d2.y;
        }
      } else if constexpr (KIND == 1) {
        // causally masked for every key of this warp: the score is the (clamped) fill, a constant
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          pt[hh * 4 + u] = ex2(-FLT_MAX + nls[u]);
          ds[hh * 4 + u] = 0.f;
        }
      } else {
        const float4 nd = lds128f(cx.del_s + 4 * col);
        const float nds[4] = {nd.x, nd.y, nd.z, nd.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = g * 8 + hh * 4 + u;
          const int qg = q0 + col + u;
          const bool fut = cx.causal && (cx.jg > qg + cx.off);
          const float v = score2(__uint_as_float(rs[e]), cx.sl2, cx.kb, fut, cx.cf2, false);
          float pe = ex2(v + nls[u]);
          if (cx.key_oob) pe = 0.f;
          pt[hh * 4 + u] = pe;
          // a causally masked score is a constant in the reference (modeling_gpt.py:89 `w*b`,
          // modeling_bloom.py:108 masked_fill): P still feeds dV, but no gradient reaches q.k
          ds[hh * 4 + u] = fut ? 0.f : pe * fmaf(__uint_as_float(rd[e]), cx.scale, nds[u]);
        }
      }
    }
    fb7_store_pair<BF16>(cx, c, g, pt, ds);
  }
}

// MODE bit 0: tiled dQ workspace; bit 1: P^T / dS^T (and the statistics) double-buffered, so the element math of
// query tile it+1 runs under the dQ / dV / dK MMAs of tile it (v2 waits for them before touching the tiles).
template <bool BF16, int MODE>
__global__ void __launch_bounds__(FB_THREADS, 1)
    attn_bwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                        const AttnBwdP bp) {
  const AttnP& p = bp.f;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sK = base, sV = base + FA_TILE;
  const uint32_t sQ = base + 2 * FA_TILE;   // 2 stages, each Q then dO
  constexpr bool DQT = (MODE & 1) != 0, PIPE = (MODE & 2) != 0;
  constexpr int N_TILES = PIPE ? 14 : 10;
  constexpr uint32_t PDS_STRIDE = PIPE ? 4 * FA_TILE : 0;  // buffer (it & 1) of the P^T / dS^T pair
  constexpr uint32_t STAT_STRIDE = PIPE ? 1024 : 0;
  const uint32_t sPT = base + 6 * FA_TILE;  // 2 panels
  const uint32_t sDS = base + 8 * FA_TILE;  // 2 panels
  const uint32_t bars = base + N_TILES * FA_TILE;
  const uint32_t kv_full = bars, qdo_full = bars + 8, qdo_empty = bars + 24, sdp_full = bars + 40,
                 pds_ready = bars + 48, mma_done = bars + 56, dkv_full = bars + 64, tmem_slot = bars + 72,
                 sdp_free = bars + 80, lse_s = bars + 128, del_s = bars + 640;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem + N_TILES * FA_TILE + 72);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv_tiles = (p.Sk + 127) / 128;
  const int kv_tile = blockIdx.x % n_kv_tiles;
  const int bh = blockIdx.x / n_kv_tiles;
  const int h = bh % p.H, b = bh / p.H;
  const int kv0 = kv_tile * 128;
  const int n_q_tiles = (p.Sq + 127) / 128;
  int i_start = 0;
  if (p.causal) {
    const bool full_sweep = p.first_valid && (p.first_valid[b] > p.off);
    if (!full_sweep) i_start = max(0, (kv0 - p.off) / 128);
  }
  const int n_it = max(0, n_q_tiles - i_start);
  CT_DBG_CTA(0);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(qdo_full + 8 * s, 1); mbar_init(qdo_empty + 8 * s, 1); }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, 256);
    mbar_init(pds_ready, 256);
    mbar_init(mma_done, 1);
    mbar_init(dkv_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t T_ST = tmem, T_DPT = tmem + 128, T_DV = tmem + 256, T_DK = tmem + 320, T_DQ = tmem + 384;

  if (warp == 0) {
    if (lane == 0 && n_it > 0) {
      mbar_expect_tx(kv_full, 2 * FA_TILE);
      tma_load_4d(sK, &tmK, kv_full, 0, kv0, h, b);
      tma_load_4d(sV, &tmV, kv_full, 0, kv0, h, b);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1, q0 = (i_start + it) * 128;
        mbar_wait(qdo_empty + 8 * s, ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(qdo_full + 8 * s, 2 * FA_TILE);
        tma_load_4d(sQ + s * 2 * FA_TILE, &tmQ, qdo_full + 8 * s, 0, q0, h, b);
        tma_load_4d(sQ + s * 2 * FA_TILE + FA_TILE, &tmDO, qdo_full + 8 * s, 0, q0, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_it > 0) {
      const uint32_t idesc_kk = umma_idesc_f16(BF16 ? 1 : 0, 0, 0, 128, 128);  // S^T, dP^T
      const uint32_t idesc_km = umma_idesc_f16(BF16 ? 1 : 0, 0, 1, 128, 64);   // dV, dK
      const uint32_t idesc_mm = umma_idesc_f16(BF16 ? 1 : 0, 1, 1, 128, 64);   // dQ
      auto issue_sdp = [&](int it) {
        const int s = it & 1;
        const uint32_t q = sQ + s * 2 * FA_TILE, d_o = q + FA_TILE;
        mbar_wait(qdo_full + 8 * s, (it >> 1) & 1);
        if (it > 0) mbar_wait(sdp_free, (it - 1) & 1);  // every compute thread holds S^T/dP^T(it-1) in registers
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)  // S^T[kv, q] = K[kv, d] . Q[q, d]
          umma_f16(T_ST, umma_smem_desc_sw128(sK + k * 32, 0, 1024), umma_smem_desc_sw128(q + k * 32, 0, 1024),
                   idesc_kk, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dP^T[kv, q] = V[kv, d] . dO[q, d]
          umma_f16(T_DPT, umma_smem_desc_sw128(sV + k * 32, 0, 1024),
                   umma_smem_desc_sw128(d_o + k * 32, 0, 1024), idesc_kk, k > 0);
        umma_commit(sdp_full);
      };
      mbar_wait(kv_full, 0);
      issue_sdp(0);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1;
        const uint32_t q = sQ + s * 2 * FA_TILE, d_o = q + FA_TILE;
        if (it + 1 < n_it) issue_sdp(it + 1);  // runs under the element math of tile it
        mbar_wait(pds_ready, it & 1);
        tc_fence_after();
        const uint32_t pt = sPT + (it & 1) * PDS_STRIDE, dst = sDS + (it & 1) * PDS_STRIDE;
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dQ[q, d] = dS[q, kv] . K[kv, d]  (A MN-major view of dS^T)
          umma_f16(T_DQ, umma_smem_desc_sw128(dst + k * 2048, FA_TILE, 1024),
                   umma_smem_desc_sw128(sK + k * 2048, 64 * 128, 1024), idesc_mm, k > 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dV[kv, d] += P^T[kv, q] . dO[q, d]   (B MN-major: rows = q)
          umma_f16(T_DV, umma_smem_desc_sw128(pt + (k >> 2) * FA_TILE + (k & 3) * 32, 0, 1024),
                   umma_smem_desc_sw128(d_o + k * 2048, 64 * 128, 1024), idesc_km, (it > 0 || k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dK[kv, d] += dS^T[kv, q] . Q[q, d]
          umma_f16(T_DK, umma_smem_desc_sw128(dst + (k >> 2) * FA_TILE + (k & 3) * 32, 0, 1024),
                   umma_smem_desc_sw128(q + k * 2048, 64 * 128, 1024), idesc_km, (it > 0 || k > 0));
        umma_commit(qdo_empty + 8 * s);
        umma_commit(mma_done);  // dQ(it) readable; P^T / dS^T tiles free for tile it+1
      }
      umma_commit(dkv_full);
    }
  } else {
    const int wq = warp & 3;
    const int rr = wq * 32 + lane;   // key row inside the tile (S^T) / query row (dQ)
    const int hf = (warp - 2) >> 2;  // which pair of 32-query chunks this warp owns
    const int jg = kv0 + rr;
    const uint32_t t_lane = (uint32_t)(wq * 32) << 16;
    FbCtx cx;
    cx.kb = (p.kbias2 && jg < p.Sk) ? __ldg(p.kbias2 + (int64_t)b * p.kb_sb + (int64_t)h * p.kb_sh + jg) : 0.f;
    cx.sl2 = p.sl2; cx.scale = p.scale; cx.cf2 = p.causal_fill2;
    cx.jg = jg; cx.off = p.off; cx.causal = p.causal; cx.key_oob = jg >= p.Sk;
    cx.lse_s = lse_s; cx.del_s = del_s; cx.sPT = sPT; cx.sDS = sDS; cx.rr = rr; cx.sw = rr & 7;
    // generic arithmetic for the whole warp when a key is masked (the reference's finite fill matters on
    // fully masked query rows) or the key tile is ragged
    const bool warp_generic = __any_sync(0xffffffffu, cx.kb < -1e30f) || (kv0 + 128 > p.Sk);
    const bool fill_is_ninf = p.causal_fill2 == -INFINITY;
    const float* lse_bh = p.lse2 + ((int64_t)b * p.H + h) * p.Sq;
    const float* del_bh = bp.delta + ((int64_t)b * p.H + h) * p.Sq;
    // per-query statistics of the current query tile, staged as -lse2 and -delta*scale
    float nlse_next = -INFINITY, ndel_next = 0.f;
    if (hf == 0 && n_it > 0 && i_start * 128 + rr < p.Sq) {
      nlse_next = -__ldg(lse_bh + i_start * 128 + rr);
      ndel_next = -__ldg(del_bh + i_start * 128 + rr) * p.scale;
    }
    const int c0 = 2 * hf, c1 = 2 * hf + 1;
    CT_DBG_CTA(1);

    // dQ rows of query tile `itp`: this thread owns query row (q0 + rr), columns [32*hf, 32*hf + 32) of d
    auto red_dq = [&](const uint32_t (&r)[32], int itp) {
      const int qi = (i_start + itp) * 128 + rr;
      if (qi < p.Sq) {
        float* dst;
        int gs;  // distance between consecutive 4-float groups of this row
        if constexpr (DQT) {
          dst = bp.dq_accum + (((int64_t)b * p.H + h) * n_q_tiles + (i_start + itp)) * FB_DQ_TILE +
                (hf * 8 * 128 + rr) * 4;
          gs = 128 * 4;
        } else {
          dst = bp.dq_accum + (((int64_t)b * p.Sq + qi) * p.H + h) * 64 + hf * 32;
          gs = 4;
        }
#pragma unroll
        for (int g = 0; g < 8; ++g)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + gs * g),
                       "f"(__uint_as_float(r[4 * g])), "f"(__uint_as_float(r[4 * g + 1])),
                       "f"(__uint_as_float(r[4 * g + 2])), "f"(__uint_as_float(r[4 * g + 3]))
                       : "memory");
      }
    };

    for (int it = 0; it < n_it; ++it) {
      const int q0 = (i_start + it) * 128;
      if constexpr (PIPE) {
        cx.sPT = sPT + (it & 1) * PDS_STRIDE; cx.sDS = sDS + (it & 1) * PDS_STRIDE;
        cx.lse_s = lse_s + (it & 1) * STAT_STRIDE; cx.del_s = del_s + (it & 1) * STAT_STRIDE;
      }
      CT_DBG_STAMP(16 * it + 0);
      mbar_wait(sdp_full, it & 1);
      CT_DBG_STAMP(16 * it + 1);
      tc_fence_after();
      // chunk kinds (warp-uniform)
      const bool touches_diag = p.causal && (kv0 + 127 > q0 + p.off);
      const bool aligned_diag = touches_diag && fill_is_ninf && (kv0 == q0 + p.off) && !warp_generic;
      int kind0, kind1;
      if (warp_generic || (touches_diag && !aligned_diag)) {
        kind0 = kind1 = 2;
      } else if (aligned_diag) {
        // key row 32*wq+l vs queries 32*c..32*c+31: c < wq entirely future, c > wq entirely visible
        kind0 = c0 < wq ? 1 : (c0 > wq ? 0 : 2);
        kind1 = c1 < wq ? 1 : (c1 > wq ? 0 : 2);
      } else {
        kind0 = kind1 = 0;
      }
      uint32_t rs0[32], rd0[32], rs1[32], rd1[32];
      if (kind0 != 1) { tmem_ld_32x32(T_ST + t_lane + c0 * 32, rs0); tmem_ld_32x32(T_DPT + t_lane + c0 * 32, rd0); }
      if (kind1 != 1) { tmem_ld_32x32(T_ST + t_lane + c1 * 32, rs1); tmem_ld_32x32(T_DPT + t_lane + c1 * 32, rd1); }
      if constexpr (!PIPE) {
        if (it > 0) {
          // all MMAs of tile it-1 retired: dQ(it-1) is complete and the P^T / dS^T tiles may be overwritten
          mbar_wait(mma_done, (it - 1) & 1);
          tc_fence_after();
        }
      }
      CT_DBG_STAMP(16 * it + 2);
      // v2: mma_done(it-1) implies pds_ready(it-1): every thread has finished the element math of tile it-1,
      // nobody still reads its statistics. PIPE: buffer (it & 1) was last read by tile it-2, and every thread
      // finished tile it-2 before it arrived at the named barrier of tile it-1, which this thread has passed.
      if (hf == 0) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(cx.lse_s + 4 * rr), "f"(nlse_next) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(cx.del_s + 4 * rr), "f"(ndel_next) : "memory");
      }
      tmem_ld_wait();
      CT_DBG_STAMP(16 * it + 3);
      tc_fence_before();
      mbar_arrive(sdp_free);        // S^T / dP^T are in registers: the next tile's MMAs may overwrite them
      bar_sync_named(1, 256);       // statistics staged by the hf == 0 warps are visible
      CT_DBG_STAMP(16 * it + 4);
      if (hf == 0) {
        const int nq = q0 + 128 + rr;
        const bool ok = (it + 1 < n_it) && nq < p.Sq;
        nlse_next = ok ? -__ldg(lse_bh + nq) : -INFINITY;
        ndel_next = ok ? -__ldg(del_bh + nq) * p.scale : 0.f;
      }
      if (kind0 == 0) fb2_chunk<0, BF16>(cx, rs0, rd0, c0, q0);
      else if (kind0 == 1) fb2_chunk<1, BF16>(cx, rs0, rd0, c0, q0);
      else fb2_chunk<2, BF16>(cx, rs0, rd0, c0, q0);
      CT_DBG_STAMP(16 * it + 5);
      if (it > 0) {
        // drain dQ(it-1) between the two chunks (T_DQ is only rewritten after pds_ready(it)): the
        // red.global.add traffic overlaps the second chunk's math
        if constexpr (PIPE) {
          // MMAs of tile it-1 ran under chunk 0. Waiting for every phase in order also proves that buffer
          // ((it+1) & 1) of P^T / dS^T — read by the MMAs of tile it-1 — is free when tile it+1 writes it.
          mbar_wait(mma_done, (it - 1) & 1);
          tc_fence_after();
          CT_DBG_STAMP(16 * it + 9);
        }
        uint32_t rq[32];
        tmem_ld_32x32(T_DQ + t_lane + hf * 32, rq);
        tmem_ld_wait();
        red_dq(rq, it - 1);
      }
      CT_DBG_STAMP(16 * it + 6);
      if (kind1 == 0) fb2_chunk<0, BF16>(cx, rs1, rd1, c1, q0);
      else if (kind1 == 1) fb2_chunk<1, BF16>(cx, rs1, rd1, c1, q0);
      else fb2_chunk<2, BF16>(cx, rs1, rd1, c1, q0);
      CT_DBG_STAMP(16 * it + 7);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_ready);
      CT_DBG_STAMP(16 * it + 8);
    }
    CT_DBG_CTA(2);
    // ---- last dQ tile, then dK / dV for this key row: 32 of the 64 head-dim columns per thread ----
    if (n_it > 0) {
      mbar_wait(mma_done, (n_it - 1) & 1);
      tc_fence_after();
      uint32_t rq[32];
      tmem_ld_32x32(T_DQ + t_lane + hf * 32, rq);
      tmem_ld_wait();
      red_dq(rq, n_it - 1);
      mbar_wait(dkv_full, 0);
      tc_fence_after();
    }
    CT_DBG_CTA(3);
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      uint32_t r[32];
      if (n_it > 0) {
        tmem_ld_32x32((which == 0 ? T_DV : T_DK) + t_lane + hf * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int t = 0; t < 32; ++t) r[t] = 0u;
      }
      if (!cx.key_oob) {
        void* basep = which == 0 ? bp.dv : bp.dk;
        const int64_t eo = which == 0
            ? (int64_t)b * bp.dv_sb + (int64_t)h * bp.dv_sh + (int64_t)jg * bp.dv_ss
            : (int64_t)b * bp.dk_sb + (int64_t)h * bp.dk_sh + (int64_t)jg * bp.dk_ss;
        uint8_t* row = reinterpret_cast<uint8_t*>(basep) + 2 * (eo + hf * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          const float f0 = __uint_as_float(r[8 * g]), f1 = __uint_as_float(r[8 * g + 1]),
                      f2 = __uint_as_float(r[8 * g + 2]), f3 = __uint_as_float(r[8 * g + 3]),
                      f4 = __uint_as_float(r[8 * g + 4]), f5 = __uint_as_float(r[8 * g + 5]),
                      f6 = __uint_as_float(r[8 * g + 6]), f7 = __uint_as_float(r[8 * g + 7]);
          if constexpr (BF16) {
            w.x = pack_bf16x2(f0, f1); w.y = pack_bf16x2(f2, f3); w.z = pack_bf16x2(f4, f5); w.w = pack_bf16x2(f6, f7);
          } else {
            __half2 x;
            x = __floats2half2_rn(f0, f1); w.x = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(f2, f3); w.y = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(f4, f5); w.z = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(f6, f7); w.w = *reinterpret_cast<uint32_t*>(&x);
          }
          *reinterpret_cast<uint4*>(row + 16 * g) = w;
        }
      }
    }
  }
  CT_DBG_CTA(4);
  tc_fence_before();
  __syncthreads();
  CT_DBG_CTA(5);
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// v7 (default): v3 with the two global-memory tails of the compute warps handed to the TMA engine: the dQ tile of every
// query tile leaves through shared memory as cp.reduce.async.bulk (fp32 add in the L2) instead of eight
// red.global.add.v4 per thread (r01f stamps: 750-1500 of ~4700 cycles per tile sat in that drain), and dK / dV leave
// as one TMA store per tile instead of 16-byte row-strided stores (3.5-4.9 k cycles of epilogue per CTA).
template <bool BF16>
__global__ void __launch_bounds__(FB_THREADS, 1)
    attn_bwd_tc7_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                        const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDK,
                        const __grid_constant__ CUtensorMap tmDV, const AttnBwdP bp) {
  const AttnP& p = bp.f;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sK = base, sV = base + FA_TILE;
  const uint32_t sQ = base + 2 * FA_TILE;   // 2 stages, each Q then dO
  constexpr bool PIPE = true;
  constexpr int N_TILES = PIPE ? 14 : 10;
  constexpr uint32_t PDS_STRIDE = PIPE ? 4 * FA_TILE : 0;  // buffer (it & 1) of the P^T / dS^T pair
  constexpr uint32_t STAT_STRIDE = PIPE ? 1024 : 0;
  const uint32_t sPT = base + 6 * FA_TILE;  // 2 panels
  const uint32_t sDS = base + 8 * FA_TILE;  // 2 panels
  const uint32_t bars = base + N_TILES * FA_TILE;
  const uint32_t kv_full = bars, qdo_full = bars + 8, qdo_empty = bars + 24, sdp_full = bars + 40,
                 pds_ready = bars + 48, mma_done = bars + 56, dkv_full = bars + 64, tmem_slot = bars + 72,
                 sdp_free = bars + 80, lse_s = bars + 128, del_s = bars + 640;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem + N_TILES * FA_TILE + 72);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv_tiles = (p.Sk + 127) / 128;
  const int kv_tile = blockIdx.x % n_kv_tiles;
  const int bh = blockIdx.x / n_kv_tiles;
  const int h = bh % p.H, b = bh / p.H;
  const int kv0 = kv_tile * 128;
  const int n_q_tiles = (p.Sq + 127) / 128;
  int i_start = 0;
  if (p.causal) {
    const bool full_sweep = p.first_valid && (p.first_valid[b] > p.off);
    if (!full_sweep) i_start = max(0, (kv0 - p.off) / 128);
  }
  const int n_it = max(0, n_q_tiles - i_start);
  CT_DBG_CTA(0);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(qdo_full + 8 * s, 1); mbar_init(qdo_empty + 8 * s, 1); }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, 256);
    mbar_init(pds_ready, 256);
    mbar_init(mma_done, 1);
    mbar_init(dkv_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t T_ST = tmem, T_DPT = tmem + 128, T_DV = tmem + 256, T_DK = tmem + 320, T_DQ = tmem + 384;

  if (warp == 0) {
    if (lane == 0 && n_it > 0) {
      mbar_expect_tx(kv_full, 2 * FA_TILE);
      tma_load_4d(sK, &tmK, kv_full, 0, kv0, h, b);
      tma_load_4d(sV, &tmV, kv_full, 0, kv0, h, b);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1, q0 = (i_start + it) * 128;
        mbar_wait(qdo_empty + 8 * s, ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(qdo_full + 8 * s, 2 * FA_TILE);
        tma_load_4d(sQ + s * 2 * FA_TILE, &tmQ, qdo_full + 8 * s, 0, q0, h, b);
        tma_load_4d(sQ + s * 2 * FA_TILE + FA_TILE, &tmDO, qdo_full + 8 * s, 0, q0, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_it > 0) {
      const uint32_t idesc_kk = umma_idesc_f16(BF16 ? 1 : 0, 0, 0, 128, 128);  // S^T, dP^T
      const uint32_t idesc_km = umma_idesc_f16(BF16 ? 1 : 0, 0, 1, 128, 64);   // dV, dK
      const uint32_t idesc_mm = umma_idesc_f16(BF16 ? 1 : 0, 1, 1, 128, 64);   // dQ
      auto issue_sdp = [&](int it) {
        const int s = it & 1;
        const uint32_t q = sQ + s * 2 * FA_TILE, d_o = q + FA_TILE;
        mbar_wait(qdo_full + 8 * s, (it >> 1) & 1);
        if (it > 0) mbar_wait(sdp_free, (it - 1) & 1);  // every compute thread holds S^T/dP^T(it-1) in registers
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)  // S^T[kv, q] = K[kv, d] . Q[q, d]
          umma_f16(T_ST, umma_smem_desc_sw128(sK + k * 32, 0, 1024), umma_smem_desc_sw128(q + k * 32, 0, 1024),
                   idesc_kk, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dP^T[kv, q] = V[kv, d] . dO[q, d]
          umma_f16(T_DPT, umma_smem_desc_sw128(sV + k * 32, 0, 1024),
                   umma_smem_desc_sw128(d_o + k * 32, 0, 1024), idesc_kk, k > 0);
        umma_commit(sdp_full);
      };
      mbar_wait(kv_full, 0);
      issue_sdp(0);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1;
        const uint32_t q = sQ + s * 2 * FA_TILE, d_o = q + FA_TILE;
        if (it + 1 < n_it) issue_sdp(it + 1);  // runs under the element math of tile it
        mbar_wait(pds_ready, it & 1);
        tc_fence_after();
        const uint32_t pt = sPT + (it & 1) * PDS_STRIDE, dst = sDS + (it & 1) * PDS_STRIDE;
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dQ[q, d] = dS[q, kv] . K[kv, d]  (A MN-major view of dS^T)
          umma_f16(T_DQ, umma_smem_desc_sw128(dst + k * 2048, FA_TILE, 1024),
                   umma_smem_desc_sw128(sK + k * 2048, 64 * 128, 1024), idesc_mm, k > 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dV[kv, d] += P^T[kv, q] . dO[q, d]   (B MN-major: rows = q)
          umma_f16(T_DV, umma_smem_desc_sw128(pt + (k >> 2) * FA_TILE + (k & 3) * 32, 0, 1024),
                   umma_smem_desc_sw128(d_o + k * 2048, 64 * 128, 1024), idesc_km, (it > 0 || k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dK[kv, d] += dS^T[kv, q] . Q[q, d]
          umma_f16(T_DK, umma_smem_desc_sw128(dst + (k >> 2) * FA_TILE + (k & 3) * 32, 0, 1024),
                   umma_smem_desc_sw128(q + k * 2048, 64 * 128, 1024), idesc_km, (it > 0 || k > 0));
        umma_commit(qdo_empty + 8 * s);
        umma_commit(mma_done);  // dQ(it) readable; P^T / dS^T tiles free for tile it+1
      }
      umma_commit(dkv_full);
    }
  } else {
    const int wq = warp & 3;
    const int rr = wq * 32 + lane;   // key row inside the tile (S^T) / query row (dQ)
    const int hf = (warp - 2) >> 2;  // which pair of 32-query chunks this warp owns
    const int jg = kv0 + rr;
    const uint32_t t_lane = (uint32_t)(wq * 32) << 16;
    FbCtx cx;
    cx.kb = (p.kbias2 && jg < p.Sk) ? __ldg(p.kbias2 + (int64_t)b * p.kb_sb + (int64_t)h * p.kb_sh + jg) : 0.f;
    cx.sl2 = p.sl2; cx.scale = p.scale; cx.cf2 = p.causal_fill2;
    cx.jg = jg; cx.off = p.off; cx.causal = p.causal; cx.key_oob = jg >= p.Sk;
    cx.lse_s = lse_s; cx.del_s = del_s; cx.sPT = sPT; cx.sDS = sDS; cx.rr = rr; cx.sw = rr & 7;
    // generic arithmetic for the whole warp when a key is masked (the reference's finite fill matters on
    // fully masked query rows) or the key tile is ragged
    const bool warp_generic = __any_sync(0xffffffffu, cx.kb < -1e30f) || (kv0 + 128 > p.Sk);
    const bool fill_is_ninf = p.causal_fill2 == -INFINITY;
    const float* lse_bh = p.lse2 + ((int64_t)b * p.H + h) * p.Sq;
    const float* del_bh = bp.delta + ((int64_t)b * p.H + h) * p.Sq;
    // per-query statistics of the current query tile, staged as -lse2 and -delta*scale
    float nlse_next = -INFINITY, ndel_next = 0.f;
    if (hf == 0 && n_it > 0 && i_start * 128 + rr < p.Sq) {
      nlse_next = -__ldg(lse_bh + i_start * 128 + rr);
      ndel_next = -__ldg(del_bh + i_start * 128 + rr) * p.scale;
    }
    const int c0 = 2 * hf, c1 = 2 * hf + 1;
    CT_DBG_CTA(1);

    // dQ rows of query tile `itp`: this thread owns query row rr, columns [32*hf, 32*hf + 32) of d. The 32 KB tile
    // goes through the dS^T buffer of that tile (free: every MMA of tile itp has retired) and leaves as two TMA
    // reduce-adds (cp.reduce.async.bulk: fp32 adds in the L2) into the [B,Sq,H,64] workspace — no per-thread
    // red.global.add, the LSU and the issue slots stay with the element math.
    auto red_dq = [&](const uint32_t (&r)[32], int itp) {
      const uint32_t stg = sDS + (itp & 1) * PDS_STRIDE;
      const uint32_t row = stg + hf * FA_TILE + rr * 128;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((g ^ cx.sw) << 4)), "r"(r[4 * g]),
                     "r"(r[4 * g + 1]), "r"(r[4 * g + 2]), "r"(r[4 * g + 3])
                     : "memory");
      fence_proxy_async_smem();
      bar_sync_named(2, 256);
      if (threadIdx.x == 64) {
        const int q0p = (i_start + itp) * 128;
        tma_reduce_add_4d(&tmDQ, stg, 0, q0p, h, b);
        tma_reduce_add_4d(&tmDQ, stg + FA_TILE, 32, q0p, h, b);
        tma_commit_group();
        tma_wait_group_read0();  // before this thread reaches the next tile's barrier: the buffer may then be rewritten
      }
    };

    for (int it = 0; it < n_it; ++it) {
      const int q0 = (i_start + it) * 128;
      if constexpr (PIPE) {
        cx.sPT = sPT + (it & 1) * PDS_STRIDE; cx.sDS = sDS + (it & 1) * PDS_STRIDE;
        cx.lse_s = lse_s + (it & 1) * STAT_STRIDE; cx.del_s = del_s + (it & 1) * STAT_STRIDE;
      }
      CT_DBG_STAMP(16 * it + 0);
      mbar_wait(sdp_full, it & 1);
      CT_DBG_STAMP(16 * it + 1);
      tc_fence_after();
      // chunk kinds (warp-uniform)
      const bool touches_diag = p.causal && (kv0 + 127 > q0 + p.off);
      const bool aligned_diag = touches_diag && fill_is_ninf && (kv0 == q0 + p.off) && !warp_generic;
      int kind0, kind1;
      if (warp_generic || (touches_diag && !aligned_diag)) {
        kind0 = kind1 = 2;
      } else if (aligned_diag) {
        // key row 32*wq+l vs queries 32*c..32*c+31: c < wq entirely future, c > wq entirely visible
        kind0 = c0 < wq ? 1 : (c0 > wq ? 0 : 2);
        kind1 = c1 < wq ? 1 : (c1 > wq ? 0 : 2);
      } else {
        kind0 = kind1 = 0;
      }
      uint32_t rs0[32], rd0[32], rs1[32], rd1[32];
      if (kind0 != 1) { tmem_ld_32x32(T_ST + t_lane + c0 * 32, rs0); tmem_ld_32x32(T_DPT + t_lane + c0 * 32, rd0); }
      if (kind1 != 1) { tmem_ld_32x32(T_ST + t_lane + c1 * 32, rs1); tmem_ld_32x32(T_DPT + t_lane + c1 * 32, rd1); }
      if constexpr (!PIPE) {
        if (it > 0) {
          // all MMAs of tile it-1 retired: dQ(it-1) is complete and the P^T / dS^T tiles may be overwritten
          mbar_wait(mma_done, (it - 1) & 1);
          tc_fence_after();
        }
      }
      CT_DBG_STAMP(16 * it + 2);
      // v2: mma_done(it-1) implies pds_ready(it-1): every thread has finished the element math of tile it-1,
      // nobody still reads its statistics. PIPE: buffer (it & 1) was last read by tile it-2, and every thread
      // finished tile it-2 before it arrived at the named barrier of tile it-1, which this thread has passed.
      if (hf == 0) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(cx.lse_s + 4 * rr), "f"(nlse_next) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(cx.del_s + 4 * rr), "f"(ndel_next) : "memory");
      }
      tmem_ld_wait();
      CT_DBG_STAMP(16 * it + 3);
      tc_fence_before();
      mbar_arrive(sdp_free);        // S^T / dP^T are in registers: the next tile's MMAs may overwrite them
      bar_sync_named(1, 256);       // statistics staged by the hf == 0 warps are visible
      CT_DBG_STAMP(16 * it + 4);
      if (hf == 0) {
        const int nq = q0 + 128 + rr;
        const bool ok = (it + 1 < n_it) && nq < p.Sq;
        nlse_next = ok ? -__ldg(lse_bh + nq) : -INFINITY;
        ndel_next = ok ? -__ldg(del_bh + nq) * p.scale : 0.f;
      }
      if (kind0 == 0) fb7_chunk<0, BF16>(cx, rs0, rd0, c0, q0);
      else if (kind0 == 1) fb7_chunk<1, BF16>(cx, rs0, rd0, c0, q0);
      else fb7_chunk<2, BF16>(cx, rs0, rd0, c0, q0);
      CT_DBG_STAMP(16 * it + 5);
      if (it > 0) {
        // drain dQ(it-1) between the two chunks (T_DQ is only rewritten after pds_ready(it)): the
        // red.global.add traffic overlaps the second chunk's math
        if constexpr (PIPE) {
          // MMAs of tile it-1 ran under chunk 0. Waiting for every phase in order also proves that buffer
          // ((it+1) & 1) of P^T / dS^T — read by the MMAs of tile it-1 — is free when tile it+1 writes it.
          mbar_wait(mma_done, (it - 1) & 1);
          tc_fence_after();
          CT_DBG_STAMP(16 * it + 9);
        }
        uint32_t rq[32];
        tmem_ld_32x32(T_DQ + t_lane + hf * 32, rq);
        tmem_ld_wait();
        red_dq(rq, it - 1);
      }
      CT_DBG_STAMP(16 * it + 6);
      if (kind1 == 0) fb7_chunk<0, BF16>(cx, rs1, rd1, c1, q0);
      else if (kind1 == 1) fb7_chunk<1, BF16>(cx, rs1, rd1, c1, q0);
      else fb7_chunk<2, BF16>(cx, rs1, rd1, c1, q0);
      CT_DBG_STAMP(16 * it + 7);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_ready);
      CT_DBG_STAMP(16 * it + 8);
    }
    CT_DBG_CTA(2);
    // ---- last dQ tile, then dK / dV for this key row: 32 of the 64 head-dim columns per thread ----
    if (n_it > 0) {
      mbar_wait(mma_done, (n_it - 1) & 1);
      tc_fence_after();
      uint32_t rq[32];
      tmem_ld_32x32(T_DQ + t_lane + hf * 32, rq);
      tmem_ld_wait();
      red_dq(rq, n_it - 1);
      mbar_wait(dkv_full, 0);
      tc_fence_after();
    }
    CT_DBG_CTA(3);
    // ---- dK / dV: TMEM -> registers -> swizzled 16 KB tiles in the (now idle) K / V buffers -> one TMA store each
    // (rows beyond Sk are clipped by the tensor map) instead of eight row-strided 16-byte stores per thread ----
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      uint32_t r[32];
      if (n_it > 0) {
        tmem_ld_32x32((which == 0 ? T_DV : T_DK) + t_lane + hf * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int t = 0; t < 32; ++t) r[t] = 0u;
      }
      const uint32_t row = (which == 0 ? sV : sK) + rr * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float f0 = __uint_as_float(r[8 * g + 2 * u]), f1 = __uint_as_float(r[8 * g + 2 * u + 1]);
          if constexpr (BF16) {
            w[u] = pack_bf16x2(f0, f1);
          } else {
            __half2 x = __floats2half2_rn(f0, f1);
            w[u] = *reinterpret_cast<uint32_t*>(&x);
          }
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + (((4 * hf + g) ^ cx.sw) << 4)), "r"(w[0]),
                     "r"(w[1]), "r"(w[2]), "r"(w[3])
                     : "memory");
      }
    }
    fence_proxy_async_smem();
    bar_sync_named(2, 256);
    if (threadIdx.x == 64) {
      tma_store_4d(&tmDV, sV, 0, kv0, h, b);
      tma_store_4d(&tmDK, sK, 0, kv0, h, b);
      tma_commit_group();
      tma_wait_group_read0();  // the CTA's shared memory must outlive the reads
    }
  }
  CT_DBG_CTA(4);
  tc_fence_before();
  __syncthreads();
  CT_DBG_CTA(5);
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// v4: v3 (tiled dQ workspace, P^T / dS^T double-buffered) with the dQ drain moved off the compute warps. The
// r01f stamps put 900-1900 of v3's ~4700 cycles per query tile into the eight red.global.add.v4 per compute
// thread (LSU-bound: the warp sits in the issue queue while its MUFU / FMA work waits). Here a fourth warpgroup
// owns the drain: T_DQ is double-buffered in TMEM (the last 64 free columns), the drain warps pull tile it out
// as soon as its MMAs retire and release the buffer before issuing their reds, and the compute warps never touch
// dQ. 512 threads = 4 warpgroups; setmaxnreg moves registers from the TMA/MMA and drain groups to the compute
// groups (launch bound 128/thread -> 40 / 72 / 200).
constexpr int FB3_THREADS = 512;
template <uint32_t N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <uint32_t N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <bool BF16>
__global__ void __launch_bounds__(FB3_THREADS, 1)
    attn_bwd_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                        const AttnBwdP bp) {
  const AttnP& p = bp.f;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sK = base, sV = base + FA_TILE;
  const uint32_t sQ = base + 2 * FA_TILE;   // 2 stages, each Q then dO
  constexpr int N_TILES = 14;
  constexpr uint32_t PDS_STRIDE = 4 * FA_TILE;  // buffer (it & 1) of the P^T / dS^T pair
  constexpr uint32_t STAT_STRIDE = 1024;
  const uint32_t sPT = base + 6 * FA_TILE;  // 2 panels
  const uint32_t sDS = base + 8 * FA_TILE;  // 2 panels
  const uint32_t bars = base + N_TILES * FA_TILE;
  const uint32_t kv_full = bars, qdo_full = bars + 8, qdo_empty = bars + 24, sdp_full = bars + 40,
                 pds_ready = bars + 48, dkv_full = bars + 64, tmem_slot = bars + 72, sdp_free = bars + 80,
                 tile_done = bars + 88 /* 2: all MMAs of tile it retired (buffer it & 1) */,
                 dq_free = bars + 104 /* 2: T_DQ[it & 1] is in the drain warps' registers */,
                 lse_s = bars + 128, del_s = bars + 640;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem + N_TILES * FA_TILE + 72);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv_tiles = (p.Sk + 127) / 128;
  const int kv_tile = blockIdx.x % n_kv_tiles;
  const int bh = blockIdx.x / n_kv_tiles;
  const int h = bh % p.H, b = bh / p.H;
  const int kv0 = kv_tile * 128;
  const int n_q_tiles = (p.Sq + 127) / 128;
  int i_start = 0;
  if (p.causal) {
    const bool full_sweep = p.first_valid && (p.first_valid[b] > p.off);
    if (!full_sweep) i_start = max(0, (kv0 - p.off) / 128);
  }
  const int n_it = max(0, n_q_tiles - i_start);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(qdo_full + 8 * s, 1); mbar_init(qdo_empty + 8 * s, 1); }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, 256);
    mbar_init(pds_ready, 256);
    for (int s = 0; s < 2; ++s) { mbar_init(tile_done + 8 * s, 1); mbar_init(dq_free + 8 * s, 128); }
    mbar_init(dkv_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t T_ST = tmem, T_DPT = tmem + 128, T_DV = tmem + 256, T_DK = tmem + 320, T_DQ = tmem + 384;  // 2 x 64

  // setmaxnreg inside each role branch: ptxas sizes a region's registers by the setmaxnreg that dominates it
  const int wg = warp >> 2;
  if (wg == 3) {
    setmaxnreg_dec<72>();
    // ------------------------------ dQ drain: one thread per query row, all 64 columns ------------------------------
    const int wq = warp & 3;
    const int rr = wq * 32 + lane;
    const uint32_t t_lane = (uint32_t)(wq * 32) << 16;
    for (int it = 0; it < n_it; ++it) {
      const int s = it & 1;
      mbar_wait(tile_done + 8 * s, (it >> 1) & 1);
      tc_fence_after();
      const int qi = (i_start + it) * 128 + rr;
      float* dst = bp.dq_accum + (((int64_t)b * p.H + h) * n_q_tiles + (i_start + it)) * FB_DQ_TILE + rr * 4;
      uint32_t r[32];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        tmem_ld_32x32(T_DQ + s * 64 + t_lane + half * 32, r);
        tmem_ld_wait();
        if (half == 1) {
          tc_fence_before();
          mbar_arrive(dq_free + 8 * s);  // dQ(it+2) may overwrite the buffer; the reds below run under later tiles
        }
        if (qi < p.Sq) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 512 * (8 * half + g)),
                         "f"(__uint_as_float(r[4 * g])), "f"(__uint_as_float(r[4 * g + 1])),
                         "f"(__uint_as_float(r[4 * g + 2])), "f"(__uint_as_float(r[4 * g + 3]))
                         : "memory");
        }
      }
    }
  } else if (wg == 0) {
   setmaxnreg_dec<56>();
   if (warp == 0) {
    if (lane == 0 && n_it > 0) {
      mbar_expect_tx(kv_full, 2 * FA_TILE);
      tma_load_4d(sK, &tmK, kv_full, 0, kv0, h, b);
      tma_load_4d(sV, &tmV, kv_full, 0, kv0, h, b);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1, q0 = (i_start + it) * 128;
        mbar_wait(qdo_empty + 8 * s, ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(qdo_full + 8 * s, 2 * FA_TILE);
        tma_load_4d(sQ + s * 2 * FA_TILE, &tmQ, qdo_full + 8 * s, 0, q0, h, b);
        tma_load_4d(sQ + s * 2 * FA_TILE + FA_TILE, &tmDO, qdo_full + 8 * s, 0, q0, h, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_it > 0) {
      const uint32_t idesc_kk = umma_idesc_f16(BF16 ? 1 : 0, 0, 0, 128, 128);  // S^T, dP^T
      const uint32_t idesc_km = umma_idesc_f16(BF16 ? 1 : 0, 0, 1, 128, 64);   // dV, dK
      const uint32_t idesc_mm = umma_idesc_f16(BF16 ? 1 : 0, 1, 1, 128, 64);   // dQ
      auto issue_sdp = [&](int it) {
        const int s = it & 1;
        const uint32_t q = sQ + s * 2 * FA_TILE, d_o = q + FA_TILE;
        mbar_wait(qdo_full + 8 * s, (it >> 1) & 1);
        if (it > 0) mbar_wait(sdp_free, (it - 1) & 1);  // every compute thread holds S^T/dP^T(it-1) in registers
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)  // S^T[kv, q] = K[kv, d] . Q[q, d]
          umma_f16(T_ST, umma_smem_desc_sw128(sK + k * 32, 0, 1024), umma_smem_desc_sw128(q + k * 32, 0, 1024),
                   idesc_kk, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dP^T[kv, q] = V[kv, d] . dO[q, d]
          umma_f16(T_DPT, umma_smem_desc_sw128(sV + k * 32, 0, 1024),
                   umma_smem_desc_sw128(d_o + k * 32, 0, 1024), idesc_kk, k > 0);
        umma_commit(sdp_full);
      };
      mbar_wait(kv_full, 0);
      issue_sdp(0);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1;
        const uint32_t q = sQ + s * 2 * FA_TILE, d_o = q + FA_TILE;
        if (it + 1 < n_it) issue_sdp(it + 1);  // runs under the element math of tile it
        mbar_wait(pds_ready, it & 1);
        tc_fence_after();
        const uint32_t pt = sPT + (it & 1) * PDS_STRIDE, dst = sDS + (it & 1) * PDS_STRIDE;
        if (it >= 2) {  // dQ(it-2) has left T_DQ[it & 1]
          mbar_wait(dq_free + 8 * s, ((it >> 1) - 1) & 1);
          tc_fence_after();
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dQ[q, d] = dS[q, kv] . K[kv, d]  (A MN-major view of dS^T)
          umma_f16(T_DQ + s * 64, umma_smem_desc_sw128(dst + k * 2048, FA_TILE, 1024),
                   umma_smem_desc_sw128(sK + k * 2048, 64 * 128, 1024), idesc_mm, k > 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dV[kv, d] += P^T[kv, q] . dO[q, d]   (B MN-major: rows = q)
          umma_f16(T_DV, umma_smem_desc_sw128(pt + (k >> 2) * FA_TILE + (k & 3) * 32, 0, 1024),
                   umma_smem_desc_sw128(d_o + k * 2048, 64 * 128, 1024), idesc_km, (it > 0 || k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dK[kv, d] += dS^T[kv, q] . Q[q, d]
          umma_f16(T_DK, umma_smem_desc_sw128(dst + (k >> 2) * FA_TILE + (k & 3) * 32, 0, 1024),
                   umma_smem_desc_sw128(q + k * 2048, 64 * 128, 1024), idesc_km, (it > 0 || k > 0));
        umma_commit(qdo_empty + 8 * s);
        umma_commit(tile_done + 8 * s);  // dQ(it) readable; P^T / dS^T buffer (it & 1) free for tile it+2
      }
      umma_commit(dkv_full);
    }
   }
  } else {
    setmaxnreg_inc<184>();
    const int wq = warp & 3;
    const int rr = wq * 32 + lane;   // key row inside the tile (S^T)
    const int hf = wg - 1;           // which pair of 32-query chunks this warp owns
    const int jg = kv0 + rr;
    const uint32_t t_lane = (uint32_t)(wq * 32) << 16;
    FbCtx cx;
    cx.kb = (p.kbias2 && jg < p.Sk) ? __ldg(p.kbias2 + (int64_t)b * p.kb_sb + (int64_t)h * p.kb_sh + jg) : 0.f;
    cx.sl2 = p.sl2; cx.scale = p.scale; cx.cf2 = p.causal_fill2;
    cx.jg = jg; cx.off = p.off; cx.causal = p.causal; cx.key_oob = jg >= p.Sk;
    cx.lse_s = lse_s; cx.del_s = del_s; cx.sPT = sPT; cx.sDS = sDS; cx.rr = rr; cx.sw = rr & 7;
    // generic arithmetic for the whole warp when a key is masked (the reference's finite fill matters on
    // fully masked query rows) or the key tile is ragged
    const bool warp_generic = __any_sync(0xffffffffu, cx.kb < -1e30f) || (kv0 + 128 > p.Sk);
    const bool fill_is_ninf = p.causal_fill2 == -INFINITY;
    const float* lse_bh = p.lse2 + ((int64_t)b * p.H + h) * p.Sq;
    const float* del_bh = bp.delta + ((int64_t)b * p.H + h) * p.Sq;
    // per-query statistics of the current query tile, staged as -lse2 and -delta*scale
    float nlse_next = -INFINITY, ndel_next = 0.f;
    if (hf == 0 && n_it > 0 && i_start * 128 + rr < p.Sq) {
      nlse_next = -__ldg(lse_bh + i_start * 128 + rr);
      ndel_next = -__ldg(del_bh + i_start * 128 + rr) * p.scale;
    }
    const int c0 = 2 * hf, c1 = 2 * hf + 1;

    for (int it = 0; it < n_it; ++it) {
      const int q0 = (i_start + it) * 128;
      cx.sPT = sPT + (it & 1) * PDS_STRIDE; cx.sDS = sDS + (it & 1) * PDS_STRIDE;
      cx.lse_s = lse_s + (it & 1) * STAT_STRIDE; cx.del_s = del_s + (it & 1) * STAT_STRIDE;
      CT_DBG_STAMP(16 * it + 0);
      mbar_wait(sdp_full, it & 1);
      CT_DBG_STAMP(16 * it + 1);
      tc_fence_after();
      // chunk kinds (warp-uniform)
      const bool touches_diag = p.causal && (kv0 + 127 > q0 + p.off);
      const bool aligned_diag = touches_diag && fill_is_ninf && (kv0 == q0 + p.off) && !warp_generic;
      int kind0, kind1;
      if (warp_generic || (touches_diag && !aligned_diag)) {
        kind0 = kind1 = 2;
      } else if (aligned_diag) {
        // key row 32*wq+l vs queries 32*c..32*c+31: c < wq entirely future, c > wq entirely visible
        kind0 = c0 < wq ? 1 : (c0 > wq ? 0 : 2);
        kind1 = c1 < wq ? 1 : (c1 > wq ? 0 : 2);
      } else {
        kind0 = kind1 = 0;
      }
      uint32_t rs0[32], rd0[32], rs1[32], rd1[32];
      if (kind0 != 1) { tmem_ld_32x32(T_ST + t_lane + c0 * 32, rs0); tmem_ld_32x32(T_DPT + t_lane + c0 * 32, rd0); }
      if (kind1 != 1) { tmem_ld_32x32(T_ST + t_lane + c1 * 32, rs1); tmem_ld_32x32(T_DPT + t_lane + c1 * 32, rd1); }
      CT_DBG_STAMP(16 * it + 2);
      // statistics buffer (it & 1) was last read by tile it-2, and every thread
      // finished tile it-2 before it arrived at the named barrier of tile it-1, which this thread has passed.
      if (hf == 0) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(cx.lse_s + 4 * rr), "f"(nlse_next) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(cx.del_s + 4 * rr), "f"(ndel_next) : "memory");
      }
      tmem_ld_wait();
      CT_DBG_STAMP(16 * it + 3);
      tc_fence_before();
      mbar_arrive(sdp_free);        // S^T / dP^T are in registers: the next tile's MMAs may overwrite them
      bar_sync_named(1, 256);       // statistics staged by the hf == 0 warps are visible
      CT_DBG_STAMP(16 * it + 4);
      if (hf == 0) {
        const int nq = q0 + 128 + rr;
        const bool ok = (it + 1 < n_it) && nq < p.Sq;
        nlse_next = ok ? -__ldg(lse_bh + nq) : -INFINITY;
        ndel_next = ok ? -__ldg(del_bh + nq) * p.scale : 0.f;
      }
      if (kind0 == 0) fb2_chunk<0, BF16>(cx, rs0, rd0, c0, q0);
      else if (kind0 == 1) fb2_chunk<1, BF16>(cx, rs0, rd0, c0, q0);
      else fb2_chunk<2, BF16>(cx, rs0, rd0, c0, q0);
      CT_DBG_STAMP(16 * it + 5);
      if (it > 0) {
        // MMAs of tile it-1 ran under chunk 0; buffer ((it+1) & 1) of P^T / dS^T, which they read, is written
        // by tile it+1. (The barrier cannot run a phase ahead: its next completion needs pds_ready(it+1).)
        mbar_wait(tile_done + 8 * ((it - 1) & 1), ((it - 1) >> 1) & 1);
        tc_fence_after();
      }
      CT_DBG_STAMP(16 * it + 6);
      if (kind1 == 0) fb2_chunk<0, BF16>(cx, rs1, rd1, c1, q0);
      else if (kind1 == 1) fb2_chunk<1, BF16>(cx, rs1, rd1, c1, q0);
      else fb2_chunk<2, BF16>(cx, rs1, rd1, c1, q0);
      CT_DBG_STAMP(16 * it + 7);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_ready);
      CT_DBG_STAMP(16 * it + 8);
    }
    // ---- dK / dV for this key row: 32 of the 64 head-dim columns per thread ----
    if (n_it > 0) {
      mbar_wait(dkv_full, 0);
      tc_fence_after();
    }
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      uint32_t r[32];
      if (n_it > 0) {
        tmem_ld_32x32((which == 0 ? T_DV : T_DK) + t_lane + hf * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int t = 0; t < 32; ++t) r[t] = 0u;
      }
      if (!cx.key_oob) {
        void* basep = which == 0 ? bp.dv : bp.dk;
        const int64_t eo = which == 0
            ? (int64_t)b * bp.dv_sb + (int64_t)h * bp.dv_sh + (int64_t)jg * bp.dv_ss
            : (int64_t)b * bp.dk_sb + (int64_t)h * bp.dk_sh + (int64_t)jg * bp.dk_ss;
        uint8_t* row = reinterpret_cast<uint8_t*>(basep) + 2 * (eo + hf * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          const float f0 = __uint_as_float(r[8 * g]), f1 = __uint_as_float(r[8 * g + 1]),
                      f2 = __uint_as_float(r[8 * g + 2]), f3 = __uint_as_float(r[8 * g + 3]),
                      f4 = __uint_as_float(r[8 * g + 4]), f5 = __uint_as_float(r[8 * g + 5]),
                      f6 = __uint_as_float(r[8 * g + 6]), f7 = __uint_as_float(r[8 * g + 7]);
          if constexpr (BF16) {
            w.x = pack_bf16x2(f0, f1); w.y = pack_bf16x2(f2, f3); w.z = pack_bf16x2(f4, f5); w.w = pack_bf16x2(f6, f7);
          } else {
            __half2 x;
            x = __floats2half2_rn(f0, f1); w.x = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(f2, f3); w.y = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(f4, f5); w.z = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(f6, f7); w.w = *reinterpret_cast<uint32_t*>(&x);
          }
          *reinterpret_cast<uint4*>(row + 16 * g) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// v5: persistent v3. r01g's per-CTA timeline puts ~9 k of the 31 k cycles of a 4-tile CTA into work nothing
// overlaps at one CTA per SM: set-up and first loads (2.7 k), the wait for the last dQ / dV / dK MMAs (2.2 k)
// and the dQ drain + dK / dV epilogue (4 k). Here a CTA loops over work items (kv tile, b, h) — kv tiles in
// ascending order = heaviest first under the causal mask — with every barrier phase derived from RUNNING tile /
// item counters that all roles advance identically, so that
//   * the TMA warp requests the next item's K / V the moment the last MMAs of the current item retire, and its
//     Q / dO tiles through the same two-stage ring,
//   * the MMA warp issues the next item's first S^T / dP^T right behind them (the score buffers were released
//     when the compute warps took the last tile into registers),
//   * the compute warps' epilogue (last dQ drain, dV / dK out of TMEM) runs under both; the first dV / dK MMA of
//     the next item waits until the accumulators have been read (dkv_free).
template <bool BF16>
__global__ void __launch_bounds__(FB_THREADS, 1)
    attn_bwd_tc5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                        const AttnBwdP bp, const int n_items) {
  const AttnP& p = bp.f;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sK = base, sV = base + FA_TILE;
  const uint32_t sQ = base + 2 * FA_TILE;   // 2 stages, each Q then dO
  constexpr int N_TILES = 14;
  constexpr uint32_t PDS_STRIDE = 4 * FA_TILE;  // buffer (g & 1) of the P^T / dS^T pair
  constexpr uint32_t STAT_STRIDE = 1024;
  const uint32_t sPT = base + 6 * FA_TILE;  // 2 panels
  const uint32_t sDS = base + 8 * FA_TILE;  // 2 panels
  const uint32_t bars = base + N_TILES * FA_TILE;
  const uint32_t kv_full = bars, qdo_full = bars + 8, qdo_empty = bars + 24, sdp_full = bars + 40,
                 pds_ready = bars + 48, mma_done = bars + 56, dkv_full = bars + 64, tmem_slot = bars + 72,
                 sdp_free = bars + 80, dkv_free = bars + 88, lse_s = bars + 128, del_s = bars + 640;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem + N_TILES * FA_TILE + 72);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_q_tiles = (p.Sq + 127) / 128;
  const int BH = p.B * p.H;
  // work item -> coordinates; identical in every role
  auto coords = [&](int item, int& kv0, int& b, int& h, int& i_start, int& n_it) {
    const int kv_tile = item / BH, bh = item - kv_tile * BH;
    h = bh % p.H; b = bh / p.H;
    kv0 = kv_tile * 128;
    i_start = 0;
    if (p.causal) {
      const bool full_sweep = p.first_valid && (p.first_valid[b] > p.off);
      if (!full_sweep) i_start = max(0, (kv0 - p.off) / 128);
    }
    n_it = max(0, n_q_tiles - i_start);
  };

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(qdo_full + 8 * s, 1); mbar_init(qdo_empty + 8 * s, 1); }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, 256);
    mbar_init(pds_ready, 256);
    mbar_init(mma_done, 1);
    mbar_init(dkv_full, 1);
    mbar_init(dkv_free, 256);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t T_ST = tmem, T_DPT = tmem + 128, T_DV = tmem + 256, T_DK = tmem + 320, T_DQ = tmem + 384;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0, I = 0;  // query tiles / items (with work) requested so far
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int kv0, b, h, i_start, n_it;
        coords(item, kv0, b, h, i_start, n_it);
        if (n_it == 0) continue;
        if (I > 0) mbar_wait(dkv_full, (I - 1) & 1);  // the previous item's last MMAs retired: K / V are free
        mbar_expect_tx(kv_full, 2 * FA_TILE);
        tma_load_4d(sK, &tmK, kv_full, 0, kv0, h, b);
        tma_load_4d(sV, &tmV, kv_full, 0, kv0, h, b);
        for (int it = 0; it < n_it; ++it, ++g) {
          const uint32_t s = g & 1;
          const int q0 = (i_start + it) * 128;
          mbar_wait(qdo_empty + 8 * s, ((g >> 1) & 1) ^ 1);
          mbar_expect_tx(qdo_full + 8 * s, 2 * FA_TILE);
          tma_load_4d(sQ + s * 2 * FA_TILE, &tmQ, qdo_full + 8 * s, 0, q0, h, b);
          tma_load_4d(sQ + s * 2 * FA_TILE + FA_TILE, &tmDO, qdo_full + 8 * s, 0, q0, h, b);
        }
        ++I;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_kk = umma_idesc_f16(BF16 ? 1 : 0, 0, 0, 128, 128);  // S^T, dP^T
      const uint32_t idesc_km = umma_idesc_f16(BF16 ? 1 : 0, 0, 1, 128, 64);   // dV, dK
      const uint32_t idesc_mm = umma_idesc_f16(BF16 ? 1 : 0, 1, 1, 128, 64);   // dQ
      auto issue_sdp = [&](uint32_t gq) {
        const uint32_t s = gq & 1;
        const uint32_t q = sQ + s * 2 * FA_TILE, d_o = q + FA_TILE;
        mbar_wait(qdo_full + 8 * s, (gq >> 1) & 1);
        if (gq > 0) mbar_wait(sdp_free, (gq - 1) & 1);  // every compute thread holds S^T/dP^T(gq-1) in registers
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)  // S^T[kv, q] = K[kv, d] . Q[q, d]
          umma_f16(T_ST, umma_smem_desc_sw128(sK + k * 32, 0, 1024), umma_smem_desc_sw128(q + k * 32, 0, 1024),
                   idesc_kk, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dP^T[kv, q] = V[kv, d] . dO[q, d]
          umma_f16(T_DPT, umma_smem_desc_sw128(sV + k * 32, 0, 1024),
                   umma_smem_desc_sw128(d_o + k * 32, 0, 1024), idesc_kk, k > 0);
        umma_commit(sdp_full);
      };
      uint32_t g = 0, I = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int kv0, b, h, i_start, n_it;
        coords(item, kv0, b, h, i_start, n_it);
        if (n_it == 0) continue;
        mbar_wait(kv_full, I & 1);
        issue_sdp(g);
        for (int it = 0; it < n_it; ++it) {
          const uint32_t gq = g + it, s = gq & 1;
          const uint32_t q = sQ + s * 2 * FA_TILE, d_o = q + FA_TILE;
          if (it + 1 < n_it) issue_sdp(gq + 1);  // runs under the element math of tile gq
          mbar_wait(pds_ready, gq & 1);
          if (it == 0 && I > 0) mbar_wait(dkv_free, (I - 1) & 1);  // the previous item's dV / dK have left TMEM
          tc_fence_after();
          const uint32_t pt = sPT + s * PDS_STRIDE, dst = sDS + s * PDS_STRIDE;
#pragma unroll
          for (int k = 0; k < 8; ++k)  // dQ[q, d] = dS[q, kv] . K[kv, d]  (A MN-major view of dS^T)
            umma_f16(T_DQ, umma_smem_desc_sw128(dst + k * 2048, FA_TILE, 1024),
                     umma_smem_desc_sw128(sK + k * 2048, 64 * 128, 1024), idesc_mm, k > 0);
#pragma unroll
          for (int k = 0; k < 8; ++k)  // dV[kv, d] += P^T[kv, q] . dO[q, d]   (B MN-major: rows = q)
            umma_f16(T_DV, umma_smem_desc_sw128(pt + (k >> 2) * FA_TILE + (k & 3) * 32, 0, 1024),
                     umma_smem_desc_sw128(d_o + k * 2048, 64 * 128, 1024), idesc_km, (it > 0 || k > 0));
#pragma unroll
          for (int k = 0; k < 8; ++k)  // dK[kv, d] += dS^T[kv, q] . Q[q, d]
            umma_f16(T_DK, umma_smem_desc_sw128(dst + (k >> 2) * FA_TILE + (k & 3) * 32, 0, 1024),
                     umma_smem_desc_sw128(q + k * 2048, 64 * 128, 1024), idesc_km, (it > 0 || k > 0));
          umma_commit(qdo_empty + 8 * s);
          umma_commit(mma_done);  // dQ(gq) readable; P^T / dS^T buffer (gq & 1) free for tile gq+2
        }
        umma_commit(dkv_full);
        g += n_it;
        ++I;
      }
    }
  } else {
    const int wq = warp & 3;
    const int rr = wq * 32 + lane;   // key row inside the tile (S^T) / query row (dQ)
    const int hf = (warp - 2) >> 2;  // which pair of 32-query chunks this warp owns
    const uint32_t t_lane = (uint32_t)(wq * 32) << 16;
    const bool fill_is_ninf = p.causal_fill2 == -INFINITY;
    const int c0 = 2 * hf, c1 = 2 * hf + 1;
    uint32_t g = 0, I = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int kv0, b, h, i_start, n_it;
      coords(item, kv0, b, h, i_start, n_it);
      const int jg = kv0 + rr;
      FbCtx cx;
      cx.kb = (p.kbias2 && jg < p.Sk) ? __ldg(p.kbias2 + (int64_t)b * p.kb_sb + (int64_t)h * p.kb_sh + jg) : 0.f;
      cx.sl2 = p.sl2; cx.scale = p.scale; cx.cf2 = p.causal_fill2;
      cx.jg = jg; cx.off = p.off; cx.causal = p.causal; cx.key_oob = jg >= p.Sk;
      cx.lse_s = lse_s; cx.del_s = del_s; cx.sPT = sPT; cx.sDS = sDS; cx.rr = rr; cx.sw = rr & 7;
      // generic arithmetic for the whole warp when a key is masked (the reference's finite fill matters on
      // fully masked query rows) or the key tile is ragged
      const bool warp_generic = __any_sync(0xffffffffu, cx.kb < -1e30f) || (kv0 + 128 > p.Sk);
      const float* lse_bh = p.lse2 + ((int64_t)b * p.H + h) * p.Sq;
      const float* del_bh = bp.delta + ((int64_t)b * p.H + h) * p.Sq;
      // per-query statistics of the first query tile, staged as -lse2 and -delta*scale
      float nlse_next = -INFINITY, ndel_next = 0.f;
      if (hf == 0 && n_it > 0 && i_start * 128 + rr < p.Sq) {
        nlse_next = -__ldg(lse_bh + i_start * 128 + rr);
        ndel_next = -__ldg(del_bh + i_start * 128 + rr) * p.scale;
      }
      // dQ rows of query tile `itp`: this thread owns query row (q0 + rr), columns [32*hf, 32*hf + 32) of d
      auto red_dq = [&](const uint32_t (&r)[32], int itp) {
        const int qi = (i_start + itp) * 128 + rr;
        if (qi < p.Sq) {
          float* dst = bp.dq_accum + (((int64_t)b * p.H + h) * n_q_tiles + (i_start + itp)) * FB_DQ_TILE +
                       (hf * 8 * 128 + rr) * 4;
#pragma unroll
          for (int gg = 0; gg < 8; ++gg)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 512 * gg),
                         "f"(__uint_as_float(r[4 * gg])), "f"(__uint_as_float(r[4 * gg + 1])),
                         "f"(__uint_as_float(r[4 * gg + 2])), "f"(__uint_as_float(r[4 * gg + 3]))
                         : "memory");
        }
      };

      for (int it = 0; it < n_it; ++it) {
        const uint32_t gq = g + it;
        const int q0 = (i_start + it) * 128;
        cx.sPT = sPT + (gq & 1) * PDS_STRIDE; cx.sDS = sDS + (gq & 1) * PDS_STRIDE;
        cx.lse_s = lse_s + (gq & 1) * STAT_STRIDE; cx.del_s = del_s + (gq & 1) * STAT_STRIDE;
        mbar_wait(sdp_full, gq & 1);
        tc_fence_after();
        // chunk kinds (warp-uniform)
        const bool touches_diag = p.causal && (kv0 + 127 > q0 + p.off);
        const bool aligned_diag = touches_diag && fill_is_ninf && (kv0 == q0 + p.off) && !warp_generic;
        int kind0, kind1;
        if (warp_generic || (touches_diag && !aligned_diag)) {
          kind0 = kind1 = 2;
        } else if (aligned_diag) {
          // key row 32*wq+l vs queries 32*c..32*c+31: c < wq entirely future, c > wq entirely visible
          kind0 = c0 < wq ? 1 : (c0 > wq ? 0 : 2);
          kind1 = c1 < wq ? 1 : (c1 > wq ? 0 : 2);
        } else {
          kind0 = kind1 = 0;
        }
        uint32_t rs0[32], rd0[32], rs1[32], rd1[32];
        if (kind0 != 1) { tmem_ld_32x32(T_ST + t_lane + c0 * 32, rs0); tmem_ld_32x32(T_DPT + t_lane + c0 * 32, rd0); }
        if (kind1 != 1) { tmem_ld_32x32(T_ST + t_lane + c1 * 32, rs1); tmem_ld_32x32(T_DPT + t_lane + c1 * 32, rd1); }
        // statistics buffer (gq & 1) was last read by tile gq-2, and every thread finished tile gq-2 before it
        // arrived at the named barrier of tile gq-1, which this thread has passed (also across items)
        if (hf == 0) {
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(cx.lse_s + 4 * rr), "f"(nlse_next) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(cx.del_s + 4 * rr), "f"(ndel_next) : "memory");
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(sdp_free);        // S^T / dP^T are in registers: the next tile's MMAs may overwrite them
        bar_sync_named(1, 256);       // statistics staged by the hf == 0 warps are visible
        if (hf == 0) {
          const int nq = q0 + 128 + rr;
          const bool ok = (it + 1 < n_it) && nq < p.Sq;
          nlse_next = ok ? -__ldg(lse_bh + nq) : -INFINITY;
          ndel_next = ok ? -__ldg(del_bh + nq) * p.scale : 0.f;
        }
        if (kind0 == 0) fb2_chunk<0, BF16>(cx, rs0, rd0, c0, q0);
        else if (kind0 == 1) fb2_chunk<1, BF16>(cx, rs0, rd0, c0, q0);
        else fb2_chunk<2, BF16>(cx, rs0, rd0, c0, q0);
        if (it > 0) {
          // MMAs of tile gq-1 ran under chunk 0; waiting for every phase in order (the last tile of an item is
          // waited for in its epilogue) also proves that buffer ((gq+1) & 1) of P^T / dS^T is free for tile gq+1
          mbar_wait(mma_done, (gq - 1) & 1);
          tc_fence_after();
          uint32_t rq[32];
          tmem_ld_32x32(T_DQ + t_lane + hf * 32, rq);
          tmem_ld_wait();
          red_dq(rq, it - 1);
        }
        if (kind1 == 0) fb2_chunk<0, BF16>(cx, rs1, rd1, c1, q0);
        else if (kind1 == 1) fb2_chunk<1, BF16>(cx, rs1, rd1, c1, q0);
        else fb2_chunk<2, BF16>(cx, rs1, rd1, c1, q0);
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(pds_ready);
      }
      // ---- epilogue of this item, under the next item's loads and first S^T / dP^T ----
      uint32_t rv[32], rk[32];
      if (n_it > 0) {
        mbar_wait(mma_done, (g + n_it - 1) & 1);
        tc_fence_after();
        {
          uint32_t rq[32];
          tmem_ld_32x32(T_DQ + t_lane + hf * 32, rq);
          tmem_ld_wait();
          red_dq(rq, n_it - 1);
        }
        mbar_wait(dkv_full, I & 1);
        tc_fence_after();
        tmem_ld_32x32(T_DV + t_lane + hf * 32, rv);
        tmem_ld_32x32(T_DK + t_lane + hf * 32, rk);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(dkv_free);  // the next item's first dV / dK MMAs may overwrite the accumulators
      } else {
#pragma unroll
        for (int t = 0; t < 32; ++t) { rv[t] = 0u; rk[t] = 0u; }
      }
      if (!cx.key_oob) {
#pragma unroll
        for (int which = 0; which < 2; ++which) {
          void* basep = which == 0 ? bp.dv : bp.dk;
          const int64_t eo = which == 0
              ? (int64_t)b * bp.dv_sb + (int64_t)h * bp.dv_sh + (int64_t)jg * bp.dv_ss
              : (int64_t)b * bp.dk_sb + (int64_t)h * bp.dk_sh + (int64_t)jg * bp.dk_ss;
          uint8_t* row = reinterpret_cast<uint8_t*>(basep) + 2 * (eo + hf * 32);
#pragma unroll
          for (int gg = 0; gg < 4; ++gg) {
            uint4 w;
            float f[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(which == 0 ? rv[8 * gg + u] : rk[8 * gg + u]);
            if constexpr (BF16) {
              w.x = pack_bf16x2(f[0], f[1]); w.y = pack_bf16x2(f[2], f[3]);
              w.z = pack_bf16x2(f[4], f[5]); w.w = pack_bf16x2(f[6], f[7]);
            } else {
              __half2 x;
              x = __floats2half2_rn(f[0], f[1]); w.x = *reinterpret_cast<uint32_t*>(&x);
              x = __floats2half2_rn(f[2], f[3]); w.y = *reinterpret_cast<uint32_t*>(&x);
              x = __floats2half2_rn(f[4], f[5]); w.z = *reinterpret_cast<uint32_t*>(&x);
              x = __floats2half2_rn(f[6], f[7]); w.w = *reinterpret_cast<uint32_t*>(&x);
            }
            *reinterpret_cast<uint4*>(row + 16 * gg) = w;
          }
        }
      }
      if (n_it > 0) { g += n_it; ++I; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// v6: v3 with SIXTEEN compute warps. The r01f / r01g stamps put ~3 k of v3's ~4.7 k cycles per query tile into the
// element math of its eight compute warps (two per scheduler, each working through two 32-query chunks one after
// the other: MUFU and FMA latencies are barely hidden) and ~1 k into the dQ red.add burst. Here every compute warp
// owns ONE 32-query chunk (64 score / dP registers instead of 128), so four warps per scheduler interleave, and the
// dQ drain is spread over sixteen warps (16 head-dim columns each). 640 threads = 5 warpgroups: the TMA / MMA group
// gives its registers to the four compute groups with setmaxnreg (launch bound 96 -> 40 / 104).
// Written after round 1's GPU budget was spent: compiled, NOT yet run (ATTN_BWD_IMPL=7, opt-in tests only).
constexpr int FB6_THREADS = 640;

template <bool BF16>
__global__ void __launch_bounds__(FB6_THREADS, 1)
    attn_bwd_tc6_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                        const AttnBwdP bp) {
  const AttnP& p = bp.f;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sK = base, sV = base + FA_TILE;
  const uint32_t sQ = base + 2 * FA_TILE;   // 2 stages, each Q then dO
  constexpr int N_TILES = 14;
  constexpr uint32_t PDS_STRIDE = 4 * FA_TILE;  // buffer (it & 1) of the P^T / dS^T pair
  constexpr uint32_t STAT_STRIDE = 1024;
  constexpr int NCOMP = 512;                    // compute threads
  const uint32_t sPT = base + 6 * FA_TILE;  // 2 panels
  const uint32_t sDS = base + 8 * FA_TILE;  // 2 panels
  const uint32_t bars = base + N_TILES * FA_TILE;
  const uint32_t kv_full = bars, qdo_full = bars + 8, qdo_empty = bars + 24, sdp_full = bars + 40,
                 pds_ready = bars + 48, mma_done = bars + 56, dkv_full = bars + 64, tmem_slot = bars + 72,
                 sdp_free = bars + 80, lse_s = bars + 128, del_s = bars + 640;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem + N_TILES * FA_TILE + 72);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv_tiles = (p.Sk + 127) / 128;
  const int kv_tile = blockIdx.x % n_kv_tiles;
  const int bh = blockIdx.x / n_kv_tiles;
  const int h = bh % p.H, b = bh / p.H;
  const int kv0 = kv_tile * 128;
  const int n_q_tiles = (p.Sq + 127) / 128;
  int i_start = 0;
  if (p.causal) {
    const bool full_sweep = p.first_valid && (p.first_valid[b] > p.off);
    if (!full_sweep) i_start = max(0, (kv0 - p.off) / 128);
  }
  const int n_it = max(0, n_q_tiles - i_start);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmDO);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(qdo_full + 8 * s, 1); mbar_init(qdo_empty + 8 * s, 1); }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, NCOMP);
    mbar_init(pds_ready, NCOMP);
    mbar_init(mma_done, 1);
    mbar_init(dkv_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t T_ST = tmem, T_DPT = tmem + 128, T_DV = tmem + 256, T_DK = tmem + 320, T_DQ = tmem + 384;

  // setmaxnreg inside each role branch: ptxas sizes a region's registers by the setmaxnreg that dominates it
  const int wg = warp >> 2;
  if (wg == 0) {
   setmaxnreg_dec<40>();
   if (warp == 0) {
    if (lane == 0 && n_it > 0) {
      mbar_expect_tx(kv_full, 2 * FA_TILE);
      tma_load_4d(sK, &tmK, kv_full, 0, kv0, h, b);
      tma_load_4d(sV, &tmV, kv_full, 0, kv0, h, b);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1, q0 = (i_start + it) * 128;
        mbar_wait(qdo_empty + 8 * s, ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(qdo_full + 8 * s, 2 * FA_TILE);
        tma_load_4d(sQ + s * 2 * FA_TILE, &tmQ, qdo_full + 8 * s, 0, q0, h, b);
        tma_load_4d(sQ + s * 2 * FA_TILE + FA_TILE, &tmDO, qdo_full + 8 * s, 0, q0, h, b);
      }
    }
   } else if (warp == 1) {
    if (lane == 0 && n_it > 0) {
      const uint32_t idesc_kk = umma_idesc_f16(BF16 ? 1 : 0, 0, 0, 128, 128);  // S^T, dP^T
      const uint32_t idesc_km = umma_idesc_f16(BF16 ? 1 : 0, 0, 1, 128, 64);   // dV, dK
      const uint32_t idesc_mm = umma_idesc_f16(BF16 ? 1 : 0, 1, 1, 128, 64);   // dQ
      auto issue_sdp = [&](int it) {
        const int s = it & 1;
        const uint32_t q = sQ + s * 2 * FA_TILE, d_o = q + FA_TILE;
        mbar_wait(qdo_full + 8 * s, (it >> 1) & 1);
        if (it > 0) mbar_wait(sdp_free, (it - 1) & 1);  // every compute thread holds S^T/dP^T(it-1) in registers
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)  // S^T[kv, q] = K[kv, d] . Q[q, d]
          umma_f16(T_ST, umma_smem_desc_sw128(sK + k * 32, 0, 1024), umma_smem_desc_sw128(q + k * 32, 0, 1024),
                   idesc_kk, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dP^T[kv, q] = V[kv, d] . dO[q, d]
          umma_f16(T_DPT, umma_smem_desc_sw128(sV + k * 32, 0, 1024),
                   umma_smem_desc_sw128(d_o + k * 32, 0, 1024), idesc_kk, k > 0);
        umma_commit(sdp_full);
      };
      mbar_wait(kv_full, 0);
      issue_sdp(0);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1;
        const uint32_t q = sQ + s * 2 * FA_TILE, d_o = q + FA_TILE;
        if (it + 1 < n_it) issue_sdp(it + 1);  // runs under the element math of tile it
        mbar_wait(pds_ready, it & 1);
        tc_fence_after();
        const uint32_t pt = sPT + (it & 1) * PDS_STRIDE, dst = sDS + (it & 1) * PDS_STRIDE;
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dQ[q, d] = dS[q, kv] . K[kv, d]  (A MN-major view of dS^T)
          umma_f16(T_DQ, umma_smem_desc_sw128(dst + k * 2048, FA_TILE, 1024),
                   umma_smem_desc_sw128(sK + k * 2048, 64 * 128, 1024), idesc_mm, k > 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dV[kv, d] += P^T[kv, q] . dO[q, d]   (B MN-major: rows = q)
          umma_f16(T_DV, umma_smem_desc_sw128(pt + (k >> 2) * FA_TILE + (k & 3) * 32, 0, 1024),
                   umma_smem_desc_sw128(d_o + k * 2048, 64 * 128, 1024), idesc_km, (it > 0 || k > 0));
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dK[kv, d] += dS^T[kv, q] . Q[q, d]
          umma_f16(T_DK, umma_smem_desc_sw128(dst + (k >> 2) * FA_TILE + (k & 3) * 32, 0, 1024),
                   umma_smem_desc_sw128(q + k * 2048, 64 * 128, 1024), idesc_km, (it > 0 || k > 0));
        umma_commit(qdo_empty + 8 * s);
        umma_commit(mma_done);  // dQ(it) readable; P^T / dS^T buffer (it & 1) free for tile it+2
      }
      umma_commit(dkv_full);
    }
   }
  } else {
    setmaxnreg_inc<104>();  // 128 x 40 + 512 x 104 = 58 368 <= 640 x 96 (the CTA's pool at launch)
    const int wq = warp & 3;
    const int rr = wq * 32 + lane;   // key row inside the tile (S^T) / query row (dQ)
    const int c = wg - 1;            // the 32-query chunk this warp owns; also its 16 head-dim columns of dQ / dK / dV
    const int jg = kv0 + rr;
    const uint32_t t_lane = (uint32_t)(wq * 32) << 16;
    FbCtx cx;
    cx.kb = (p.kbias2 && jg < p.Sk) ? __ldg(p.kbias2 + (int64_t)b * p.kb_sb + (int64_t)h * p.kb_sh + jg) : 0.f;
    cx.sl2 = p.sl2; cx.scale = p.scale; cx.cf2 = p.causal_fill2;
    cx.jg = jg; cx.off = p.off; cx.causal = p.causal; cx.key_oob = jg >= p.Sk;
    cx.lse_s = lse_s; cx.del_s = del_s; cx.sPT = sPT; cx.sDS = sDS; cx.rr = rr; cx.sw = rr & 7;
    // generic arithmetic for the whole warp when a key is masked (the reference's finite fill matters on
    // fully masked query rows) or the key tile is ragged
    const bool warp_generic = __any_sync(0xffffffffu, cx.kb < -1e30f) || (kv0 + 128 > p.Sk);
    const bool fill_is_ninf = p.causal_fill2 == -INFINITY;
    const float* lse_bh = p.lse2 + ((int64_t)b * p.H + h) * p.Sq;
    const float* del_bh = bp.delta + ((int64_t)b * p.H + h) * p.Sq;
    // per-query statistics of the current query tile, staged as -lse2 and -delta*scale by the chunk-0 warps
    float nlse_next = -INFINITY, ndel_next = 0.f;
    if (c == 0 && n_it > 0 && i_start * 128 + rr < p.Sq) {
      nlse_next = -__ldg(lse_bh + i_start * 128 + rr);
      ndel_next = -__ldg(del_bh + i_start * 128 + rr) * p.scale;
    }
    // dQ of query tile `itp`: this thread owns query row rr, head-dim columns [16c, 16c + 16) = d/4 planes 4c..4c+3
    auto red_dq = [&](const uint32_t (&r)[16], int itp) {
      const int qi = (i_start + itp) * 128 + rr;
      if (qi < p.Sq) {
        float* dst = bp.dq_accum + (((int64_t)b * p.H + h) * n_q_tiles + (i_start + itp)) * FB_DQ_TILE +
                     (4 * c * 128 + rr) * 4;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 512 * g),
                       "f"(__uint_as_float(r[4 * g])), "f"(__uint_as_float(r[4 * g + 1])),
                       "f"(__uint_as_float(r[4 * g + 2])), "f"(__uint_as_float(r[4 * g + 3]))
                       : "memory");
      }
    };

    for (int it = 0; it < n_it; ++it) {
      const int q0 = (i_start + it) * 128;
      cx.sPT = sPT + (it & 1) * PDS_STRIDE; cx.sDS = sDS + (it & 1) * PDS_STRIDE;
      cx.lse_s = lse_s + (it & 1) * STAT_STRIDE; cx.del_s = del_s + (it & 1) * STAT_STRIDE;
      mbar_wait(sdp_full, it & 1);
      tc_fence_after();
      // chunk kind (warp-uniform)
      const bool touches_diag = p.causal && (kv0 + 127 > q0 + p.off);
      const bool aligned_diag = touches_diag && fill_is_ninf && (kv0 == q0 + p.off) && !warp_generic;
      int kind;
      if (warp_generic || (touches_diag && !aligned_diag)) kind = 2;
      else if (aligned_diag) kind = c < wq ? 1 : (c > wq ? 0 : 2);  // key row 32*wq+l vs queries 32*c..32*c+31
      else kind = 0;
      uint32_t rs[32], rd[32];
      if (kind != 1) { tmem_ld_32x32(T_ST + t_lane + c * 32, rs); tmem_ld_32x32(T_DPT + t_lane + c * 32, rd); }
      // statistics buffer (it & 1) was last read by tile it-2, and every thread finished tile it-2 before it
      // arrived at the named barrier of tile it-1, which this thread has passed
      if (c == 0) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(cx.lse_s + 4 * rr), "f"(nlse_next) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(cx.del_s + 4 * rr), "f"(ndel_next) : "memory");
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(sdp_free);        // S^T / dP^T are in registers: the next tile's MMAs may overwrite them
      bar_sync_named(1, NCOMP);     // statistics staged by the chunk-0 warps are visible
      if (c == 0) {
        const int nq = q0 + 128 + rr;
        const bool ok = (it + 1 < n_it) && nq < p.Sq;
        nlse_next = ok ? -__ldg(lse_bh + nq) : -INFINITY;
        ndel_next = ok ? -__ldg(del_bh + nq) * p.scale : 0.f;
      }
      if (kind == 0) fb2_chunk<0, BF16>(cx, rs, rd, c, q0);
      else if (kind == 1) fb2_chunk<1, BF16>(cx, rs, rd, c, q0);
      else fb2_chunk<2, BF16>(cx, rs, rd, c, q0);
      if (it > 0) {
        // The MMAs of tile it-1 ran under the chunk above: dQ(it-1) is complete (T_DQ is only rewritten after
        // pds_ready(it)). Waiting for every phase in order also proves that buffer ((it+1) & 1) of P^T / dS^T,
        // read by the MMAs of tile it-1, is free when tile it+1 writes it.
        mbar_wait(mma_done, (it - 1) & 1);
        tc_fence_after();
        uint32_t rq[16];
        tmem_ld_32x16(T_DQ + t_lane + c * 16, rq);
        tmem_ld_wait();
        red_dq(rq, it - 1);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_ready);
    }
    // ---- last dQ tile, then dK / dV for this key row: 16 of the 64 head-dim columns per thread ----
    if (n_it > 0) {
      mbar_wait(mma_done, (n_it - 1) & 1);
      tc_fence_after();
      uint32_t rq[16];
      tmem_ld_32x16(T_DQ + t_lane + c * 16, rq);
      tmem_ld_wait();
      red_dq(rq, n_it - 1);
      mbar_wait(dkv_full, 0);
      tc_fence_after();
    }
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      uint32_t r[16];
      if (n_it > 0) {
        tmem_ld_32x16((which == 0 ? T_DV : T_DK) + t_lane + c * 16, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int t = 0; t < 16; ++t) r[t] = 0u;
      }
      if (!cx.key_oob) {
        void* basep = which == 0 ? bp.dv : bp.dk;
        const int64_t eo = which == 0
            ? (int64_t)b * bp.dv_sb + (int64_t)h * bp.dv_sh + (int64_t)jg * bp.dv_ss
            : (int64_t)b * bp.dk_sb + (int64_t)h * bp.dk_sh + (int64_t)jg * bp.dk_ss;
        uint8_t* row = reinterpret_cast<uint8_t*>(basep) + 2 * (eo + c * 16);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint4 w;
          float f[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) f[u] = __uint_as_float(r[8 * g + u]);
          if constexpr (BF16) {
            w.x = pack_bf16x2(f[0], f[1]); w.y = pack_bf16x2(f[2], f[3]);
            w.z = pack_bf16x2(f[4], f[5]); w.w = pack_bf16x2(f[6], f[7]);
          } else {
            __half2 x;
            x = __floats2half2_rn(f[0], f[1]); w.x = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(f[2], f[3]); w.y = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(f[4], f[5]); w.z = *reinterpret_cast<uint32_t*>(&x);
            x = __floats2half2_rn(f[6], f[7]); w.w = *reinterpret_cast<uint32_t*>(&x);
          }
          *reinterpret_cast<uint4*>(row + 16 * g) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]   (one warp per (b,i,h); D <= 128)
__global__ void __launch_bounds__(256)
    attn_delta_kernel(const void* __restrict__ dout, const void* __restrict__ o, int fmt, int64_t sb,
                      int64_t sh, int64_t ss, float* __restrict__ delta, int B, int H, int Sq, int D) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (int64_t)B * H * Sq) return;
  const int i = (int)(w % Sq);
  const int h = (int)((w / Sq) % H);
  const int b = (int)(w / ((int64_t)Sq * H));
  const int64_t off = (int64_t)b * sb + (int64_t)h * sh + (int64_t)i * ss;
  float acc = 0.f;
  for (int d = lane; d < D; d += 32) {
    float x, y;
    if (fmt == 1) {
      x = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(dout)[off + d]);
      y = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(o)[off + d]);
    } else {
      x = __half2float(reinterpret_cast<const __half*>(dout)[off + d]);
      y = __half2float(reinterpret_cast<const __half*>(o)[off + d]);
    }
    acc = fmaf(x, y, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) delta[((int64_t)b * H + h) * Sq + i] = acc;
}

// D == 64, bf16, merged-head layout: 8 lanes x 16 bytes cover one (b,i,h) row; a warp handles 4 rows
__global__ void __launch_bounds__(256)
    attn_delta64_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ o, int64_t sb,
                        int64_t sh, int64_t ss, float* __restrict__ delta, int B, int H, int Sq) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = t >> 3;  // (b, i, h) flattened with h fastest: coalesced along the merged head dim
  const int part = (int)(t & 7);
  const bool ok = row < (int64_t)B * Sq * H;
  float acc = 0.f;
  int h = 0, i = 0, b = 0;
  if (ok) {
    h = (int)(row % H); i = (int)((row / H) % Sq); b = (int)(row / ((int64_t)H * Sq));
    const int64_t off = (int64_t)b * sb + (int64_t)h * sh + (int64_t)i * ss + part * 8;
    const uint4 x = *reinterpret_cast<const uint4*>(dout + off), y = *reinterpret_cast<const uint4*>(o + off);
    float2 a, c;
    a = unpack_bf16x2(x.x); c = unpack_bf16x2(y.x); acc = fmaf(a.x, c.x, fmaf(a.y, c.y, acc));
    a = unpack_bf16x2(x.y); c = unpack_bf16x2(y.y); acc = fmaf(a.x, c.x, fmaf(a.y, c.y, acc));
    a = unpack_bf16x2(x.z); c = unpack_bf16x2(y.z); acc = fmaf(a.x, c.x, fmaf(a.y, c.y, acc));
    a = unpack_bf16x2(x.w); c = unpack_bf16x2(y.w); acc = fmaf(a.x, c.x, fmaf(a.y, c.y, acc));
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if (ok && part == 0) delta[((int64_t)b * H + h) * Sq + i] = acc;
}

// dq[b,h,i,:] = (bf16) dq_accum[b,i,h,:]
__global__ void __launch_bounds__(256)
    attn_dq_convert_kernel(const float* __restrict__ acc, void* __restrict__ dq, int fmt, int64_t sb,
                           int64_t sh, int64_t ss, int B, int H, int Sq, int D) {
  // 8 elements per thread: two 16-byte loads, one 16-byte store (D % 8 == 0)
  const int64_t nvec = (int64_t)B * Sq * H * D / 8;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nvec; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t el = e * 8;
    const int d = (int)(el % D);
    const int h = (int)((el / D) % H);
    const int i = (int)((el / ((int64_t)D * H)) % Sq);
    const int b = (int)(el / ((int64_t)D * H * Sq));
    const float4 lo = __ldcs(reinterpret_cast<const float4*>(acc + el));
    const float4 hi = __ldcs(reinterpret_cast<const float4*>(acc + el) + 1);
    const int64_t off = (int64_t)b * sb + (int64_t)h * sh + (int64_t)i * ss + d;
    uint4 w;
    if (fmt == 1) {
      w.x = pack_bf16x2(lo.x, lo.y); w.y = pack_bf16x2(lo.z, lo.w);
      w.z = pack_bf16x2(hi.x, hi.y); w.w = pack_bf16x2(hi.z, hi.w);
    } else {
      __half2 t;
      t = __floats2half2_rn(lo.x, lo.y); w.x = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2half2_rn(lo.z, lo.w); w.y = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2half2_rn(hi.x, hi.y); w.z = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2half2_rn(hi.z, hi.w); w.w = *reinterpret_cast<uint32_t*>(&t);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(dq) + off) = w;
  }
}

// dq[b,h,i,:] = (bf16) of the TILED workspace [b][h][query tile][d/4][row][4] (D == 64). Eight threads cover one
// query row (coalesced 128-byte stores); their 16-byte loads pair up with the next row's in full 32-byte sectors.
__global__ void __launch_bounds__(256)
    attn_dq_convert_tiled_kernel(const float* __restrict__ acc, void* __restrict__ dq, int fmt, int64_t sb,
                                 int64_t sh, int64_t ss, int B, int H, int Sq) {
  const int nqt = (Sq + 127) / 128;
  const int64_t n = (int64_t)B * H * nqt * 1024;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int d8 = (int)(e & 7), row = (int)((e >> 3) & 127);
    const int64_t tile = e >> 10;
    const int qt = (int)(tile % nqt);
    const int64_t bh = tile / nqt;
    const int h = (int)(bh % H), b = (int)(bh / H);
    const int i = qt * 128 + row;
    if (i >= Sq) continue;
    const float* src = acc + tile * FB_DQ_TILE + ((2 * d8) * 128 + row) * 4;
    const float4 lo = __ldcs(reinterpret_cast<const float4*>(src));
    const float4 hi = __ldcs(reinterpret_cast<const float4*>(src + 128 * 4));
    const int64_t off = (int64_t)b * sb + (int64_t)h * sh + (int64_t)i * ss + d8 * 8;
    uint4 w;
    if (fmt == 1) {
      w.x = pack_bf16x2(lo.x, lo.y); w.y = pack_bf16x2(lo.z, lo.w);
      w.z = pack_bf16x2(hi.x, hi.y); w.w = pack_bf16x2(hi.z, hi.w);
    } else {
      __half2 t;
      t = __floats2half2_rn(lo.x, lo.y); w.x = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2half2_rn(lo.z, lo.w); w.y = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2half2_rn(hi.x, hi.y); w.z = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2half2_rn(hi.z, hi.w); w.w = *reinterpret_cast<uint32_t*>(&t);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(dq) + off) = w;
  }
}

// =================================================================================================
// SIMT kernels (any D <= 128): one warp per (b, h, query row) / per (b, h, key row)
// =================================================================================================
__device__ __forceinline__ float ld16(const void* p, int fmt, int64_t idx) {
  return fmt == 1 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx])
                  : __half2float(reinterpret_cast<const __half*>(p)[idx]);
}
__device__ __forceinline__ void st16(void* p, int fmt, int64_t idx, float v) {
  if (fmt == 1) reinterpret_cast<__nv_bfloat16*>(p)[idx] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(p)[idx] = __float2half_rn(v);
}

struct SimtP {
  AttnP a;
  int D;
  const void *q, *k, *v;
  int64_t q_sb, q_sh, q_ss, k_sb, k_sh, k_ss, v_sb, v_sh, v_ss;
};

constexpr int SIMT_WARPS = 4;

__global__ void __launch_bounds__(SIMT_WARPS * 32)
    attn_fwd_simt_kernel(const SimtP sp) {
  const AttnP& p = sp.a;
  __shared__ float qs[SIMT_WARPS][128];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * SIMT_WARPS + wib;
  if (w >= (int64_t)p.B * p.H * p.Sq) return;
  const int i = (int)(w % p.Sq);
  const int h = (int)((w / p.Sq) % p.H);
  const int b = (int)(w / ((int64_t)p.Sq * p.H));
  const int D = sp.D;
  const int64_t qoff = (int64_t)b * sp.q_sb + (int64_t)h * sp.q_sh + (int64_t)i * sp.q_ss;
  for (int d = lane; d < D; d += 32) qs[wib][d] = ld16(sp.q, p.fmt, qoff + d);
  __syncwarp();
  const float* kb_row = p.kbias2 ? p.kbias2 + (int64_t)b * p.kb_sb + (int64_t)h * p.kb_sh : nullptr;
  float m = -INFINITY, l = 0.f;
  float o_acc[4] = {0.f, 0.f, 0.f, 0.f};  // lane owns d = lane, lane+32, lane+64, lane+96
  for (int j0 = 0; j0 < p.Sk; j0 += 32) {
    const int j = j0 + lane;
    float v = -INFINITY;
    if (j < p.Sk) {
      const int64_t koff = (int64_t)b * sp.k_sb + (int64_t)h * sp.k_sh + (int64_t)j * sp.k_ss;
      float acc = 0.f;
      for (int d = 0; d < D; ++d) acc = fmaf(qs[wib][d], ld16(sp.k, p.fmt, koff + d), acc);
      v = score2(acc, p.sl2, kb_row ? kb_row[j] : 0.f, p.causal && (j > i + p.off), p.causal_fill2, false);
    }
    const float m_new = fmaxf(m, warp_max(v));
    const float alpha = ex2(m - m_new);
    const float e = ex2(v - m_new);
    // P is rounded to the activation dtype before the PV product, like the tensor-core path
    const float er = p.fmt == 1 ? __bfloat162float(__float2bfloat16_rn(e)) : __half2float(__float2half_rn(e));
    l = l * alpha + warp_sum(e);
    m = m_new;
#pragma unroll
    for (int u = 0; u < 4; ++u) o_acc[u] *= alpha;
    const int lim = min(32, p.Sk - j0);
    for (int t = 0; t < lim; ++t) {
      const float pj = __shfl_sync(0xffffffffu, er, t);
      const int64_t voff = (int64_t)b * sp.v_sb + (int64_t)h * sp.v_sh + (int64_t)(j0 + t) * sp.v_ss;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int d = lane + 32 * u;
        if (d < D) o_acc[u] = fmaf(pj, ld16(sp.v, p.fmt, voff + d), o_acc[u]);
      }
    }
  }
  const float inv = 1.f / l;
  const int64_t ooff = (int64_t)b * p.o_sb + (int64_t)h * p.o_sh + (int64_t)i * p.o_ss;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int d = lane + 32 * u;
    if (d < D) st16(p.o, p.fmt, ooff + d, o_acc[u] * inv);
  }
  if (p.lse2 && lane == 0) p.lse2[((int64_t)b * p.H + h) * p.Sq + i] = m + log2f(l);
}

struct SimtBwdP {
  SimtP s;
  const void* dout;
  const float* delta;
  void *dq, *dk, *dv;
  int64_t dq_sb, dq_sh, dq_ss, dk_sb, dk_sh, dk_ss, dv_sb, dv_sh, dv_ss;
};

// dq row: warp per (b,h,i)
__global__ void __launch_bounds__(SIMT_WARPS * 32)
    attn_bwd_dq_simt_kernel(const SimtBwdP bp) {
  const SimtP& sp = bp.s;
  const AttnP& p = sp.a;
  __shared__ float qs[SIMT_WARPS][128];
  __shared__ float dos[SIMT_WARPS][128];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * SIMT_WARPS + wib;
  if (w >= (int64_t)p.B * p.H * p.Sq) return;
  const int i = (int)(w % p.Sq);
  const int h = (int)((w / p.Sq) % p.H);
  const int b = (int)(w / ((int64_t)p.Sq * p.H));
  const int D = sp.D;
  const int64_t qoff = (int64_t)b * sp.q_sb + (int64_t)h * sp.q_sh + (int64_t)i * sp.q_ss;
  const int64_t ooff = (int64_t)b * p.o_sb + (int64_t)h * p.o_sh + (int64_t)i * p.o_ss;
  for (int d = lane; d < D; d += 32) {
    qs[wib][d] = ld16(sp.q, p.fmt, qoff + d);
    dos[wib][d] = ld16(bp.dout, p.fmt, ooff + d);
  }
  __syncwarp();
  const float* kb_row = p.kbias2 ? p.kbias2 + (int64_t)b * p.kb_sb + (int64_t)h * p.kb_sh : nullptr;
  const float lse = p.lse2[((int64_t)b * p.H + h) * p.Sq + i];
  const float dl = bp.delta[((int64_t)b * p.H + h) * p.Sq + i];
  float acc4[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j0 = 0; j0 < p.Sk; j0 += 32) {
    const int j = j0 + lane;
    float ds = 0.f;
    if (j < p.Sk) {
      const int64_t koff = (int64_t)b * sp.k_sb + (int64_t)h * sp.k_sh + (int64_t)j * sp.k_ss;
      const int64_t voff = (int64_t)b * sp.v_sb + (int64_t)h * sp.v_sh + (int64_t)j * sp.v_ss;
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < D; ++d) {
        s = fmaf(qs[wib][d], ld16(sp.k, p.fmt, koff + d), s);
        dp = fmaf(dos[wib][d], ld16(sp.v, p.fmt, voff + d), dp);
      }
      const bool fut = p.causal && (j > i + p.off);
      const float v = score2(s, p.sl2, kb_row ? kb_row[j] : 0.f, fut, p.causal_fill2, false);
      ds = fut ? 0.f : ex2(v - lse) * (dp - dl) * p.scale;
    }
    const int lim = min(32, p.Sk - j0);
    for (int t = 0; t < lim; ++t) {
      const float dsj = __shfl_sync(0xffffffffu, ds, t);
      const int64_t koff = (int64_t)b * sp.k_sb + (int64_t)h * sp.k_sh + (int64_t)(j0 + t) * sp.k_ss;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int d = lane + 32 * u;
        if (d < D) acc4[u] = fmaf(dsj, ld16(sp.k, p.fmt, koff + d), acc4[u]);
      }
    }
  }
  const int64_t dqoff = (int64_t)b * bp.dq_sb + (int64_t)h * bp.dq_sh + (int64_t)i * bp.dq_ss;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int d = lane + 32 * u;
    if (d < D) st16(bp.dq, p.fmt, dqoff + d, acc4[u]);
  }
}

// dk, dv rows: warp per (b,h,j)
__global__ void __launch_bounds__(SIMT_WARPS * 32)
    attn_bwd_dkv_simt_kernel(const SimtBwdP bp) {
  const SimtP& sp = bp.s;
  const AttnP& p = sp.a;
  __shared__ float ks[SIMT_WARPS][128];
  __shared__ float vs[SIMT_WARPS][128];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * SIMT_WARPS + wib;
  if (w >= (int64_t)p.B * p.H * p.Sk) return;
  const int j = (int)(w % p.Sk);
  const int h = (int)((w / p.Sk) % p.H);
  const int b = (int)(w / ((int64_t)p.Sk * p.H));
  const int D = sp.D;
  const int64_t koff = (int64_t)b * sp.k_sb + (int64_t)h * sp.k_sh + (int64_t)j * sp.k_ss;
  const int64_t voff = (int64_t)b * sp.v_sb + (int64_t)h * sp.v_sh + (int64_t)j * sp.v_ss;
  for (int d = lane; d < D; d += 32) {
    ks[wib][d] = ld16(sp.k, p.fmt, koff + d);
    vs[wib][d] = ld16(sp.v, p.fmt, voff + d);
  }
  __syncwarp();
  const float kb = p.kbias2 ? p.kbias2[(int64_t)b * p.kb_sb + (int64_t)h * p.kb_sh + j] : 0.f;
  float dk4[4] = {0.f, 0.f, 0.f, 0.f}, dv4[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i0 = 0; i0 < p.Sq; i0 += 32) {
    const int i = i0 + lane;
    float pe = 0.f, ds = 0.f;
    if (i < p.Sq) {
      const int64_t qoff = (int64_t)b * sp.q_sb + (int64_t)h * sp.q_sh + (int64_t)i * sp.q_ss;
      const int64_t ooff = (int64_t)b * p.o_sb + (int64_t)h * p.o_sh + (int64_t)i * p.o_ss;
      float s = 0.f, dp = 0.f;
      for (int d = 0; d < D; ++d) {
        s = fmaf(ld16(sp.q, p.fmt, qoff + d), ks[wib][d], s);
        dp = fmaf(ld16(bp.dout, p.fmt, ooff + d), vs[wib][d], dp);
      }
      const bool fut = p.causal && (j > i + p.off);
      const float v = score2(s, p.sl2, kb, fut, p.causal_fill2, false);
      pe = ex2(v - p.lse2[((int64_t)b * p.H + h) * p.Sq + i]);
      ds = fut ? 0.f : pe * (dp - bp.delta[((int64_t)b * p.H + h) * p.Sq + i]) * p.scale;
    }
    const int lim = min(32, p.Sq - i0);
    for (int t = 0; t < lim; ++t) {
      const float pi = __shfl_sync(0xffffffffu, pe, t);
      const float dsi = __shfl_sync(0xffffffffu, ds, t);
      const int64_t qoff = (int64_t)b * sp.q_sb + (int64_t)h * sp.q_sh + (int64_t)(i0 + t) * sp.q_ss;
      const int64_t ooff = (int64_t)b * p.o_sb + (int64_t)h * p.o_sh + (int64_t)(i0 + t) * p.o_ss;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int d = lane + 32 * u;
        if (d < D) {
          dv4[u] = fmaf(pi, ld16(bp.dout, p.fmt, ooff + d), dv4[u]);
          dk4[u] = fmaf(dsi, ld16(sp.q, p.fmt, qoff + d), dk4[u]);
        }
      }
    }
  }
  const int64_t dkoff = (int64_t)b * bp.dk_sb + (int64_t)h * bp.dk_sh + (int64_t)j * bp.dk_ss;
  const int64_t dvoff = (int64_t)b * bp.dv_sb + (int64_t)h * bp.dv_sh + (int64_t)j * bp.dv_ss;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int d = lane + 32 * u;
    if (d < D) {
      st16(bp.dk, p.fmt, dkoff + d, dk4[u]);
      st16(bp.dv, p.fmt, dvoff + d, dv4[u]);
    }
  }
}

// =================================================================================================
// mask preparation: one block per batch row
// =================================================================================================
__global__ void __launch_bounds__(256)
    attn_mask_prep_kernel(const void* __restrict__ mask, int mask_dtype, int Sk, int H, int mode,
                          const float* __restrict__ slopes, float* __restrict__ kbias2,
                          int32_t* __restrict__ first_valid) {
  extern __shared__ int pos_s[];  // [Sk] inclusive cumsum - 1
  __shared__ int first_s;
  const int b = blockIdx.x;
  auto mval = [&](int j) -> float {
    if (mask_dtype == 3) return (float)reinterpret_cast<const long long*>(mask)[(int64_t)b * Sk + j];
    if (mask_dtype == 4) return (float)reinterpret_cast<const int*>(mask)[(int64_t)b * Sk + j];
    return reinterpret_cast<const float*>(mask)[(int64_t)b * Sk + j];
  };
  {
    // inclusive scan of the mask row: each thread owns a contiguous segment, segment totals are scanned
    // with warp shuffles (256 threads = 8 warps), first valid key = block-wide min
    __shared__ int wsum_s[8];
    const int seg = (Sk + (int)blockDim.x - 1) / (int)blockDim.x;
    const int j0 = min(Sk, (int)threadIdx.x * seg), j1 = min(Sk, j0 + seg);
    int tot = 0, first = Sk;
    for (int j = j0; j < j1; ++j) {
      const int mv = (int)mval(j);
      tot += mv;
      if (mv != 0 && first == Sk) first = j;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    if (threadIdx.x == 0) first_s = Sk;
    __syncthreads();
    if (lane == 31) wsum_s[wid] = inc;
    if (lane == 0) atomicMin(&first_s, first);
    __syncthreads();
    int run = inc - tot;  // exclusive prefix inside the warp
    for (int w = 0; w < wid; ++w) run += wsum_s[w];
    for (int j = j0; j < j1; ++j) {
      run += (int)mval(j);
      pos_s[j] = run - 1;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && first_valid) first_valid[b] = first_s;
  const int heads = (mode == 0) ? H : 1;
  for (int e = threadIdx.x; e < heads * Sk; e += blockDim.x) {
    const int h = e / Sk, j = e % Sk;
    const float mv = mval(j);
    float add, pos = 0.f;
    if (mode == 0) {  // modeling_bloom.py:179-185 (fill finfo.min) + :328-331 (alibi)
      add = (mv != 0.f) ? 0.f : -FLT_MAX;
      pos = slopes[h] * ((float)pos_s[j] * mv);
    } else if (mode == 1) {  // modeling_gpt.py:176-179
      add = (1.0f - mv) * -FLT_MAX;
    } else {  // modeling_bert.py:303-304
      add = (1.0f - mv) * -10000.0f;
    }
    kbias2[((int64_t)b * heads + h) * Sk + j] = (pos + add) * LOG2E;
  }
}


// cache[b,h,pos+s,:] = src[b,h,s,:]  — in-place K/V append into a preallocated [B,H,T_max,D] cache
// (replaces the O(ctx) torch.concat of modeling_bloom.py:88-92 / modeling_gpt.py:76-80 per layer per step)
__global__ void __launch_bounds__(256)
    kv_append_kernel(const uint16_t* __restrict__ src, int64_t s_sb, int64_t s_sh, int64_t s_ss,
                     uint16_t* __restrict__ cache, int64_t c_sb, int64_t c_sh, int64_t c_ss, int B, int H,
                     int S, int D, int pos) {
  const int64_t n = (int64_t)B * H * S * D;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(e % D);
    const int sidx = (int)((e / D) % S);
    const int h = (int)((e / ((int64_t)D * S)) % H);
    const int b = (int)(e / ((int64_t)D * S * H));
    cache[(int64_t)b * c_sb + (int64_t)h * c_sh + (int64_t)(pos + sidx) * c_ss + d] =
        src[(int64_t)b * s_sb + (int64_t)h * s_sh + (int64_t)sidx * s_ss + d];
  }
}

static int make_qkv_tmap(CUtensorMap* tm, const void* base, int64_t sb, int64_t sh, int64_t ss, int B,
                         int H, int S, int D) {
  uint64_t dims[4] = {(uint64_t)D, (uint64_t)S, (uint64_t)H, (uint64_t)B};
  uint64_t strides[4] = {2, (uint64_t)ss * 2, (uint64_t)sh * 2, (uint64_t)sb * 2};
  uint32_t box[4] = {64, 128, 1, 1};
  return make_tmap(tm, base, 2, 4, dims, strides, box, 1);
}

static bool tma_ok4(const void* base, int64_t sb, int64_t sh, int64_t ss) {
  return (((uintptr_t)base & 15) == 0) && (sb % 8 == 0) && (sh % 8 == 0) && (ss % 8 == 0);
}

static void fill_common(AttnP& p, const ct_attn_args& a) {
  p.B = a.B; p.H = a.H; p.Sq = a.Sq; p.Sk = a.Sk;
  p.fmt = a.dtype == DT_BF16 ? 1 : 0;
  p.scale = a.scale;
  p.sl2 = a.scale * LOG2E;
  p.causal = a.causal;
  p.causal_fill2 = a.causal_fill * LOG2E;  // -FLT_MAX * log2e -> -inf; clamped in score2
  p.off = a.Sk - a.Sq;
  p.kbias2 = a.kbias2; p.kb_sb = a.kb_sb; p.kb_sh = a.kb_sh;
  p.first_valid = a.first_valid;
  p.o = a.o; p.o_sb = a.o_sb; p.o_sh = a.o_sh; p.o_ss = a.o_ss;
  p.lse2 = a.lse2;
}

static int check_args(const ct_attn_args& a, const char* who) {
  CT_REQUIRE(a.q && a.k && a.v && a.o, CT_ERR_BAD_ARG, "%s: null tensor", who);
  CT_REQUIRE(a.B > 0 && a.H > 0 && a.Sq > 0 && a.Sk > 0 && a.D > 0 && a.D <= 128, CT_ERR_BAD_ARG,
             "%s: bad shape B=%d H=%d Sq=%d Sk=%d D=%d", who, a.B, a.H, a.Sq, a.Sk, a.D);
  CT_REQUIRE(a.dtype == DT_BF16 || a.dtype == DT_F16, CT_ERR_UNSUPPORTED, "%s: dtype must be bf16/f16", who);
  return 0;
}

}  // namespace ct

using namespace ct;

extern "C" int ct_attn_fwd(const ct_attn_args* args, void* stream) {
  CT_REQUIRE(args != nullptr, CT_ERR_BAD_ARG, "ct_attn_fwd: null args");
  const ct_attn_args& a = *args;
  int rc = check_args(a, "ct_attn_fwd");
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc_ok = a.D == 64 && tma_ok4(a.q, a.q_sb, a.q_sh, a.q_ss) &&
                     tma_ok4(a.k, a.k_sb, a.k_sh, a.k_ss) && tma_ok4(a.v, a.v_sb, a.v_sh, a.v_ss) &&
                     tma_ok4(a.o, a.o_sb, a.o_sh, a.o_ss);
  bool use_tc;
  if (a.impl == 1) {
    CT_REQUIRE(tc_ok, CT_ERR_UNSUPPORTED, "ct_attn_fwd: tcgen05 path needs D=64 and 16-byte aligned strides");
    use_tc = true;
  } else if (a.impl == 2) {
    use_tc = false;
  } else {
    use_tc = tc_ok && a.Sq >= 16;
  }
  if (use_tc) {
    AttnP p;
    fill_common(p, a);
    CUtensorMap tmQ, tmK, tmV;
    if ((rc = make_qkv_tmap(&tmQ, a.q, a.q_sb, a.q_sh, a.q_ss, a.B, a.H, a.Sq, 64))) return rc;
    if ((rc = make_qkv_tmap(&tmK, a.k, a.k_sb, a.k_sh, a.k_ss, a.B, a.H, a.Sk, 64))) return rc;
    if ((rc = make_qkv_tmap(&tmV, a.v, a.v_sb, a.v_sh, a.v_ss, a.B, a.H, a.Sk, 64))) return rc;
    static bool attr = false;
    if (!attr) {
      CT_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc2_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc2_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
      attr = true;
    }
    const int64_t grid = (int64_t)a.B * a.H * ((a.Sq + 127) / 128);
    // ATTN_FWD_IMPL: 0 = auto (generation 4: two threads per query row), 4 = the same, 3 = generation 2;
    if ((option(OPT_ATTN_FWD_IMPL) == 0 || option(OPT_ATTN_FWD_IMPL) == 4) && a.scale > 0.f) {
      static bool attr4 = false;
      if (!attr4) {
        CT_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc4_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F4_SMEM));
        CT_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc4_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F4_SMEM));
        CT_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc4_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F4_SMEM));
        CT_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc4_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F4_SMEM));
        attr4 = true;
      }
      const bool kb = a.kbias2 != nullptr, bf = p.fmt == 1;
      if (kb && bf) attn_fwd_tc4_kernel<true, true><<<(unsigned)grid, F4_THREADS, F4_SMEM, st>>>(tmQ, tmK, tmV, p);
      else if (kb) attn_fwd_tc4_kernel<true, false><<<(unsigned)grid, F4_THREADS, F4_SMEM, st>>>(tmQ, tmK, tmV, p);
      else if (bf) attn_fwd_tc4_kernel<false, true><<<(unsigned)grid, F4_THREADS, F4_SMEM, st>>>(tmQ, tmK, tmV, p);
      else attn_fwd_tc4_kernel<false, false><<<(unsigned)grid, F4_THREADS, F4_SMEM, st>>>(tmQ, tmK, tmV, p);
      CT_LAUNCH_OK();
      return 0;
    }
    // ATTN_FWD_IMPL (older generations): 3 = v2 (register-resident score rows, O in TMEM), 1 = v1 (two TMEM passes),
    //                2 = v3 (v2 + lazy reference maximum + P handed over per 64-key panel; bf16 only; compiled,
    //                    not yet run on a GPU)
    const bool v2 = option(OPT_ATTN_FWD_IMPL) != 1 && a.scale > 0.f;
    if (v2 && option(OPT_ATTN_FWD_IMPL) == 2 && p.fmt == 1) {
      if (a.kbias2 != nullptr) attn_fwd_tc2_kernel<true, true, true><<<(unsigned)grid, FA_THREADS, FA_SMEM, st>>>(tmQ, tmK, tmV, p);
      else attn_fwd_tc2_kernel<false, true, true><<<(unsigned)grid, FA_THREADS, FA_SMEM, st>>>(tmQ, tmK, tmV, p);
    } else if (v2) {
      const bool kb = a.kbias2 != nullptr, bf = p.fmt == 1;
      if (kb && bf) attn_fwd_tc2_kernel<true, true><<<(unsigned)grid, FA_THREADS, FA_SMEM, st>>>(tmQ, tmK, tmV, p);
      else if (kb) attn_fwd_tc2_kernel<true, false><<<(unsigned)grid, FA_THREADS, FA_SMEM, st>>>(tmQ, tmK, tmV, p);
      else if (bf) attn_fwd_tc2_kernel<false, true><<<(unsigned)grid, FA_THREADS, FA_SMEM, st>>>(tmQ, tmK, tmV, p);
      else attn_fwd_tc2_kernel<false, false><<<(unsigned)grid, FA_THREADS, FA_SMEM, st>>>(tmQ, tmK, tmV, p);
    } else {
      attn_fwd_tc_kernel<<<(unsigned)grid, FA_THREADS, FA_SMEM, st>>>(tmQ, tmK, tmV, p);
    }
    CT_LAUNCH_OK();
    return 0;
  }
  SimtP sp;
  fill_common(sp.a, a);
  sp.D = a.D;
  sp.q = a.q; sp.k = a.k; sp.v = a.v;
  sp.q_sb = a.q_sb; sp.q_sh = a.q_sh; sp.q_ss = a.q_ss;
  sp.k_sb = a.k_sb; sp.k_sh = a.k_sh; sp.k_ss = a.k_ss;
  sp.v_sb = a.v_sb; sp.v_sh = a.v_sh; sp.v_ss = a.v_ss;
  const int64_t warps = (int64_t)a.B * a.H * a.Sq;
  attn_fwd_simt_kernel<<<(unsigned)((warps + SIMT_WARPS - 1) / SIMT_WARPS), SIMT_WARPS * 32, 0, st>>>(sp);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_attn_bwd(const ct_attn_bwd_args* args, void* stream) {
  CT_REQUIRE(args != nullptr, CT_ERR_BAD_ARG, "ct_attn_bwd: null args");
  const ct_attn_args& a = args->f;
  int rc = check_args(a, "ct_attn_bwd");
  if (rc) return rc;
  CT_REQUIRE(args->dout && args->dq && args->dk && args->dv && args->delta && a.lse2, CT_ERR_BAD_ARG,
             "ct_attn_bwd: null tensor");
  cudaStream_t st = (cudaStream_t)stream;
  const int fmt = a.dtype == DT_BF16 ? 1 : 0;
  {
    const int64_t warps = (int64_t)a.B * a.H * a.Sq;
    if (a.D == 64 && fmt == 1 && tma_ok4(args->dout, a.o_sb, a.o_sh, a.o_ss) && tma_ok4(a.o, a.o_sb, a.o_sh, a.o_ss))
      attn_delta64_kernel<<<(unsigned)((warps * 8 + 255) / 256), 256, 0, st>>>(
          (const __nv_bfloat16*)args->dout, (const __nv_bfloat16*)a.o, a.o_sb, a.o_sh, a.o_ss, args->delta, a.B,
          a.H, a.Sq);
    else
      attn_delta_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(
          args->dout, a.o, fmt, a.o_sb, a.o_sh, a.o_ss, args->delta, a.B, a.H, a.Sq, a.D);
    CT_LAUNCH_OK();
  }
  const bool tc_ok = a.D == 64 && tma_ok4(a.q, a.q_sb, a.q_sh, a.q_ss) &&
                     tma_ok4(a.k, a.k_sb, a.k_sh, a.k_ss) && tma_ok4(a.v, a.v_sb, a.v_sh, a.v_ss) &&
                     tma_ok4(args->dout, a.o_sb, a.o_sh, a.o_ss) &&
                     tma_ok4(args->dk, args->dk_sb, args->dk_sh, args->dk_ss) &&
                     tma_ok4(args->dv, args->dv_sb, args->dv_sh, args->dv_ss) && args->dq_accum != nullptr;
  bool use_tc;
  if (a.impl == 1) {
    CT_REQUIRE(tc_ok, CT_ERR_UNSUPPORTED, "ct_attn_bwd: tcgen05 path needs D=64, aligned strides, dq_accum");
    use_tc = true;
  } else if (a.impl == 2) {
    use_tc = false;
  } else {
    use_tc = tc_ok && a.Sq >= 16;
  }
  if (use_tc) {
    AttnBwdP bp;
    fill_common(bp.f, a);
    bp.delta = args->delta;
    bp.dq_accum = args->dq_accum;
    bp.dk = args->dk; bp.dk_sb = args->dk_sb; bp.dk_sh = args->dk_sh; bp.dk_ss = args->dk_ss;
    bp.dv = args->dv; bp.dv_sb = args->dv_sb; bp.dv_sh = args->dv_sh; bp.dv_ss = args->dv_ss;
    CUtensorMap tmQ, tmK, tmV, tmDO;
    if ((rc = make_qkv_tmap(&tmQ, a.q, a.q_sb, a.q_sh, a.q_ss, a.B, a.H, a.Sq, 64))) return rc;
    if ((rc = make_qkv_tmap(&tmK, a.k, a.k_sb, a.k_sh, a.k_ss, a.B, a.H, a.Sk, 64))) return rc;
    if ((rc = make_qkv_tmap(&tmV, a.v, a.v_sb, a.v_sh, a.v_ss, a.B, a.H, a.Sk, 64))) return rc;
    if ((rc = make_qkv_tmap(&tmDO, args->dout, a.o_sb, a.o_sh, a.o_ss, a.B, a.H, a.Sq, 64))) return rc;
    // ATTN_BWD_IMPL: 0 = auto (v3), 1 = v1, 2 = v2 (row-major dQ workspace), 3 = v2 + tiled dQ workspace,
    //                4 = v3 (tiled dQ workspace, P^T / dS^T double-buffered), 5 = v4 (v3 + dedicated dQ drain warpgroup),
    //                6 = v5 (persistent v3: the next work item's loads and first MMAs run under the epilogue; NOT yet
    //                    run on a GPU — written after the round's GPU budget was spent), 7 = v6 (v3 with sixteen
    //                    compute warps, one 32-query chunk each; same status)
    int variant = option(OPT_ATTN_BWD_IMPL);
    if (variant < 1 || variant > 8) variant = 8;  // 8 = v7: v3 + TMA reduce-add dQ drain + TMA-stored dK / dV
    const bool dq_tiled = variant >= 3 && variant != 8;
    const int nqt = (a.Sq + 127) / 128;
    // the workspace is sized for whole query tiles (include/ct_b200.h): B*H*ceil(Sq/128)*128*64 floats
    const size_t dq_elems = (size_t)a.B * a.H * nqt * FB_DQ_TILE;
    CT_CUDA_OK(cudaMemsetAsync(args->dq_accum, 0,
                               sizeof(float) * (dq_tiled ? dq_elems : (size_t)a.B * a.Sq * a.H * 64), st));
    static bool attr = false;
    if (!attr) {
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc2_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc2_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc2_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc2_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc2_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_PIPE));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc2_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_PIPE));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_PIPE));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_PIPE));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_PIPE));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_PIPE));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc6_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_PIPE));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc6_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_PIPE));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc7_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_PIPE));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc7_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_PIPE));
      attr = true;
    }
    const int64_t grid = (int64_t)a.B * a.H * ((a.Sk + 127) / 128);
    const unsigned g = (unsigned)grid;
    switch (variant) {
      case 8: {
        CUtensorMap tmDQ, tmDK, tmDV;
        // f32 workspace [B, Sq, H, 64] seen as (d, s, h, b); two 32-float (128-byte, swizzled) column panels per tile
        const uint64_t dims[4] = {64, (uint64_t)a.Sq, (uint64_t)a.H, (uint64_t)a.B};
        const uint64_t str[4] = {4, (uint64_t)a.H * 64 * 4, 64 * 4, (uint64_t)a.Sq * a.H * 64 * 4};
        const uint32_t box[4] = {32, 128, 1, 1};
        if ((rc = make_tmap(&tmDQ, args->dq_accum, 4, 4, dims, str, box, 1))) return rc;
        if ((rc = make_qkv_tmap(&tmDK, args->dk, args->dk_sb, args->dk_sh, args->dk_ss, a.B, a.H, a.Sk, 64))) return rc;
        if ((rc = make_qkv_tmap(&tmDV, args->dv, args->dv_sb, args->dv_sh, args->dv_ss, a.B, a.H, a.Sk, 64))) return rc;
        if (fmt == 1) attn_bwd_tc7_kernel<true><<<g, FB_THREADS, FB_SMEM_PIPE, st>>>(tmQ, tmK, tmV, tmDO, tmDQ, tmDK, tmDV, bp);
        else attn_bwd_tc7_kernel<false><<<g, FB_THREADS, FB_SMEM_PIPE, st>>>(tmQ, tmK, tmV, tmDO, tmDQ, tmDK, tmDV, bp);
        break;
      }
      case 1: attn_bwd_tc_kernel<<<g, FB_THREADS, FB_SMEM, st>>>(tmQ, tmK, tmV, tmDO, bp); break;
      case 2:
        if (fmt == 1) attn_bwd_tc2_kernel<true, 0><<<g, FB_THREADS, FB_SMEM, st>>>(tmQ, tmK, tmV, tmDO, bp);
        else attn_bwd_tc2_kernel<false, 0><<<g, FB_THREADS, FB_SMEM, st>>>(tmQ, tmK, tmV, tmDO, bp);
        break;
      case 3:
        if (fmt == 1) attn_bwd_tc2_kernel<true, 1><<<g, FB_THREADS, FB_SMEM, st>>>(tmQ, tmK, tmV, tmDO, bp);
        else attn_bwd_tc2_kernel<false, 1><<<g, FB_THREADS, FB_SMEM, st>>>(tmQ, tmK, tmV, tmDO, bp);
        break;
      case 7:
        if (fmt == 1) attn_bwd_tc6_kernel<true><<<g, FB6_THREADS, FB_SMEM_PIPE, st>>>(tmQ, tmK, tmV, tmDO, bp);
        else attn_bwd_tc6_kernel<false><<<g, FB6_THREADS, FB_SMEM_PIPE, st>>>(tmQ, tmK, tmV, tmDO, bp);
        break;
      case 6: {
        const int n_items = (int)grid;
        const unsigned pg = (unsigned)(n_items < sm_count() ? n_items : sm_count());
        if (fmt == 1) attn_bwd_tc5_kernel<true><<<pg, FB_THREADS, FB_SMEM_PIPE, st>>>(tmQ, tmK, tmV, tmDO, bp, n_items);
        else attn_bwd_tc5_kernel<false><<<pg, FB_THREADS, FB_SMEM_PIPE, st>>>(tmQ, tmK, tmV, tmDO, bp, n_items);
        break;
      }
      case 5:
        if (fmt == 1) attn_bwd_tc3_kernel<true><<<g, FB3_THREADS, FB_SMEM_PIPE, st>>>(tmQ, tmK, tmV, tmDO, bp);
        else attn_bwd_tc3_kernel<false><<<g, FB3_THREADS, FB_SMEM_PIPE, st>>>(tmQ, tmK, tmV, tmDO, bp);
        break;
      default:
        if (fmt == 1) attn_bwd_tc2_kernel<true, 3><<<g, FB_THREADS, FB_SMEM_PIPE, st>>>(tmQ, tmK, tmV, tmDO, bp);
        else attn_bwd_tc2_kernel<false, 3><<<g, FB_THREADS, FB_SMEM_PIPE, st>>>(tmQ, tmK, tmV, tmDO, bp);
        break;
    }
    CT_LAUNCH_OK();
    const int64_t n = dq_tiled ? (int64_t)(dq_elems / 8) : (int64_t)a.B * a.Sq * a.H * 64 / 8;
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
    if (dq_tiled)
      attn_dq_convert_tiled_kernel<<<(unsigned)blocks, 256, 0, st>>>(args->dq_accum, args->dq, fmt, args->dq_sb,
                                                                     args->dq_sh, args->dq_ss, a.B, a.H, a.Sq);
    else
      attn_dq_convert_kernel<<<(unsigned)blocks, 256, 0, st>>>(args->dq_accum, args->dq, fmt, args->dq_sb,
                                                               args->dq_sh, args->dq_ss, a.B, a.H, a.Sq, 64);
    CT_LAUNCH_OK();
    return 0;
  }
  SimtBwdP bp;
  fill_common(bp.s.a, a);
  bp.s.D = a.D;
  bp.s.q = a.q; bp.s.k = a.k; bp.s.v = a.v;
  bp.s.q_sb = a.q_sb; bp.s.q_sh = a.q_sh; bp.s.q_ss = a.q_ss;
  bp.s.k_sb = a.k_sb; bp.s.k_sh = a.k_sh; bp.s.k_ss = a.k_ss;
  bp.s.v_sb = a.v_sb; bp.s.v_sh = a.v_sh; bp.s.v_ss = a.v_ss;
  bp.dout = args->dout; bp.delta = args->delta;
  bp.dq = args->dq; bp.dk = args->dk; bp.dv = args->dv;
  bp.dq_sb = args->dq_sb; bp.dq_sh = args->dq_sh; bp.dq_ss = args->dq_ss;
  bp.dk_sb = args->dk_sb; bp.dk_sh = args->dk_sh; bp.dk_ss = args->dk_ss;
  bp.dv_sb = args->dv_sb; bp.dv_sh = args->dv_sh; bp.dv_ss = args->dv_ss;
  const int64_t wq = (int64_t)a.B * a.H * a.Sq, wk = (int64_t)a.B * a.H * a.Sk;
  attn_bwd_dq_simt_kernel<<<(unsigned)((wq + SIMT_WARPS - 1) / SIMT_WARPS), SIMT_WARPS * 32, 0, st>>>(bp);
  CT_LAUNCH_OK();
  attn_bwd_dkv_simt_kernel<<<(unsigned)((wk + SIMT_WARPS - 1) / SIMT_WARPS), SIMT_WARPS * 32, 0, st>>>(bp);
  CT_LAUNCH_OK();
  return 0;
}

// Diagnostic: resident CTAs per SM of the default tcgen05 forward / backward kernels (cudaOccupancy API, no launch).
// detail (optional, 8 ints): forward kernel numRegs, static shared bytes, dynamic shared bytes asked for, and its
// occupancy with the dynamic shared memory reduced by 0 / 1 / 2 / 4 / 16 KB (what limits it: registers or smem?).
extern "C" int ct_attn_occupancy(int* fwd_ctas_per_sm, int* bwd_ctas_per_sm, int* detail) {
  CT_REQUIRE(fwd_ctas_per_sm && bwd_ctas_per_sm, CT_ERR_BAD_ARG, "ct_attn_occupancy: null out");
  CT_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc4_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F4_SMEM));
  CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc7_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_PIPE));
  CT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(fwd_ctas_per_sm, attn_fwd_tc4_kernel<true, true>, F4_THREADS,
                                                           F4_SMEM));
  CT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(bwd_ctas_per_sm, attn_bwd_tc7_kernel<true>, FB_THREADS,
                                                           FB_SMEM_PIPE));
  if (detail) {
    cudaFuncAttributes fa;
    CT_CUDA_OK(cudaFuncGetAttributes(&fa, attn_fwd_tc4_kernel<true, true>));
    detail[0] = fa.numRegs; detail[1] = (int)fa.sharedSizeBytes; detail[2] = F4_SMEM;
    const int cut[5] = {0, 1024, 2048, 4096, 16384};
    for (int i = 0; i < 5; ++i)
      CT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(detail + 3 + i, attn_fwd_tc4_kernel<true, true>,
                                                               F4_THREADS, F4_SMEM - cut[i]));
  }
  return 0;
}

extern "C" int ct_attn_mask_prep(const void* attention_mask, int mask_dtype, int64_t B, int64_t Sk,
                                 int64_t H, int mode, const float* slopes, float* kbias2,
                                 int32_t* first_valid, void* stream) {
  CT_REQUIRE(attention_mask && kbias2, CT_ERR_BAD_ARG, "ct_attn_mask_prep: null pointer");
  CT_REQUIRE(mask_dtype == DT_F32 || mask_dtype == 3 || mask_dtype == 4, CT_ERR_UNSUPPORTED,
             "ct_attn_mask_prep: mask dtype must be f32/int64/int32");
  CT_REQUIRE(mode >= 0 && mode <= 2 && (mode != 0 || slopes), CT_ERR_BAD_ARG, "ct_attn_mask_prep: bad mode");
  CT_REQUIRE(B > 0 && Sk > 0 && Sk <= 12000 && H > 0, CT_ERR_BAD_ARG, "ct_attn_mask_prep: bad shape");
  attn_mask_prep_kernel<<<(unsigned)B, 256, sizeof(int) * Sk, (cudaStream_t)stream>>>(
      attention_mask, mask_dtype, (int)Sk, (int)H, mode, slopes, kbias2, first_valid);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_kv_append(const void* src, int64_t s_sb, int64_t s_sh, int64_t s_ss, void* cache,
                            int64_t c_sb, int64_t c_sh, int64_t c_ss, int B, int H, int S_new, int D,
                            int pos, int t_max, void* stream) {
  CT_REQUIRE(src && cache, CT_ERR_BAD_ARG, "ct_kv_append: null pointer");
  CT_REQUIRE(B > 0 && H > 0 && S_new >= 0 && D > 0 && pos >= 0 && pos + S_new <= t_max, CT_ERR_BAD_ARG,
             "ct_kv_append: rows [%d,%d) do not fit a cache of %d", pos, pos + S_new, t_max);
  if (S_new == 0) return 0;
  const int64_t n = (int64_t)B * H * S_new * D;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  kv_append_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const uint16_t*)src, s_sb, s_sh, s_ss, (uint16_t*)cache, c_sb, c_sh, c_ss, B, H, S_new, D, pos);
  CT_LAUNCH_OK();
  return 0;
}

// q_len = 1 (or a few) against a cache: same contract as ct_attn_fwd, warp-per-query-row kernel.
extern "C" int ct_attn_decode(const ct_attn_args* args, void* stream) {
  CT_REQUIRE(args != nullptr, CT_ERR_BAD_ARG, "ct_attn_decode: null args");
  ct_attn_args a = *args;
  a.impl = 2;
  a.lse2 = nullptr;
  return ct_attn_fwd(&a, stream);
}

#ifdef CT_DEBUG_TIMING
extern "C" int ct_debug_timing(long long* out, int n) {
  if (n > 4096) n = 4096;
  return (int)cudaMemcpyFromSymbol(out, ct::ct_dbg_clk, sizeof(long long) * n);
}
#endif
