// attention.cu — fused QK^T -> (+ALiBi / masks) -> online softmax -> PV, forward and backward.
//
// Replaces the materialised [B,H,Sq,Sk] score pipeline of
//   CleanTransformer/models/modeling_bloom.py:99-116, modeling_gpt.py:83-103, transformer.py:41-57
// (see include/ct_b200.h for the exact score definition shared by all three variants).
//
// tcgen05 forward  (head_dim 64): CTA = 128 query rows of one (b,h), 2 CTAs per SM; a TMA producer warp (Q once,
//   K/V double-buffered 128-key tiles), a single-thread MMA issuer (S = Q K^T into TMEM; O += P V accumulated in
//   TMEM), a bias-staging warp, and EIGHT softmax warps: two threads per query row (64 key columns each), the row
//   maximum crosses between them through a spare TMEM column, the running maximum is a lazily updated reference
//   (see attn_fwd_tc4_kernel).
// tcgen05 backward (head_dim 64): CTA = 128 keys of one (b,h), loops over query tiles:
//   S^T = K Q^T, dP^T = V dO^T (TMEM) -> P^T, dS^T (bf16, swizzled smem, double-buffered) -> dV += P^T dO,
//   dK += dS^T Q (TMEM accumulators), dQ_tile = dS K -> red.global.add.f32 into a tiled fp32 dQ workspace.
//   The Q/dO tiles are read through two descriptor views (K-major for the first pair of MMAs,
//   MN-major for the second), dS^T likewise (K-major for dK, MN-major for dQ): no transposes.
// Earlier generations (one thread per score row in the forward; register-resident / dedicated-drain / persistent /
// 16-warp / TMA-reduce backward variants) lost their A/B and were removed; the measurements are in profiles/
// (r01g_ab_attention.jsonl, r02a_*, r02c_*, r02h_*) and DESIGN.md §5 says what each one taught.
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
  const int32_t* sk_dev;  // decode from a CUDA graph: the number of cached keys lives in device memory (else null)
  DropKey drop;           // attention-probability dropout (thr == 0: off); element (b,h,i,j): hi = b*H + h, lo = i*Sk + j
  const uint16_t *k_new, *v_new;  // decode: the new token's rows, stored into cache row Sk - 1 by the kernel (else null)
  int64_t kn_sb, kn_sh, vn_sb, vn_sh;
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
constexpr int FA_TILE = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16, SWIZZLE_128B

// ---- helpers shared by the tcgen05 kernels ----
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
// DROP: probabilities are zeroed AFTER they were added to the row sum (drop_keep_pre over lo_term0 + e * 0x9E3779B1 for
// element e of the chunk); the 1 / (1 - p) factor is applied once, in the epilogue.
template <bool SCALED, bool BF16, bool DROP = false>
__device__ __forceinline__ void f4_exp_store(const uint32_t (&r)[32], int c, float m, float sl2, uint32_t p_row, int sw,
                                             float2& acc0, float2& acc1, uint32_t lo_term0 = 0u, uint32_t drop_pre = 0u,
                                             uint32_t drop_thr = 0u) {
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
    if constexpr (DROP) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t lt = lo_term0 + (uint32_t)(8 * g + 2 * u) * 0x9E3779B1u;
        if (!drop_keep_pre(lt, drop_pre, drop_thr)) v[u].x = 0.f;
        if (!drop_keep_pre(lt + 0x9E3779B1u, drop_pre, drop_thr)) v[u].y = 0.f;
      }
    }
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

template <bool HAS_KB, bool BF16, bool DROP>
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
  pdl_wait();               // programmatic dependent launch: see ct_common.cuh (first_valid is read right below)
  pdl_launch_dependents();
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
    const uint32_t drop_pre = (uint32_t)bh * 0x85EBCA77u + p.drop.key1;

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
      uint32_t lt0 = 0u;  // dropout: lo = i * Sk + column, as lo * 0x9E3779B1 + key0 for this thread's first column
      if constexpr (DROP) lt0 = ((uint32_t)i * (uint32_t)p.Sk + (uint32_t)col0) * 0x9E3779B1u + p.drop.key0;
      if (scaled) {
        f4_exp_store<true, BF16, DROP>(ra, 0, m_ref, p.sl2, p_row, sw, acc0, acc1, lt0, drop_pre, p.drop.thr);
        f4_exp_store<true, BF16, DROP>(rb, 1, m_ref, p.sl2, p_row, sw, acc0, acc1, lt0 + 32u * 0x9E3779B1u, drop_pre,
                                       p.drop.thr);
      } else {
        f4_exp_store<false, BF16, DROP>(ra, 0, m_ref, p.sl2, p_row, sw, acc0, acc1, lt0, drop_pre, p.drop.thr);
        f4_exp_store<false, BF16, DROP>(rb, 1, m_ref, p.sl2, p_row, sw, acc0, acc1, lt0 + 32u * 0x9E3779B1u, drop_pre,
                                        p.drop.thr);
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
      const float inv = (n_kv > 0) ? (DROP ? p.drop.rscale : 1.f) / l : 0.f;
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
constexpr int FB_THREADS_WIDE = 640;  // control warpgroup (TMA, MMA, 2 idle) + 16 compute warps (4 per lane quarter)
constexpr int FB_SMEM = 2 * FA_TILE /*K,V*/ + 4 * FA_TILE /*2 x (Q,dO)*/ + 2 * FA_TILE /*P^T*/ +
                        2 * FA_TILE /*dS^T*/ + 128 /*barriers*/ + 1024 /*lse2, delta of the query tile*/;
// v3 (pipelined): second P^T / dS^T buffer pair and a second statistics buffer
constexpr int FB_SMEM_PIPE = FB_SMEM + 4 * FA_TILE + 1024;
// dQ workspace, tiled layout: [b][h][query tile][d/4 (16)][row (128)][4] f32 — a warp's red.global.add.v4 then covers
// 512 contiguous bytes (32 rows x 16 B) instead of 32 different 128-byte lines
constexpr int FB_DQ_TILE = 128 * 64;

// =================================================================================================
// tcgen05 backward
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
  // attention-probability dropout (d_thr == 0: off): lo * 0x9E3779B1 + key0 = query * d_istep + d_jterm
  uint32_t d_jterm, d_istep, d_pre, d_thr;
  float d_rs;
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

// one 32-query chunk c of this thread's key row. KIND 0 = visible, 1 = entirely future, 2 = generic,
// 3 = generic with dropout: P^T (the dV operand) holds the surviving probabilities scaled by 1 / (1 - p), and dP reaches
// dS only through them: dS = P * (keep * dP / (1 - p) - delta)
// NG groups of 8 queries starting at group g0 of the chunk (NG = 4, g0 = 0: the whole chunk; the 16-warp kernel works
// through a chunk in two halves to halve the registers that hold S^T / dP^T)
template <int KIND, bool BF16, int NG = 4>
__device__ __forceinline__ void fb2_chunk(const FbCtx& cx, const uint32_t (&rs)[8 * NG], const uint32_t (&rd)[8 * NG], int c,
                                          int q0, int g0 = 0) {
  const float2 sl2v = make_float2(cx.sl2, cx.sl2), scv = make_float2(cx.scale, cx.scale), kbv = make_float2(cx.kb, cx.kb);
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float pt[8], ds[8];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int col = c * 32 + (g0 + g) * 8 + hh * 4;
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
      } else if constexpr (KIND == 3) {
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
          const float ks = drop_keep_pre((uint32_t)qg * cx.d_istep + cx.d_jterm, cx.d_pre, cx.d_thr) ? cx.d_rs : 0.f;
          pt[hh * 4 + u] = pe * ks;
          ds[hh * 4 + u] = fut ? 0.f : pe * fmaf(__uint_as_float(rd[e]) * ks, cx.scale, nds[u]);
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
    fb2_store_pair<BF16>(cx, c, g0 + g, pt, ds);
  }
}

// MODE bit 0: tiled dQ workspace; bit 1: P^T / dS^T (and the statistics) double-buffered, so the element math of
// query tile it+1 runs under the dQ / dV / dK MMAs of tile it (v2 waits for them before touching the tiles).
// MODE bit 2 (WIDE): 16 compute warps instead of 8 — four per TMEM lane quarter, one 32-query chunk each — so that
// every scheduler has four element-math warps to switch between instead of two (ncu on v3: warps active 15 %, issue
// active 23 %, stalls on dependent MUFU / FFMA2 chains). Warps 0-3 are the control warpgroup (TMA, MMA, two idle);
// a chunk is worked through in two halves of 16 queries, which keeps the kernel at 96 registers for 640 threads.
template <bool BF16, int MODE>
__global__ void __launch_bounds__((MODE & 4) ? FB_THREADS_WIDE : FB_THREADS, 1)
    attn_bwd_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDO,
                        const AttnBwdP bp) {
  const AttnP& p = bp.f;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  const uint32_t sK = base, sV = base + FA_TILE;
  const uint32_t sQ = base + 2 * FA_TILE;   // 2 stages, each Q then dO
  constexpr bool DQT = (MODE & 1) != 0, PIPE = (MODE & 2) != 0, WIDE = (MODE & 4) != 0;
  static_assert(!WIDE || (DQT && PIPE), "the 16-warp variant builds on the tiled dQ workspace and the double buffers");
  constexpr int N_COMPUTE = WIDE ? 512 : 256;  // compute threads
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
  pdl_wait();               // programmatic dependent launch: see ct_common.cuh (first_valid is read right below)
  pdl_launch_dependents();
  const int n_kv_tiles = (p.Sk + 127) / 128;
  // Under the causal mask key tile 0 meets every query tile and the last key tile only one: WIDE launches all heads'
  // heaviest tiles first (longest-processing-time order) so that the last wave holds one-tile CTAs, and the CTAs that add
  // into the same dQ tile are spread over the launch instead of running side by side.
  const int n_bh_ = p.B * p.H;
  const int kv_tile = WIDE ? (int)(blockIdx.x / n_bh_) : (int)(blockIdx.x % n_kv_tiles);
  const int bh = WIDE ? (int)(blockIdx.x % n_bh_) : (int)(blockIdx.x / n_kv_tiles);
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
    mbar_init(sdp_free, N_COMPUTE);
    mbar_init(pds_ready, N_COMPUTE);
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
  } else if (WIDE && warp < 4) {
    // idle members of the control warpgroup
  } else if constexpr (WIDE) {
    const int wq = warp & 3;
    const int rr = wq * 32 + lane;   // key row inside the tile (S^T) / query row (dQ)
    const int c = (warp - 4) >> 2;   // the 32-query chunk this warp owns; also its 16 of the 64 dQ / dK / dV columns
    const int jg = kv0 + rr;
    const uint32_t t_lane = (uint32_t)(wq * 32) << 16;
    FbCtx cx;
    cx.kb = (p.kbias2 && jg < p.Sk) ? __ldg(p.kbias2 + (int64_t)b * p.kb_sb + (int64_t)h * p.kb_sh + jg) : 0.f;
    cx.sl2 = p.sl2; cx.scale = p.scale; cx.cf2 = p.causal_fill2;
    cx.jg = jg; cx.off = p.off; cx.causal = p.causal; cx.key_oob = jg >= p.Sk;
    cx.lse_s = lse_s; cx.del_s = del_s; cx.sPT = sPT; cx.sDS = sDS; cx.rr = rr; cx.sw = rr & 7;
    cx.d_jterm = (uint32_t)jg * 0x9E3779B1u + p.drop.key0; cx.d_istep = (uint32_t)p.Sk * 0x9E3779B1u;
    cx.d_pre = (uint32_t)bh * 0x85EBCA77u + p.drop.key1; cx.d_thr = p.drop.thr; cx.d_rs = p.drop.rscale;
    const bool warp_generic = __any_sync(0xffffffffu, cx.kb < -1e30f) || (kv0 + 128 > p.Sk);
    const bool fill_is_ninf = p.causal_fill2 == -INFINITY;
    const float* lse_bh = p.lse2 + ((int64_t)b * p.H + h) * p.Sq;
    const float* del_bh = bp.delta + ((int64_t)b * p.H + h) * p.Sq;
    // Per-query statistics of this warp's chunk, staged as -lse2 and -delta*scale: lane l fetches query 32*c + l. The four
    // warps that share chunk c write identical values to the same words, so a __syncwarp is all a warp needs before it
    // reads them back (no CTA-wide barrier per tile); buffer (it & 1) is rewritten at tile it+2, which a warp reaches only
    // after every warp has arrived at sdp_free(it+1), i.e. has finished reading tile it's statistics.
    const int sq = 32 * c + lane;  // this lane's query inside the tile
    float nlse_next = -INFINITY, ndel_next = 0.f;
    if (n_it > 0 && i_start * 128 + sq < p.Sq) {
      nlse_next = -__ldg(lse_bh + i_start * 128 + sq);
      ndel_next = -__ldg(del_bh + i_start * 128 + sq) * p.scale;
    }
    // dQ of query tile `itp`: this thread owns query row (q0 + rr), columns [16*c, 16*c + 16) of d
    auto red_dq16 = [&](const uint32_t (&r)[16], int itp) {
      const int qi = (i_start + itp) * 128 + rr;
      if (qi < p.Sq) {
        float* dst = bp.dq_accum + (((int64_t)b * p.H + h) * n_q_tiles + (i_start + itp)) * FB_DQ_TILE +
                     (c * 4 * 128 + rr) * 4;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 128 * 4 * g),
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
      const bool touches_diag = p.causal && (kv0 + 127 > q0 + p.off);
      const bool aligned_diag = touches_diag && fill_is_ninf && (kv0 == q0 + p.off) && !warp_generic;
      int kind;
      if (warp_generic || (touches_diag && !aligned_diag)) kind = 2;
      else if (aligned_diag) kind = c < wq ? 1 : (c > wq ? 0 : 2);
      else kind = 0;
      // the chunk in two halves of 16 queries (32 live S^T / dP^T registers instead of 64)
      uint32_t rs[16], rd[16];
      if (kind != 1) { tmem_ld_32x16(T_ST + t_lane + c * 32, rs); tmem_ld_32x16(T_DPT + t_lane + c * 32, rd); }
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(cx.lse_s + 4 * sq), "f"(nlse_next) : "memory");
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(cx.del_s + 4 * sq), "f"(ndel_next) : "memory");
      tmem_ld_wait();
      __syncwarp();                       // this warp's 32 statistics are visible to its lanes
      {
        const int nq = q0 + 128 + sq;
        const bool ok = (it + 1 < n_it) && nq < p.Sq;
        nlse_next = ok ? -__ldg(lse_bh + nq) : -INFINITY;
        ndel_next = ok ? -__ldg(del_bh + nq) * p.scale : 0.f;
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (half == 1) {
          if (kind != 1) {
            tmem_ld_32x16(T_ST + t_lane + c * 32 + 16, rs);
            tmem_ld_32x16(T_DPT + t_lane + c * 32 + 16, rd);
          }
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(sdp_free);          // S^T / dP^T are in registers: the next tile's MMAs may overwrite them
        }
        if (cx.d_thr) fb2_chunk<3, BF16, 2>(cx, rs, rd, c, q0, 2 * half);
        else if (kind == 0) fb2_chunk<0, BF16, 2>(cx, rs, rd, c, q0, 2 * half);
        else if (kind == 1) fb2_chunk<1, BF16, 2>(cx, rs, rd, c, q0, 2 * half);
        else fb2_chunk<2, BF16, 2>(cx, rs, rd, c, q0, 2 * half);
      }
      fence_proxy_async_smem();
      if (it > 0) {
        // dQ(it-1) must leave TMEM before pds_ready(it) lets the MMA warp overwrite it; its MMAs ran under this tile's
        // element math. Waiting for every mma_done phase in order also proves that buffer ((it+1) & 1) of P^T / dS^T —
        // read by the MMAs of tile it-1 — is free when tile it+1 writes it.
        mbar_wait(mma_done, (it - 1) & 1);
        tc_fence_after();
        uint32_t rq[16];
        tmem_ld_32x16(T_DQ + t_lane + c * 16, rq);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(pds_ready);
        red_dq16(rq, it - 1);             // the atomics drain while the next tile starts
      } else {
        tc_fence_before();
        mbar_arrive(pds_ready);
      }
    }
    if (n_it > 0) {
      mbar_wait(mma_done, (n_it - 1) & 1);
      tc_fence_after();
      uint32_t rq[16];
      tmem_ld_32x16(T_DQ + t_lane + c * 16, rq);
      tmem_ld_wait();
      red_dq16(rq, n_it - 1);
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
    cx.d_jterm = (uint32_t)jg * 0x9E3779B1u + p.drop.key0; cx.d_istep = (uint32_t)p.Sk * 0x9E3779B1u;
    cx.d_pre = (uint32_t)bh * 0x85EBCA77u + p.drop.key1; cx.d_thr = p.drop.thr; cx.d_rs = p.drop.rscale;
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
      if (cx.d_thr) fb2_chunk<3, BF16>(cx, rs0, rd0, c0, q0);
      else if (kind0 == 0) fb2_chunk<0, BF16>(cx, rs0, rd0, c0, q0);
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
      if (cx.d_thr) fb2_chunk<3, BF16>(cx, rs1, rd1, c1, q0);
      else if (kind1 == 0) fb2_chunk<0, BF16>(cx, rs1, rd1, c1, q0);
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
  pdl_wait();               // programmatic dependent launch: see ct_common.cuh
  pdl_launch_dependents();
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

// dq[b,h,i,:] = (bf16) of the TILED workspace [b][h][query tile][d/4][row][4] (D == 64). Eight threads cover one
// query row (coalesced 128-byte stores); their 16-byte loads pair up with the next row's in full 32-byte sectors.
__global__ void __launch_bounds__(256)
    attn_dq_convert_tiled_kernel(const float* __restrict__ acc, void* __restrict__ dq, int fmt, int64_t sb,
                                 int64_t sh, int64_t ss, int B, int H, int Sq) {
  pdl_wait();               // programmatic dependent launch: see ct_common.cuh
  pdl_launch_dependents();
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
    float er = p.fmt == 1 ? __bfloat162float(__float2bfloat16_rn(e)) : __half2float(__float2half_rn(e));
    if (p.drop.thr && !drop_keep(p.drop, (uint32_t)(b * p.H + h), (uint32_t)i * (uint32_t)p.Sk + (uint32_t)j))
      er = 0.f;  // dropped after the softmax: the row sum below still counts the key
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
  const float inv = p.drop.rscale / l;  // survivors of the dropout are scaled by 1 / (1 - p)
  const int64_t ooff = (int64_t)b * p.o_sb + (int64_t)h * p.o_sh + (int64_t)i * p.o_ss;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int d = lane + 32 * u;
    if (d < D) st16(p.o, p.fmt, ooff + d, o_acc[u] * inv);
  }
  if (p.lse2 && lane == 0) p.lse2[((int64_t)b * p.H + h) * p.Sq + i] = m + log2f(l);
}

// q_len = 1 decode against a [b,h,t,d] cache (modeling_gpt.py:76-103 / modeling_bloom.py:88-116 with one new token),
// head_dim D in {32, 64, 128}: one CTA of DEC_WARPS warps per (b, h), the keys dealt to the warps in groups of
// 4 * KPI (KPI = keys per warp instruction). Keys and values are read as whole rows (D/8 lanes x 16 bytes per key); the
// key count may live in device memory (sk_dev) so that one captured step serves every position of a generation.
// Same score definition as every other path (score2); P is rounded to the activation dtype before P.V.
constexpr int DEC_WARPS = 8;

template <int D>
__global__ void __launch_bounds__(DEC_WARPS * 32, 4)  // 4 CTAs / SM: 592 (b, h) pairs in one wave
    attn_decode_kernel(const SimtP sp) {
  constexpr int LPK = D / 8;     // lanes per key row
  constexpr int KPI = 32 / LPK;  // keys per warp-wide load instruction
  constexpr int G = 4;           // load groups in flight per iteration
  const AttnP& p = sp.a;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = (int)(blockIdx.x % p.H), b = (int)(blockIdx.x / p.H);
  pdl_wait();               // programmatic dependent launch: see ct_common.cuh
  pdl_launch_dependents();
  const int Sk = p.sk_dev ? min(*p.sk_dev, p.Sk) : p.Sk;  // p.Sk is the capacity when the count is on the device
  const int sub = lane / LPK, part = lane % LPK;  // key sub-slot inside a group, 16-byte piece of the row
  if (p.k_new) {
    // the reference's torch.concat((past, new)): this CTA owns cache rows (b, h, :), so it stores the new token's key and
    // value at row Sk - 1 itself and reads them back after the barrier like any other cached row
    if (Sk >= 1 && threadIdx.x < 2 * LPK) {
      const int which = threadIdx.x / LPK, pc = threadIdx.x % LPK;
      const uint16_t* src = which == 0 ? p.k_new + (int64_t)b * p.kn_sb + (int64_t)h * p.kn_sh
                                       : p.v_new + (int64_t)b * p.vn_sb + (int64_t)h * p.vn_sh;
      uint16_t* dst = which == 0
          ? const_cast<uint16_t*>(reinterpret_cast<const uint16_t*>(sp.k)) + (int64_t)b * sp.k_sb + (int64_t)h * sp.k_sh + (int64_t)(Sk - 1) * sp.k_ss
          : const_cast<uint16_t*>(reinterpret_cast<const uint16_t*>(sp.v)) + (int64_t)b * sp.v_sb + (int64_t)h * sp.v_sh + (int64_t)(Sk - 1) * sp.v_ss;
      *reinterpret_cast<uint4*>(dst + pc * 8) = *reinterpret_cast<const uint4*>(src + pc * 8);
    }
    __syncthreads();
  }
  auto unpack2 = [&](uint32_t w) {
    return p.fmt == 1 ? unpack_bf16x2(w) : __half22float2(*reinterpret_cast<const __half2*>(&w));
  };
  float qf[8];  // the 8 query elements this lane multiplies
  {
    const uint4 qv = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(sp.q) + (int64_t)b * sp.q_sb +
                                                     (int64_t)h * sp.q_sh + part * 8);
    const uint32_t qw[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 f = unpack2(qw[u]);
      qf[2 * u] = f.x; qf[2 * u + 1] = f.y;
    }
  }
  const float* kb_row = p.kbias2 ? p.kbias2 + (int64_t)b * p.kb_sb + (int64_t)h * p.kb_sh : nullptr;
  const uint16_t* kbase = reinterpret_cast<const uint16_t*>(sp.k) + (int64_t)b * sp.k_sb + (int64_t)h * sp.k_sh + part * 8;
  const uint16_t* vbase = reinterpret_cast<const uint16_t*>(sp.v) + (int64_t)b * sp.v_sb + (int64_t)h * sp.v_sh + part * 8;
  float m = -INFINITY, l = 0.f;
  float o_acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // this lane's 8 output columns, for its key sub-slot
  for (int j0 = wib * (G * KPI); j0 < Sk; j0 += DEC_WARPS * G * KPI) {
    float sc[G];
    uint4 kv[G], vv[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {  // all loads of the iteration first: 2 * G 16-byte requests in flight per lane
      const int j = j0 + KPI * g + sub;
      kv[g] = make_uint4(0u, 0u, 0u, 0u);
      vv[g] = make_uint4(0u, 0u, 0u, 0u);
      if (j < Sk) {
        kv[g] = *reinterpret_cast<const uint4*>(kbase + (int64_t)j * sp.k_ss);
        vv[g] = *reinterpret_cast<const uint4*>(vbase + (int64_t)j * sp.v_ss);
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int j = j0 + KPI * g + sub;
      const uint32_t kw[4] = {kv[g].x, kv[g].y, kv[g].z, kv[g].w};
      float acc = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 f = unpack2(kw[u]);
        acc = fmaf(qf[2 * u], f.x, acc);
        acc = fmaf(qf[2 * u + 1], f.y, acc);
      }
#pragma unroll
      for (int o = LPK / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      // q_len == 1: the query is the last position, nothing lies in its future
      sc[g] = j < Sk ? score2(acc, p.sl2, kb_row ? __ldg(kb_row + j) : 0.f, false, p.causal_fill2, false) : -INFINITY;
    }
    float mt = sc[0];
#pragma unroll
    for (int g = 1; g < G; ++g) mt = fmaxf(mt, sc[g]);
#pragma unroll
    for (int o = LPK; o < 32; o <<= 1) mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
    const float m_new = fmaxf(m, mt);
    const float alpha = ex2(m - m_new);
    l *= alpha;
#pragma unroll
    for (int u = 0; u < 8; ++u) o_acc[u] *= alpha;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float e = ex2(sc[g] - m_new);
      const float er = p.fmt == 1 ? __bfloat162float(__float2bfloat16_rn(e)) : __half2float(__float2half_rn(e));
      if (part == 0) l += e;  // one lane per key sub-slot counts the key once
      const uint32_t vw[4] = {vv[g].x, vv[g].y, vv[g].z, vv[g].w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 f = unpack2(vw[u]);
        o_acc[2 * u] = fmaf(er, f.x, o_acc[2 * u]);
        o_acc[2 * u + 1] = fmaf(er, f.y, o_acc[2 * u + 1]);
      }
    }
    m = m_new;
  }
  // combine the key sub-slots of the warp (lanes with the same `part` share m already) ...
#pragma unroll
  for (int o = LPK; o < 32; o <<= 1) {
    l += __shfl_xor_sync(0xffffffffu, l, o);
#pragma unroll
    for (int u = 0; u < 8; ++u) o_acc[u] += __shfl_xor_sync(0xffffffffu, o_acc[u], o);
  }
  l = __shfl_sync(0xffffffffu, l, 0);  // lanes with part != 0 carried partial counts only
  // ... then the warps of the CTA
  __shared__ float s_m[DEC_WARPS], s_l[DEC_WARPS];
  __shared__ float s_o[DEC_WARPS][D];
  if (lane == 0) { s_m[wib] = m; s_l[wib] = l; }
  if (sub == 0) {
#pragma unroll
    for (int u = 0; u < 8; ++u) s_o[wib][part * 8 + u] = o_acc[u];
  }
  __syncthreads();
  if (threadIdx.x < D) {
    float mm = s_m[0];
#pragma unroll
    for (int w = 1; w < DEC_WARPS; ++w) mm = fmaxf(mm, s_m[w]);
    float lt = 0.f, ot = 0.f;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) {
      const float f = s_m[w] == -INFINITY ? 0.f : ex2(s_m[w] - mm);  // warps that saw no key contribute nothing
      lt = fmaf(s_l[w], f, lt);
      ot = fmaf(s_o[w][threadIdx.x], f, ot);
    }
    st16(p.o, p.fmt, (int64_t)b * p.o_sb + (int64_t)h * p.o_sh + threadIdx.x, ot / lt);
  }
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
      if (p.drop.thr)  // dP reaches the scores only through the surviving probabilities, scaled by 1 / (1 - p)
        dp = drop_keep(p.drop, (uint32_t)(b * p.H + h), (uint32_t)i * (uint32_t)p.Sk + (uint32_t)j) ? dp * p.drop.rscale : 0.f;
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
      float keep_scale = 1.f;
      if (p.drop.thr)
        keep_scale = drop_keep(p.drop, (uint32_t)(b * p.H + h), (uint32_t)i * (uint32_t)p.Sk + (uint32_t)j) ? p.drop.rscale : 0.f;
      ds = fut ? 0.f : pe * (dp * keep_scale - bp.delta[((int64_t)b * p.H + h) * p.Sq + i]) * p.scale;
      pe *= keep_scale;  // dV sees the dropped probabilities
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

// the same with the destination taken from device memory: rows [*len_dev - S, *len_dev) (captured decode step)
__global__ void __launch_bounds__(256)
    kv_append_dev_kernel(const uint16_t* __restrict__ src, int64_t s_sb, int64_t s_sh, int64_t s_ss,
                         uint16_t* __restrict__ cache, int64_t c_sb, int64_t c_sh, int64_t c_ss, int B, int H,
                         int S, int D, const int32_t* __restrict__ len_dev, int t_max) {
  const int pos = *len_dev - S;
  if (pos < 0 || pos + S > t_max) return;  // (a full cache: nothing is written; the host sized it for the generation)
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
  p.sk_dev = a.seq_len_dev;
  p.drop = make_drop_key(a.dropout_p, a.rng_seed, a.rng_stream);
  p.k_new = (const uint16_t*)a.k_new; p.v_new = (const uint16_t*)a.v_new;
  p.kn_sb = a.kn_sb; p.kn_sh = a.kn_sh; p.vn_sb = a.vn_sb; p.vn_sh = a.vn_sh;
}

static int check_args(const ct_attn_args& a, const char* who) {
  CT_REQUIRE(a.q && a.k && a.v && a.o, CT_ERR_BAD_ARG, "%s: null tensor", who);
  CT_REQUIRE(a.B > 0 && a.H > 0 && a.Sq > 0 && a.Sk > 0 && a.D > 0 && a.D <= 128, CT_ERR_BAD_ARG,
             "%s: bad shape B=%d H=%d Sq=%d Sk=%d D=%d", who, a.B, a.H, a.Sq, a.Sk, a.D);
  CT_REQUIRE(a.dtype == DT_BF16 || a.dtype == DT_F16, CT_ERR_UNSUPPORTED, "%s: dtype must be bf16/f16", who);
  CT_REQUIRE(a.dropout_p >= 0.f && a.dropout_p < 1.f, CT_ERR_BAD_ARG, "%s: dropout_p must be in [0, 1)", who);
  CT_REQUIRE(a.dropout_p == 0.f || a.seq_len_dev == nullptr, CT_ERR_UNSUPPORTED, "%s: no dropout in the decode step", who);
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
    use_tc = tc_ok && a.Sq >= 16 && a.scale > 0.f;
  }
  if (a.seq_len_dev != nullptr) use_tc = false;
  if (use_tc) {
    AttnP p;
    fill_common(p, a);
    CUtensorMap tmQ, tmK, tmV;
    if ((rc = make_qkv_tmap(&tmQ, a.q, a.q_sb, a.q_sh, a.q_ss, a.B, a.H, a.Sq, 64))) return rc;
    if ((rc = make_qkv_tmap(&tmK, a.k, a.k_sb, a.k_sh, a.k_ss, a.B, a.H, a.Sk, 64))) return rc;
    if ((rc = make_qkv_tmap(&tmV, a.v, a.v_sb, a.v_sh, a.v_ss, a.B, a.H, a.Sk, 64))) return rc;
    CT_REQUIRE(a.scale > 0.f, CT_ERR_BAD_ARG, "ct_attn_fwd: the tcgen05 path needs scale > 0");
    static bool attr4 = false;
    if (!attr4) {
#define CT_F4_ATTR(KB, BF, DR) \
  CT_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc4_kernel<KB, BF, DR>, cudaFuncAttributeMaxDynamicSharedMemorySize, F4_SMEM))
      CT_F4_ATTR(true, true, false); CT_F4_ATTR(true, false, false); CT_F4_ATTR(false, true, false); CT_F4_ATTR(false, false, false);
      CT_F4_ATTR(true, true, true); CT_F4_ATTR(true, false, true); CT_F4_ATTR(false, true, true); CT_F4_ATTR(false, false, true);
#undef CT_F4_ATTR
      attr4 = true;
    }
    const int64_t grid = (int64_t)a.B * a.H * ((a.Sq + 127) / 128);
    const bool kb = a.kbias2 != nullptr, bf = p.fmt == 1, dr = p.drop.thr != 0u;
#define CT_F4_GO(KB, BF, DR) \
  CT_CUDA_OK(launch_k(attn_fwd_tc4_kernel<KB, BF, DR>, dim3((unsigned)grid), dim3(F4_THREADS), F4_SMEM, st, tmQ, tmK, tmV, p))
    if (dr) {
      if (kb && bf) CT_F4_GO(true, true, true);
      else if (kb) CT_F4_GO(true, false, true);
      else if (bf) CT_F4_GO(false, true, true);
      else CT_F4_GO(false, false, true);
    } else {
      if (kb && bf) CT_F4_GO(true, true, false);
      else if (kb) CT_F4_GO(true, false, false);
      else if (bf) CT_F4_GO(false, true, false);
      else CT_F4_GO(false, false, false);
    }
#undef CT_F4_GO
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
  const bool decode = a.Sq == 1 && a.dropout_p == 0.f && (a.D == 32 || a.D == 64 || a.D == 128) && a.lse2 == nullptr &&
                      tma_ok4(a.q, a.q_sb, a.q_sh, 8) && tma_ok4(a.k, a.k_sb, a.k_sh, a.k_ss) &&
                      tma_ok4(a.v, a.v_sb, a.v_sh, a.v_ss);
  CT_REQUIRE((a.k_new == nullptr) == (a.v_new == nullptr), CT_ERR_BAD_ARG, "ct_attn_fwd: k_new and v_new come together");
  CT_REQUIRE(a.k_new == nullptr || (decode && (((uintptr_t)a.k_new | (uintptr_t)a.v_new) & 15) == 0 &&
                                    ((a.kn_sb | a.kn_sh | a.vn_sb | a.vn_sh) & 7) == 0),
             CT_ERR_UNSUPPORTED, "ct_attn_fwd: k_new / v_new need the q_len = 1 decode kernel and 16-byte aligned rows");
  CT_REQUIRE(a.seq_len_dev == nullptr || decode, CT_ERR_UNSUPPORTED,
             "ct_attn_fwd: a device-side key count needs the q_len = 1 decode kernel (head_dim 32 / 64 / 128, 16-byte "
             "aligned rows, no lse output)");
  if (decode) {
    const unsigned grid = (unsigned)((int64_t)a.B * a.H);
    if (a.D == 32) CT_CUDA_OK(launch_k(attn_decode_kernel<32>, dim3(grid), dim3(DEC_WARPS * 32), 0, st, sp));
    else if (a.D == 64) CT_CUDA_OK(launch_k(attn_decode_kernel<64>, dim3(grid), dim3(DEC_WARPS * 32), 0, st, sp));
    else CT_CUDA_OK(launch_k(attn_decode_kernel<128>, dim3(grid), dim3(DEC_WARPS * 32), 0, st, sp));
    CT_LAUNCH_OK();
    return 0;
  }
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
  // delta = rowsum(dO * O): launched right before its consumer (after the workspace memset on the tcgen05 path, so that
  // delta -> backward -> convert is an unbroken kernel chain for programmatic dependent launch)
  auto launch_delta = [&]() -> int {
    const int64_t warps = (int64_t)a.B * a.H * a.Sq;
    if (a.D == 64 && fmt == 1 && tma_ok4(args->dout, a.o_sb, a.o_sh, a.o_ss) && tma_ok4(a.o, a.o_sb, a.o_sh, a.o_ss))
      CT_CUDA_OK(launch_k(attn_delta64_kernel, dim3((unsigned)((warps * 8 + 255) / 256)), dim3(256), 0, st,
                          (const __nv_bfloat16*)args->dout, (const __nv_bfloat16*)a.o, a.o_sb, a.o_sh, a.o_ss,
                          args->delta, a.B, a.H, a.Sq));
    else
      attn_delta_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(
          args->dout, a.o, fmt, a.o_sb, a.o_sh, a.o_ss, args->delta, a.B, a.H, a.Sq, a.D);
    CT_LAUNCH_OK();
    return 0;
  };
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
    const int nqt = (a.Sq + 127) / 128;
    // the workspace is sized for whole query tiles (include/ct_b200.h): B*H*ceil(Sq/128)*128*64 floats, tiled
    // [b][h][query tile][d/4][row][4]
    const size_t dq_elems = (size_t)a.B * a.H * nqt * FB_DQ_TILE;
    CT_CUDA_OK(cudaMemsetAsync(args->dq_accum, 0, sizeof(float) * dq_elems, st));
    if ((rc = launch_delta())) return rc;
    static bool attr = false;
    if (!attr) {
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc2_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_PIPE));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc2_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_PIPE));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc2_kernel<true, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_PIPE));
      CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc2_kernel<false, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_PIPE));
      attr = true;
    }
    const unsigned g = (unsigned)((int64_t)a.B * a.H * ((a.Sk + 127) / 128));
    // ATTN_BWD_IMPL: 0 (default) / 2 = 16 compute warps (one 32-query chunk each), heaviest key tiles first: 154 vs
    // 179 us per call at the bench shape, identical bits (profiles/r02r_ab_attention.jsonl); 1 = 8 compute warps
    if (option(OPT_ATTN_BWD_IMPL) != 1) {
      if (fmt == 1)
        CT_CUDA_OK(launch_k(attn_bwd_tc2_kernel<true, 7>, dim3(g), dim3(FB_THREADS_WIDE), FB_SMEM_PIPE, st, tmQ, tmK, tmV, tmDO, bp));
      else
        CT_CUDA_OK(launch_k(attn_bwd_tc2_kernel<false, 7>, dim3(g), dim3(FB_THREADS_WIDE), FB_SMEM_PIPE, st, tmQ, tmK, tmV, tmDO, bp));
    } else {
      if (fmt == 1)
        CT_CUDA_OK(launch_k(attn_bwd_tc2_kernel<true, 3>, dim3(g), dim3(FB_THREADS), FB_SMEM_PIPE, st, tmQ, tmK, tmV, tmDO, bp));
      else
        CT_CUDA_OK(launch_k(attn_bwd_tc2_kernel<false, 3>, dim3(g), dim3(FB_THREADS), FB_SMEM_PIPE, st, tmQ, tmK, tmV, tmDO, bp));
    }
    CT_LAUNCH_OK();
    const int64_t n = (int64_t)(dq_elems / 8);
    int64_t blocks = (n + 255) / 256;
    if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
    CT_CUDA_OK(launch_k(attn_dq_convert_tiled_kernel, dim3((unsigned)blocks), dim3(256), 0, st, (const float*)args->dq_accum,
                        args->dq, fmt, args->dq_sb, args->dq_sh, args->dq_ss, a.B, a.H, a.Sq));
    return 0;
  }
  if ((rc = launch_delta())) return rc;
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
  CT_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc4_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F4_SMEM));
  CT_CUDA_OK(cudaFuncSetAttribute(attn_bwd_tc2_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_SMEM_PIPE));
  CT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(fwd_ctas_per_sm, attn_fwd_tc4_kernel<true, true, false>, F4_THREADS,
                                                           F4_SMEM));
  CT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(bwd_ctas_per_sm, attn_bwd_tc2_kernel<true, 3>, FB_THREADS,
                                                           FB_SMEM_PIPE));
  if (detail) {
    cudaFuncAttributes fa;
    CT_CUDA_OK(cudaFuncGetAttributes(&fa, attn_fwd_tc4_kernel<true, true, false>));
    detail[0] = fa.numRegs; detail[1] = (int)fa.sharedSizeBytes; detail[2] = F4_SMEM;
    const int cut[5] = {0, 1024, 2048, 4096, 16384};
    for (int i = 0; i < 5; ++i)
      CT_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(detail + 3 + i, attn_fwd_tc4_kernel<true, true, false>,
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

extern "C" int ct_kv_append_dev(const void* src, int64_t s_sb, int64_t s_sh, int64_t s_ss, void* cache,
                                int64_t c_sb, int64_t c_sh, int64_t c_ss, int B, int H, int S_new, int D,
                                const int32_t* len_dev, int t_max, void* stream) {
  CT_REQUIRE(src && cache && len_dev, CT_ERR_BAD_ARG, "ct_kv_append_dev: null pointer");
  CT_REQUIRE(B > 0 && H > 0 && S_new > 0 && D > 0 && t_max >= S_new, CT_ERR_BAD_ARG, "ct_kv_append_dev: bad shape");
  const int64_t n = (int64_t)B * H * S_new * D;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  kv_append_dev_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      (const uint16_t*)src, s_sb, s_sh, s_ss, (uint16_t*)cache, c_sb, c_sh, c_ss, B, H, S_new, D, len_dev, t_max);
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
