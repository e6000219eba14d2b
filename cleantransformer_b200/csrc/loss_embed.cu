// loss_embed.cu — the two ends of the model around the block stack:
//   * embedding gather / scatter-add   (modeling_bloom.py:190, modeling_gpt.py:169,184,
//     modeling_bert.py:297-300 and the autograd scatter they imply)
//   * shifted / plain cross-entropy    (modeling_bloom.py:224-230: torch CrossEntropyLoss, mean over
//     all B*(S-1) shifted positions) producing the loss AND dlogits in one pass over the logits.
// Both are HBM-bound. CE algorithmic bytes: rows * V * (sizeof(logit) read + sizeof(dlogit) write).
#include "ct_common.cuh"
#include "../../include/ct_b200.h"
#include <cfloat>

namespace ct {

// ---------------------------------------------------------------------------------------------
// embedding
// ---------------------------------------------------------------------------------------------
// out[t, :] (+)= W[ids[t], :]   one warp per token, float4 accesses when H % 4 == 0
__global__ void __launch_bounds__(256)
    embedding_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ W,
                         float* __restrict__ out, int64_t T, int64_t H, int64_t V, int accumulate) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= T) return;
  long long id = ids[w];
  float* dst = out + w * H;
  if (id < 0 || id >= V) {
    // torch raises (device-side assert); no host sync here, so the row is poisoned instead: a tokenizer / vocab_size
    // mismatch shows up as a NaN loss on the first step, not as silent training on row 0
    for (int64_t c = lane; c < H; c += 32) dst[c] = __int_as_float(0x7fc00000);
    return;
  }
  const float* src = W + id * H;
  if ((H & 3) == 0) {
    for (int64_t c = lane * 4; c < H; c += 128) {
      float4 v = *reinterpret_cast<const float4*>(src + c);
      if (accumulate) {
        float4 o = *reinterpret_cast<const float4*>(dst + c);
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      *reinterpret_cast<float4*>(dst + c) = v;
    }
  } else {
    for (int64_t c = lane; c < H; c += 32) dst[c] = (accumulate ? dst[c] : 0.f) + src[c];
  }
}

// dW[ids[t], :] += dout[t, :]  (red.global.add; rows equal to padding_idx are skipped like
// torch.nn.Embedding(padding_idx=...) does, modeling_bert.py:273)
__global__ void __launch_bounds__(256)
    embedding_bwd_kernel(const long long* __restrict__ ids, const float* __restrict__ dout,
                         float* __restrict__ dW, int64_t T, int64_t H, int64_t V,
                         long long padding_idx) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= T) return;
  const long long id = ids[w];
  if (id < 0 || id >= V || id == padding_idx) return;
  const float* src = dout + w * H;
  float* dst = dW + id * H;
  if ((H & 3) == 0) {
    for (int64_t c = lane * 4; c < H; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(src + c);
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(v.x), "f"(v.y),
                   "f"(v.z), "f"(v.w)
                   : "memory");
    }
  } else {
    for (int64_t c = lane; c < H; c += 32) atomicAdd(dst + c, src[c]);
  }
}

// ---------------------------------------------------------------------------------------------
// cross entropy
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long ce_target(const long long* labels, int64_t r, int64_t S, int shift) {
  if (!shift) return labels[r];
  // modeling_bloom.py:224-225: logits[..., :-1, :] vs labels[..., 1:]
  return ((r % S) < S - 1) ? labels[r + 1] : -100;
}

// stats[0] = number of rows with a valid target
__global__ void ce_count_kernel(const long long* __restrict__ labels, int64_t rows, int64_t S, int shift,
                                long long ignore_index, int64_t V, float* __restrict__ stats) {
  __shared__ int red[32];
  int cnt = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const long long t = ce_target(labels, r, S, shift);
    cnt += (t != ignore_index && t >= 0 && t < V) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    atomicAdd(stats, (float)s);
  }
}

constexpr int CE_THREADS = 512;

template <typename T>
__device__ __forceinline__ float ce_ld(const T* p, int64_t i);
template <>
__device__ __forceinline__ float ce_ld<float>(const float* p, int64_t i) { return p[i]; }
template <>
__device__ __forceinline__ float ce_ld<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) {
  return __bfloat162float(p[i]);
}
template <>
__device__ __forceinline__ float ce_ld<__half>(const __half* p, int64_t i) { return __half2float(p[i]); }
template <typename T>
struct ce_is_bf16 { static constexpr bool value = false; };
template <>
struct ce_is_bf16<__nv_bfloat16> { static constexpr bool value = true; };
template <typename T>
struct ce_is_f16 { static constexpr bool value = false; };
template <>
struct ce_is_f16<__half> { static constexpr bool value = true; };

// (max, sum of exp relative to max) pairs; a side that saw no element is (-inf, 0) and must not
// produce exp(-inf - -inf) = NaN
__device__ __forceinline__ void ms_merge(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  const float a = (m == mn) ? 1.f : __expf(m - mn);
  const float b = (m2 == mn) ? 1.f : __expf(m2 - mn);
  s = s * a + s2 * b;
  m = mn;
}

__device__ __forceinline__ float2 block_max_sum(float m, float s, float* red) {
  // combine (max, sum of exp relative to max) across the block
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    ms_merge(m, s, m2, s2);
  }
  if (lane == 0) { red[2 * w] = m; red[2 * w + 1] = s; }
  __syncthreads();
  if (w == 0) {
    m = lane < (CE_THREADS >> 5) ? red[2 * lane] : -INFINITY;
    s = lane < (CE_THREADS >> 5) ? red[2 * lane + 1] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      ms_merge(m, s, m2, s2);
    }
    if (lane == 0) { red[0] = m; red[1] = s; }
  }
  __syncthreads();
  const float2 r = make_float2(red[0], red[1]);
  __syncthreads();
  return r;
}

// One CTA per row (grid-stride). Pass 1: online max / sum-exp. Pass 2: dlogits = (softmax - onehot)
// * inv_count (row re-read, normally from L2).
template <typename T>
__global__ void __launch_bounds__(CE_THREADS)
    ce_fwd_kernel(const T* __restrict__ logits, int64_t ld, const long long* __restrict__ labels,
                  T* __restrict__ dlogits, int64_t ldd, float* __restrict__ row_loss,
                  const float* __restrict__ stats, int64_t rows, int64_t V, int64_t S, int shift,
                  long long ignore_index) {
  __shared__ float red[2 * (CE_THREADS >> 5)];
  // f16 (torch.cuda.amp.autocast() + GradScaler, examples/ft_bloom_DDP.py:108-128): 1/count (~1e-4) times a small
  // probability underflows the f16 range before the loss scale can be applied, so the f16 gradient is stored
  // UN-normalised (softmax - onehot) and the caller multiplies by dloss / count in one step (ops.py).
  const float inv_count = ce_is_f16<T>::value ? 1.f : 1.f / fmaxf(stats[0], 1.f);
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const T* x = logits + r * ld;
    const long long tgt = ce_target(labels, r, S, shift);
    const bool valid = (tgt != ignore_index && tgt >= 0 && tgt < V);
    if (!valid) {
      if (row_loss && threadIdx.x == 0) row_loss[r] = 0.f;
      if (dlogits) {
        T* d = dlogits + r * ldd;
        for (int64_t c = threadIdx.x; c < V; c += CE_THREADS) d[c] = (T)0.f;
      }
      continue;
    }
    float m = -INFINITY, s = 0.f;
    if (ce_is_bf16<T>::value && (V & 7) == 0 && (ld & 7) == 0) {
      const uint4* x8 = reinterpret_cast<const uint4*>(x);
      for (int64_t c = threadIdx.x; c < (V >> 3); c += CE_THREADS) {
        const uint4 u = x8[c];
        float v[8];
        float2 f;
        f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
        f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
        f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
        f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
        float mx = v[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) mx = fmaxf(mx, v[i]);
        const float mn = fmaxf(m, mx);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += __expf(v[i] - mn);
        s = s * ((m == mn) ? 1.f : __expf(m - mn)) + acc;
        m = mn;
      }
    } else {
      for (int64_t c = threadIdx.x; c < V; c += CE_THREADS) {
        const float v = ce_ld<T>(x, c);
        const float mn = fmaxf(m, v);
        s = s * ((m == mn) ? 1.f : __expf(m - mn)) + __expf(v - mn);
        m = mn;
      }
    }
    const float2 ms = block_max_sum(m, s, red);
    const float lse = ms.x + logf(ms.y);
    if (row_loss && threadIdx.x == 0) row_loss[r] = lse - ce_ld<T>(x, tgt);
    if (dlogits) {
      T* d = dlogits + r * ldd;
      if (ce_is_bf16<T>::value && (V & 7) == 0 && (ld & 7) == 0 && (ldd & 7) == 0) {
        const uint4* x8 = reinterpret_cast<const uint4*>(x);
        uint4* d8 = reinterpret_cast<uint4*>(d);
        for (int64_t c = threadIdx.x; c < (V >> 3); c += CE_THREADS) {
          const uint4 u = x8[c];
          float v[8];
          float2 f;
          f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
          f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
          f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
          f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
          const int64_t c0 = c << 3;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float pgrad = __expf(v[i] - lse);
            if (c0 + i == tgt) pgrad -= 1.f;
            v[i] = pgrad * inv_count;
          }
          uint4 o;
          o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
          o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
          d8[c] = o;
        }
      } else {
        for (int64_t c = threadIdx.x; c < V; c += CE_THREADS) {
          float pgrad = __expf(ce_ld<T>(x, c) - lse);
          if (c == tgt) pgrad -= 1.f;
          d[c] = (T)(pgrad * inv_count);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// cross entropy, one HBM pass: the row lives in the shared memory of a 4-CTA cluster
// ---------------------------------------------------------------------------------------------
// ce_fwd_kernel reads every logit twice (statistics, then gradients); at V = 250 880 the 592 rows in flight are
// 297 MB — more than L2 — so the second read comes from HBM again: 12.3 GB per Bloom-560M step for 8.2 GB of
// algorithmic traffic (profiles/r01e: 1.98 ms at 6.2 TB/s, i.e. already at the HBM roof for what it moves).
// Here a cluster of CEC_C CTAs owns one row at a time, each CTA keeping a quarter of it (125 KB of bf16) in
// shared memory: the slice arrives through cp.async.bulk in 16 KB chunks (one mbarrier per chunk), the
// statistics pass reads it from shared memory as the chunks land, the CTAs exchange their (max, sum) pairs
// through distributed shared memory, and the gradient pass re-reads the slice from shared memory. As soon as a
// chunk has been consumed by the gradient pass the same chunk of the NEXT row is requested, so the loads of row
// r+1 run under the gradient stores of row r. Exponentials in the log2 domain (ex2.approx), fp32 sums.
constexpr int CEC_C = 4;
constexpr int CEC_THREADS = 1024;
constexpr int CEC_MAX_CHUNKS = 13;  // 13 x 16 KB = 208 KB of shared memory for the slice

__device__ __forceinline__ float ce_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// (max, sum of 2^(t - max)); a side that saw nothing is (-inf, 0)
__device__ __forceinline__ void ms_merge2(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  const float a = (m == mn) ? 1.f : ce_ex2(m - mn);
  const float b = (m2 == mn) ? 1.f : ce_ex2(m2 - mn);
  s = s * a + s2 * b;
  m = mn;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void st_shared_cluster_f2(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}

__global__ void __cluster_dims__(CEC_C, 1, 1) __launch_bounds__(CEC_THREADS, 1)
    ce_fwd_cluster_kernel(const __nv_bfloat16* __restrict__ logits, int64_t ld, const long long* __restrict__ labels,
                          __nv_bfloat16* __restrict__ dlogits, int64_t ldd, float* __restrict__ row_loss,
                          const float* __restrict__ stats, int64_t rows, int64_t V, int64_t S, int shift,
                          long long ignore_index, int slice_vec) {
  extern __shared__ __align__(128) uint8_t ce_smem[];
  __shared__ __align__(8) unsigned long long full_bar[CEC_MAX_CHUNKS];
  __shared__ __align__(8) float xchg[2][CEC_C][2];  // [row parity][cluster rank] = (max2, sum2)
  __shared__ float red[2 * (CEC_THREADS >> 5)];
  const uint32_t buf = smem_u32(ce_smem);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t crank = cluster_ctarank();
  const int64_t cid = blockIdx.x / CEC_C, n_clusters = gridDim.x / CEC_C;
  const int64_t nvec = V >> 3;
  const int64_t v0 = min(nvec, (int64_t)crank * slice_vec);
  const int my_len = (int)(min(nvec, v0 + slice_vec) - v0);  // uint4 (8 logits) in this CTA's slice
  const int n_chunks = (my_len + CEC_THREADS - 1) / CEC_THREADS;
  const float inv_count = 1.f / fmaxf(stats[0], 1.f);
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

  if (tid == 0) {
    for (int k = 0; k < CEC_MAX_CHUNKS; ++k) mbar_init(smem_u32(&full_bar[k]), 1);
    mbar_fence_init();
  }
  __syncthreads();
  auto request = [&](int64_t r, int k) {  // thread 0: chunk k of row r -> shared memory
    const int len = min(CEC_THREADS, my_len - k * CEC_THREADS);
    const uint32_t bar = smem_u32(&full_bar[k]);
    mbar_expect_tx(bar, 16u * len);
    bulk_g2s(buf + 16u * CEC_THREADS * k, logits + r * ld + 8 * (v0 + (int64_t)k * CEC_THREADS), 16u * len, bar);
  };
  if (tid == 0 && cid < rows)
    for (int k = 0; k < n_chunks; ++k) request(cid, k);
  cluster_sync_all();  // peers' barriers and exchange slots exist before anyone writes to them

  uint32_t ph = 0;
  for (int64_t r = cid; r < rows; r += n_clusters, ph ^= 1) {
    const long long tgt = ce_target(labels, r, S, shift);
    const bool valid = (tgt != ignore_index && tgt >= 0 && tgt < V);
    // ---- pass 1: (max, sum) of this thread's elements as the chunks land ----
    float m = -INFINITY, s = 0.f;
    for (int k = 0; k < n_chunks; ++k) {
      mbar_wait(smem_u32(&full_bar[k]), ph);
      const int idx = k * CEC_THREADS + tid;
      if (idx < my_len) {
        const uint4 u = *reinterpret_cast<const uint4*>(ce_smem + 16 * idx);
        float v[8];
        float2 f;
        f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
        f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
        f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
        f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
        float mx = v[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) mx = fmaxf(mx, v[i]);
        mx *= LOG2E;
        if (mx > m) {  // rare once the running maximum has settled
          s *= ce_ex2(m - mx);
          m = mx;
        }
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += ce_ex2(fmaf(v[i], LOG2E, -m));
        s += acc;
      }
    }
    // ---- block, then cluster reduction ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      ms_merge2(m, s, m2, s2);
    }
    if (lane == 0) { red[2 * w] = m; red[2 * w + 1] = s; }
    __syncthreads();
    if (w == 0) {
      m = red[2 * lane]; s = red[2 * lane + 1];  // CEC_THREADS / 32 == 32 warps
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        ms_merge2(m, s, m2, s2);
      }
      if (lane < CEC_C) st_shared_cluster_f2(mapa_shared(smem_u32(&xchg[ph][crank][0]), lane), m, s);
    }
    cluster_sync_all();  // release/acquire: every CTA's pair is visible in every CTA's slots
    m = xchg[ph][0][0]; s = xchg[ph][0][1];
#pragma unroll
    for (int q = 1; q < CEC_C; ++q) ms_merge2(m, s, xchg[ph][q][0], xchg[ph][q][1]);
    const float lse2 = m + log2f(s);
    if (row_loss) {
      if (!valid) {
        if (crank == 0 && tid == 0) row_loss[r] = 0.f;
      } else if (tid == 0 && (tgt >> 3) >= v0 && (tgt >> 3) < v0 + my_len) {
        const __nv_bfloat16 xt = reinterpret_cast<const __nv_bfloat16*>(ce_smem)[tgt - 8 * v0];
        row_loss[r] = lse2 * LN2 - __bfloat162float(xt);
      }
    }
    // ---- pass 2: gradients from shared memory; refill each chunk with the next row as soon as it is consumed ----
    const int64_t r_next = r + n_clusters;
    uint4* d8 = dlogits ? reinterpret_cast<uint4*>(dlogits + r * ldd) + v0 : nullptr;
    for (int k = 0; k < n_chunks; ++k) {
      const int idx = k * CEC_THREADS + tid;
      if (idx < my_len && d8) {
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (valid) {
          const uint4 u = *reinterpret_cast<const uint4*>(ce_smem + 16 * idx);
          float v[8];
          float2 f;
          f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
          f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
          f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
          f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
          const int64_t c0 = (v0 + idx) << 3;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float pg = ce_ex2(fmaf(v[i], LOG2E, -lse2));
            if (c0 + i == tgt) pg -= 1.f;
            v[i] = pg * inv_count;
          }
          o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
          o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        }
        d8[idx] = o;
      }
      __syncthreads();  // every thread is done with chunk k of row r
      if (tid == 0 && r_next < rows) request(r_next, k);
    }
  }
  cluster_sync_all();  // no CTA leaves while a peer could still address its shared memory
}

// ---------------------------------------------------------------------------------------------
// cross entropy, two passes with the row kept in L2
// ---------------------------------------------------------------------------------------------
// Same arithmetic as ce_fwd_kernel, different residency: ONE 1024-thread CTA per SM, so only sm_count rows
// (148 x 0.5 MB = 74 MB at V = 250 880) are between their two passes at any time and the second read hits the
// 126 MB L2 instead of HBM (ce_fwd_kernel keeps 592 rows = 297 MB in flight: 12.3 GB of DRAM traffic per step for
// 8.2 GB of algorithmic bytes). Four independent 16-byte loads per thread per iteration keep ~64 KB in flight
// per SM; the gradient stores and the second read are streaming (.cs) so they do not push the next rows out.
constexpr int CE2_THREADS = 1024;
constexpr int CE2_UNROLL = 4;

__device__ __forceinline__ void ce_unpack8(const uint4& u, float (&v)[8]) {
  float2 f;
  f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
}

__global__ void __launch_bounds__(CE2_THREADS, 1)
    ce_fwd_l2_kernel(const __nv_bfloat16* __restrict__ logits, int64_t ld, const long long* __restrict__ labels,
                     __nv_bfloat16* __restrict__ dlogits, int64_t ldd, float* __restrict__ row_loss,
                     const float* __restrict__ stats, int64_t rows, int64_t V, int64_t S, int shift,
                     long long ignore_index) {
  __shared__ float red[2 * (CE2_THREADS >> 5)];
  const float inv_count = 1.f / fmaxf(stats[0], 1.f);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t nvec = V >> 3;
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const long long tgt = ce_target(labels, r, S, shift);
    const bool valid = (tgt != ignore_index && tgt >= 0 && tgt < V);
    const uint4* x8 = reinterpret_cast<const uint4*>(logits + r * ld);
    uint4* d8 = dlogits ? reinterpret_cast<uint4*>(dlogits + r * ldd) : nullptr;
    if (!valid) {
      if (row_loss && tid == 0) row_loss[r] = 0.f;
      if (d8)
        for (int64_t c = tid; c < nvec; c += CE2_THREADS) __stcs(d8 + c, make_uint4(0u, 0u, 0u, 0u));
      continue;
    }
    // ---- pass 1: (max, sum 2^(t - max)) in the log2 domain ----
    float m = -INFINITY, s = 0.f;
    for (int64_t c0 = tid; c0 < nvec; c0 += CE2_THREADS * CE2_UNROLL) {
      uint4 u[CE2_UNROLL];
#pragma unroll
      for (int k = 0; k < CE2_UNROLL; ++k) {
        const int64_t c = c0 + (int64_t)k * CE2_THREADS;
        u[k] = c < nvec ? x8[c] : make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);  // -inf pairs
      }
#pragma unroll
      for (int k = 0; k < CE2_UNROLL; ++k) {
        float v[8];
        ce_unpack8(u[k], v);
        float mx = v[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) mx = fmaxf(mx, v[i]);
        mx *= LOG2E;
        if (mx > m) {
          s *= ce_ex2(m - mx);
          m = mx;
        }
        if (m > -INFINITY) {
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc += ce_ex2(fmaf(v[i], LOG2E, -m));
          s += acc;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      ms_merge2(m, s, m2, s2);
    }
    if (lane == 0) { red[2 * w] = m; red[2 * w + 1] = s; }
    __syncthreads();
    m = red[2 * lane]; s = red[2 * lane + 1];  // 32 warps: every warp reduces the same 32 pairs
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      ms_merge2(m, s, m2, s2);
    }
    __syncthreads();  // red is rewritten by the next row
    const float lse2 = m + log2f(s);
    if (row_loss && tid == 0) row_loss[r] = lse2 * LN2 - __bfloat162float(logits[r * ld + tgt]);
    // ---- pass 2: gradients; the row comes back from L2 ----
    if (d8) {
      for (int64_t c0 = tid; c0 < nvec; c0 += CE2_THREADS * CE2_UNROLL) {
        uint4 u[CE2_UNROLL];
#pragma unroll
        for (int k = 0; k < CE2_UNROLL; ++k) {
          const int64_t c = c0 + (int64_t)k * CE2_THREADS;
          if (c < nvec) u[k] = __ldcs(x8 + c);
        }
#pragma unroll
        for (int k = 0; k < CE2_UNROLL; ++k) {
          const int64_t c = c0 + (int64_t)k * CE2_THREADS;
          if (c < nvec) {
            float v[8];
            ce_unpack8(u[k], v);
            const int64_t e0 = c << 3;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float pg = ce_ex2(fmaf(v[i], LOG2E, -lse2));
              if (e0 + i == tgt) pg -= 1.f;
              v[i] = pg * inv_count;
            }
            uint4 o;
            o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
            o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
            __stcs(d8 + c, o);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// cross entropy from statistics the LM-head GEMM already produced (ct_gemm_args.row_stats): ONE streaming pass
// ---------------------------------------------------------------------------------------------
// The epilogue of the logits GEMM sees every logit in registers; it leaves, per row, 2*ceil(V/256) partial
// (max2, sum2) pairs of the bf16-rounded values (128 MB for Bloom-560M's 8192 x 250 880 logits, 3 % of the logits
// themselves). Merging them gives the row's log-sum-exp without reading the row, so this kernel is a single pass:
// read logit, write (softmax - onehot) / count. SURVEY §8 (f) N1, first half.
constexpr int CES_THREADS = 512;

__global__ void __launch_bounds__(CES_THREADS)
    ce_fwd_stats_kernel(const __nv_bfloat16* __restrict__ logits, int64_t ld, const long long* __restrict__ labels,
                        __nv_bfloat16* __restrict__ dlogits, int64_t ldd, float* __restrict__ row_loss,
                        const float* __restrict__ stats, const float2* __restrict__ row_stats, int n_slots,
                        int64_t rows, int64_t V, int64_t S, int shift, long long ignore_index) {
  __shared__ float red[2 * (CES_THREADS >> 5)];
  const float inv_count = 1.f / fmaxf(stats[0], 1.f);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t nvec = V >> 3;
  constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const long long tgt = ce_target(labels, r, S, shift);
    const bool valid = (tgt != ignore_index && tgt >= 0 && tgt < V);
    const uint4* x8 = reinterpret_cast<const uint4*>(logits + r * ld);
    uint4* d8 = dlogits ? reinterpret_cast<uint4*>(dlogits + r * ldd) : nullptr;
    if (!valid) {
      if (row_loss && tid == 0) row_loss[r] = 0.f;
      if (d8)
        for (int64_t c = tid; c < nvec; c += CES_THREADS) __stcs(d8 + c, make_uint4(0u, 0u, 0u, 0u));
      continue;
    }
    float m = -INFINITY, s = 0.f;
    for (int k = tid; k < n_slots; k += CES_THREADS) {
      const float2 q = __ldg(row_stats + (int64_t)k * rows + r);
      ms_merge2(m, s, q.x, q.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      ms_merge2(m, s, m2, s2);
    }
    if (lane == 0) { red[2 * w] = m; red[2 * w + 1] = s; }
    __syncthreads();
    m = lane < (CES_THREADS >> 5) ? red[2 * lane] : -INFINITY;
    s = lane < (CES_THREADS >> 5) ? red[2 * lane + 1] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      ms_merge2(m, s, m2, s2);
    }
    __syncthreads();  // red is rewritten by the next row
    const float lse2 = m + log2f(s);
    if (row_loss && tid == 0) row_loss[r] = lse2 * LN2 - __bfloat162float(logits[r * ld + tgt]);
    if (d8) {
      for (int64_t c0 = tid; c0 < nvec; c0 += CES_THREADS * CE2_UNROLL) {
        uint4 u[CE2_UNROLL];
#pragma unroll
        for (int k = 0; k < CE2_UNROLL; ++k) {
          const int64_t c = c0 + (int64_t)k * CES_THREADS;
          if (c < nvec) u[k] = __ldcs(x8 + c);
        }
#pragma unroll
        for (int k = 0; k < CE2_UNROLL; ++k) {
          const int64_t c = c0 + (int64_t)k * CES_THREADS;
          if (c < nvec) {
            float v[8];
            ce_unpack8(u[k], v);
            const int64_t e0 = c << 3;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float pg = ce_ex2(fmaf(v[i], LOG2E, -lse2));
              if (e0 + i == tgt) pg -= 1.f;
              v[i] = pg * inv_count;
            }
            uint4 o;
            o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
            o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
            __stcs(d8 + c, o);
          }
        }
      }
    }
  }
}

// loss = sum(row_loss) / count   (single block, deterministic order)
__global__ void __launch_bounds__(1024)
    ce_finalize_kernel(const float* __restrict__ row_loss, int64_t rows, const float* __restrict__ stats,
                       float* __restrict__ loss) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t r = threadIdx.x; r < rows; r += 1024) acc += row_loss[r];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = red[threadIdx.x];
    v = warp_sum(v);
    // no valid target at all: 0/0 = NaN like torch.nn.CrossEntropyLoss (an all-masked batch must not look like loss 0)
    if (threadIdx.x == 0) loss[0] = v / stats[0];
  }
}

// x *= *scalar unless *scalar == 1 (upstream dloss, e.g. a GradScaler factor)
template <typename T>
__global__ void __launch_bounds__(256)
    scale_by_device_scalar_kernel(T* __restrict__ x, int64_t n, const float* __restrict__ scalar) {
  const float s = *scalar;
  if (s == 1.f) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = (T)((float)x[i] * s);
}


// ---------------------------------------------------------------------------------------------
// One step of GenerationMixin._greedy_search after the model call (generation_util.py:86-101, do_sample = False),
// entirely on the device so that a whole decode step can be replayed from a CUDA graph:
//   next = argmax(logits[b, :])  (first maximum, like torch.argmax) — or sampled[b] when the caller drew the token itself
//   (do_sample = True: generation_util.py:78-84);  next = next * alive + pad * (1 - alive);
//   alive[b] &= next not in end_ids;  ids_out[b, out_pos] = next;  cur_ids[b] = next;  pos_ids[b] += 1;
//   once per call (last block): out_pos += 1, seq_len += 1, done_at = out_pos when no row is alive any more.
// state (int32): [0] seq_len (cache length the next model call sees, its own token included), [1] out_pos,
//                [2] alive rows, [3] done_at (-1 until every row has finished), [4] block counter (0 between calls).
template <typename T>
__global__ void __launch_bounds__(1024)
    greedy_step_kernel(const T* __restrict__ logits, int64_t ld, int64_t V, long long* __restrict__ alive,
                       const long long* __restrict__ end_ids, int n_end, long long pad_id,
                       long long* __restrict__ ids_out, int64_t out_stride, long long* __restrict__ cur_ids,
                       long long* __restrict__ pos_ids, int* __restrict__ state,
                       const long long* __restrict__ sampled) {
  const int b = blockIdx.x;
  const T* row = logits + (int64_t)b * ld;
  float best = -INFINITY;
  long long arg = 0x7fffffffffffffffLL;
  const int64_t v_scan = sampled ? 0 : V;  // a token drawn by the caller: no scan
  for (int64_t j = threadIdx.x; j < v_scan; j += blockDim.x) {
    const float v = (float)row[j];
    if (v > best || (v == best && j < arg) || (v != v && !(best != best))) { best = v; arg = j; }  // NaN wins like torch
  }
  auto better = [](float v, long long i, float bv, long long bi) {
    const bool vn = v != v, bn = bv != bv;
    if (vn != bn) return vn;
    if (vn) return i < bi;
    return v > bv || (v == bv && i < bi);
  };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, arg, o);
    if (better(ov, oi, best, arg)) { best = ov; arg = oi; }
  }
  __shared__ float sv[32];
  __shared__ long long si[32];
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = arg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (better(sv[w], si[w], best, arg)) { best = sv[w]; arg = si[w]; }
    if (sampled) arg = sampled[b];
    const long long was_alive = alive[b];
    const long long next = arg * was_alive + pad_id * (1 - was_alive);
    bool hit = false;
    for (int e = 0; e < n_end; ++e) hit |= (next == end_ids[e]);
    const int out_pos = state[1];
    ids_out[(int64_t)b * out_stride + out_pos] = next;
    cur_ids[b] = next;
    if (pos_ids) pos_ids[b] += 1;
    if (hit && was_alive) {
      alive[b] = 0;
      atomicSub(&state[2], 1);
    }
    __threadfence();
    const int done = atomicAdd(&state[4], 1);
    if (done == (int)gridDim.x - 1) {  // every row has read out_pos and updated the alive count
      __threadfence();
      state[4] = 0;
      state[0] += 1;
      state[1] = out_pos + 1;
      if (atomicAdd(&state[2], 0) <= 0 && state[3] < 0) state[3] = out_pos + 1;
    }
  }
}

}  // namespace ct

using namespace ct;

extern "C" int ct_embedding_fwd(const int64_t* ids, const float* weight, float* out, int64_t T,
                                int64_t H, int64_t V, int accumulate, void* stream) {
  CT_REQUIRE(ids && weight && out, CT_ERR_BAD_ARG, "ct_embedding_fwd: null pointer");
  if (T <= 0) return 0;
  embedding_fwd_kernel<<<(unsigned)((T * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const long long*)ids, weight, out, T, H, V, accumulate);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_embedding_bwd(const int64_t* ids, const float* dout, float* dweight, int64_t T,
                                int64_t H, int64_t V, int64_t padding_idx, void* stream) {
  CT_REQUIRE(ids && dout && dweight, CT_ERR_BAD_ARG, "ct_embedding_bwd: null pointer");
  if (T <= 0) return 0;
  embedding_bwd_kernel<<<(unsigned)((T * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const long long*)ids, dout, dweight, T, H, V, (long long)padding_idx);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_cross_entropy_fwd(const void* logits, int dtype, int64_t ld, const int64_t* labels,
                                    void* dlogits, int64_t ldd, float* loss, float* workspace,
                                    int64_t rows, int64_t V, int64_t S, int shift,
                                    int64_t ignore_index, void* stream) {
  CT_REQUIRE(logits && labels && loss && workspace, CT_ERR_BAD_ARG, "ct_cross_entropy_fwd: null pointer");
  CT_REQUIRE(dtype == DT_F32 || dtype == DT_BF16 || dtype == DT_F16, CT_ERR_UNSUPPORTED, "ct_cross_entropy_fwd: dtype");
  CT_REQUIRE(rows > 0 && V > 0 && (!shift || (S > 0 && rows % S == 0)), CT_ERR_BAD_ARG,
             "ct_cross_entropy_fwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  float* stats = workspace;        // [0] = valid-row count
  float* row_loss = workspace + 4; // [rows]
  CT_CUDA_OK(cudaMemsetAsync(stats, 0, 16, st));
  ce_count_kernel<<<64, 256, 0, st>>>((const long long*)labels, rows, S, shift, (long long)ignore_index, V, stats);
  CT_LAUNCH_OK();
  // CE_IMPL: 0 = auto (3 when the row is 16-byte addressable bf16), 1 = two-pass kernel, 592 rows in flight (second read from HBM at Bloom's vocabulary),
  //          2 = row resident in cluster shared memory (measured slower, r01g), 3 = two passes, one row per SM
  //          in flight so that the second read hits L2
  const int64_t nvec = V >> 3;
  const int slice_vec = (int)((nvec + CEC_C - 1) / CEC_C);
  const bool cluster_ok = dtype == DT_BF16 && (V & 7) == 0 && (ld & 7) == 0 && (!dlogits || (ldd & 7) == 0) &&
                          ((uintptr_t)logits & 15) == 0 && ((uintptr_t)dlogits & 15) == 0 &&
                          slice_vec <= CEC_MAX_CHUNKS * CEC_THREADS;
  const int ce_impl = option(OPT_CE_IMPL);
  const bool vec_ok = dtype == DT_BF16 && (V & 7) == 0 && (ld & 7) == 0 && (!dlogits || (ldd & 7) == 0) &&
                      ((uintptr_t)logits & 15) == 0 && ((uintptr_t)dlogits & 15) == 0;
  if (vec_ok && (ce_impl == 3 || ce_impl == 0)) {
    int64_t grid = sm_count();
    if (grid > rows) grid = rows;
    ce_fwd_l2_kernel<<<(unsigned)grid, CE2_THREADS, 0, st>>>(
        (const __nv_bfloat16*)logits, ld, (const long long*)labels, (__nv_bfloat16*)dlogits, ldd, row_loss, stats,
        rows, V, S, shift, (long long)ignore_index);
    CT_LAUNCH_OK();
  } else if (cluster_ok && ce_impl == 2) {
    const size_t smem = (size_t)16 * ((slice_vec + CEC_THREADS - 1) / CEC_THREADS) * CEC_THREADS;
    static bool attr = false;
    if (!attr) {
      CT_CUDA_OK(cudaFuncSetAttribute(ce_fwd_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      16 * CEC_MAX_CHUNKS * CEC_THREADS));
      attr = true;
    }
    int64_t clusters = sm_count() / CEC_C;
    if (clusters > rows) clusters = rows;
    if (clusters < 1) clusters = 1;
    ce_fwd_cluster_kernel<<<(unsigned)(clusters * CEC_C), CEC_THREADS, smem, st>>>(
        (const __nv_bfloat16*)logits, ld, (const long long*)labels, (__nv_bfloat16*)dlogits, ldd, row_loss, stats,
        rows, V, S, shift, (long long)ignore_index, slice_vec);
    CT_LAUNCH_OK();
  } else {
  int64_t grid = rows;
  const int64_t cap = (int64_t)sm_count() * 4;
  if (grid > cap) grid = cap;
  if (dtype == DT_BF16)
    ce_fwd_kernel<__nv_bfloat16><<<(unsigned)grid, CE_THREADS, 0, st>>>(
        (const __nv_bfloat16*)logits, ld, (const long long*)labels, (__nv_bfloat16*)dlogits, ldd, row_loss,
        stats, rows, V, S, shift, (long long)ignore_index);
  else if (dtype == DT_F16)
    ce_fwd_kernel<__half><<<(unsigned)grid, CE_THREADS, 0, st>>>(
        (const __half*)logits, ld, (const long long*)labels, (__half*)dlogits, ldd, row_loss, stats, rows, V, S, shift,
        (long long)ignore_index);
  else
    ce_fwd_kernel<float><<<(unsigned)grid, CE_THREADS, 0, st>>>(
        (const float*)logits, ld, (const long long*)labels, (float*)dlogits, ldd, row_loss, stats, rows, V, S,
        shift, (long long)ignore_index);
  CT_LAUNCH_OK();
  }
  ce_finalize_kernel<<<1, 1024, 0, st>>>(row_loss, rows, stats, loss);
  CT_LAUNCH_OK();
  return 0;
}

// bf16 logits whose per-row softmax statistics came out of the producing GEMM (ct_gemm_args.row_stats): same contract
// as ct_cross_entropy_fwd otherwise. n_slots = 2 * ceil(V / 256).
extern "C" int ct_cross_entropy_fwd_stats(const void* logits, int64_t ld, const int64_t* labels, void* dlogits,
                                          int64_t ldd, float* loss, float* workspace, const float* row_stats,
                                          int64_t n_slots, int64_t rows, int64_t V, int64_t S, int shift,
                                          int64_t ignore_index, void* stream) {
  CT_REQUIRE(logits && labels && loss && workspace && row_stats, CT_ERR_BAD_ARG,
             "ct_cross_entropy_fwd_stats: null pointer");
  CT_REQUIRE(rows > 0 && V > 0 && (V & 7) == 0 && (ld & 7) == 0 && (!dlogits || (ldd & 7) == 0) &&
                 ((uintptr_t)logits & 15) == 0 && ((uintptr_t)dlogits & 15) == 0 && ((uintptr_t)row_stats & 7) == 0 &&
                 n_slots == 2 * ((V + 255) / 256) && (!shift || (S > 0 && rows % S == 0)),
             CT_ERR_BAD_ARG, "ct_cross_entropy_fwd_stats: bad shape / alignment");
  cudaStream_t st = (cudaStream_t)stream;
  float* stats = workspace;
  float* row_loss = workspace + 4;
  CT_CUDA_OK(cudaMemsetAsync(stats, 0, 16, st));
  ce_count_kernel<<<64, 256, 0, st>>>((const long long*)labels, rows, S, shift, (long long)ignore_index, V, stats);
  CT_LAUNCH_OK();
  int64_t grid = (int64_t)sm_count() * 4;
  if (grid > rows) grid = rows;
  ce_fwd_stats_kernel<<<(unsigned)grid, CES_THREADS, 0, st>>>(
      (const __nv_bfloat16*)logits, ld, (const long long*)labels, (__nv_bfloat16*)dlogits, ldd, row_loss, stats,
      (const float2*)row_stats, (int)n_slots, rows, V, S, shift, (long long)ignore_index);
  CT_LAUNCH_OK();
  ce_finalize_kernel<<<1, 1024, 0, st>>>(row_loss, rows, stats, loss);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_scale_by_scalar(void* x, int dtype, int64_t n, const float* device_scalar, void* stream) {
  CT_REQUIRE(x && device_scalar, CT_ERR_BAD_ARG, "ct_scale_by_scalar: null pointer");
  CT_REQUIRE(dtype == DT_F32 || dtype == DT_BF16 || dtype == DT_F16, CT_ERR_UNSUPPORTED, "ct_scale_by_scalar: dtype");
  if (n <= 0) return 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
  if (dtype == DT_BF16)
    scale_by_device_scalar_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (__nv_bfloat16*)x, n, device_scalar);
  else if (dtype == DT_F16)
    scale_by_device_scalar_kernel<__half><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((__half*)x, n, device_scalar);
  else
    scale_by_device_scalar_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((float*)x, n,
                                                                                           device_scalar);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_greedy_step(const void* logits, int logits_dtype, int64_t ld, int64_t B, int64_t V, int64_t* alive,
                              const int64_t* end_ids, int n_end, int64_t pad_id, int64_t* ids_out, int64_t out_stride,
                              int64_t* cur_ids, int64_t* pos_ids, int32_t* state, const int64_t* sampled, void* stream) {
  CT_REQUIRE((logits || sampled) && alive && ids_out && cur_ids && state && (n_end == 0 || end_ids), CT_ERR_BAD_ARG,
             "ct_greedy_step: null pointer");
  CT_REQUIRE(B > 0 && V > 0 && ld >= V && n_end >= 0, CT_ERR_BAD_ARG, "ct_greedy_step: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (logits_dtype == DT_F32)
    greedy_step_kernel<float><<<(unsigned)B, 1024, 0, st>>>((const float*)logits, ld, V, (long long*)alive,
                                                          (const long long*)end_ids, n_end, (long long)pad_id,
                                                          (long long*)ids_out, out_stride, (long long*)cur_ids,
                                                          (long long*)pos_ids, state, (const long long*)sampled);
  else if (logits_dtype == DT_BF16)
    greedy_step_kernel<__nv_bfloat16><<<(unsigned)B, 1024, 0, st>>>((const __nv_bfloat16*)logits, ld, V, (long long*)alive,
                                                                  (const long long*)end_ids, n_end, (long long)pad_id,
                                                                  (long long*)ids_out, out_stride, (long long*)cur_ids,
                                                                  (long long*)pos_ids, state, (const long long*)sampled);
  else if (logits_dtype == DT_F16)
    greedy_step_kernel<__half><<<(unsigned)B, 1024, 0, st>>>((const __half*)logits, ld, V, (long long*)alive,
                                                           (const long long*)end_ids, n_end, (long long)pad_id,
                                                           (long long*)ids_out, out_stride, (long long*)cur_ids,
                                                           (long long*)pos_ids, state, (const long long*)sampled);
  else
    CT_REQUIRE(false, CT_ERR_UNSUPPORTED, "ct_greedy_step: logits must be f32 / bf16 / f16");
  CT_LAUNCH_OK();
  return 0;
}
