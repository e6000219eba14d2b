// layernorm.cu — fused LayerNorm forward/backward (HBM-bound; one warp per row, 128-bit accesses,
// fp32 statistics via warp shuffles).
//
// Replaces the ~9 eager elementwise/reduce kernels of CleanTransformer/transformer.py:71-89
// (LayerNorm._mean + forward): mean = sum(x)/N; std = sqrt(mean((x-mean)^2 + eps));
// y = w * (x - mean)/std + b. Note eps sits inside the mean, i.e. std^2 = var + eps (biased var),
// identical to torch.nn.LayerNorm. The two-pass (mean, then centred squares) order of the
// reference is kept.
//
// Algorithmic bytes per row (cols = H): fwd  H*(sizeof(x) + sizeof(y) [+ sizeof(y2)]) + 8
//                                       bwd  H*(sizeof(dy) + sizeof(x) + sizeof(dx) [+4 dx_add]) + 8
#include "ct_common.cuh"
#include "../../include/ct_b200.h"
#include <cstdlib>
#include <cstring>

namespace ct {

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  static __device__ __forceinline__ float4 load(const float* p) {
    return *reinterpret_cast<const float4*>(p);
  }
  static __device__ __forceinline__ void store(float* p, float4 v) {
    *reinterpret_cast<float4*>(p) = v;
  }
};
template <>
struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 load(const __nv_bfloat16* p) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, float4 v) {
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

template <>
struct Vec4<__half> {
  static __device__ __forceinline__ float4 load(const __half* p) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ void store(__half* p, float4 v) {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

// run-time typed accessors: f32, bf16 (the default autocast dtype) or f16 (torch.cuda.amp.autocast(),
// examples/ft_bloom_DDP.py:108-128)
__device__ __forceinline__ float4 load4_dyn(const void* base, int dtype, int64_t idx) {
  if (dtype == DT_F32) return Vec4<float>::load(reinterpret_cast<const float*>(base) + idx);
  if (dtype == DT_BF16) return Vec4<__nv_bfloat16>::load(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
  return Vec4<__half>::load(reinterpret_cast<const __half*>(base) + idx);
}
__device__ __forceinline__ void store4_dyn(void* base, int dtype, int64_t idx, float4 v) {
  if (dtype == DT_F32)
    Vec4<float>::store(reinterpret_cast<float*>(base) + idx, v);
  else if (dtype == DT_BF16)
    Vec4<__nv_bfloat16>::store(reinterpret_cast<__nv_bfloat16*>(base) + idx, v);
  else
    Vec4<__half>::store(reinterpret_cast<__half*>(base) + idx, v);
}
__device__ __forceinline__ float load1_dyn(const void* base, int dtype, int64_t idx) {
  if (dtype == DT_F32) return reinterpret_cast<const float*>(base)[idx];
  if (dtype == DT_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
  return __half2float(reinterpret_cast<const __half*>(base)[idx]);
}
__device__ __forceinline__ void store1_dyn(void* base, int dtype, int64_t idx, float v) {
  if (dtype == DT_F32)
    reinterpret_cast<float*>(base)[idx] = v;
  else if (dtype == DT_BF16)
    reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
  else
    reinterpret_cast<__half*>(base)[idx] = __float2half_rn(v);
}

constexpr int LN_WARPS = 8;

// ---- forward, cols == 128 * VPL (row held in registers) ------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32)
    ln_fwd_vec_kernel(const void* __restrict__ x, int x_dtype, const float* __restrict__ gamma,
                      const float* __restrict__ beta, void* __restrict__ y, int y_dtype,
                      void* __restrict__ y2, int y2_dtype, float* __restrict__ mean_out,
                      float* __restrict__ rstd_out, int64_t rows, float eps) {
  constexpr int COLS = VPL * 128;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * LN_WARPS;
  pdl_wait();               // programmatic dependent launch: see ct_common.cuh
  pdl_launch_dependents();

  float4 g[VPL], b[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    g[i] = Vec4<float>::load(gamma + i * 128 + lane * 4);
    b[i] = Vec4<float>::load(beta + i * 128 + lane * 4);
  }
  for (int64_t row = warp; row < rows; row += nwarps) {
    float4 v[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      v[i] = load4_dyn(x, x_dtype, row * COLS + i * 128 + lane * 4);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.f / COLS);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float var = warp_sum(q) * (1.f / COLS) + eps;
    const float rstd = rsqrtf(var);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float4 o;
      o.x = fmaf(v[i].x * rstd, g[i].x, b[i].x);
      o.y = fmaf(v[i].y * rstd, g[i].y, b[i].y);
      o.z = fmaf(v[i].z * rstd, g[i].z, b[i].z);
      o.w = fmaf(v[i].w * rstd, g[i].w, b[i].w);
      const int64_t idx = row * COLS + i * 128 + lane * 4;
      if (y) store4_dyn(y, y_dtype, idx, o);
      if (y2) store4_dyn(y2, y2_dtype, idx, o);
    }
  }
}

// ---- embedding gather (up to three tables) + LayerNorm in one pass -------------------
// modeling_bloom.py:190-191 (word_embeddings -> word_embeddings_layernorm), modeling_bert.py:297-301
// (word + segment + position tables -> embedding_post LayerNorm). One warp per token: the table rows are
// summed in registers, the sum is written once (fp32, only when a backward will need it as the LayerNorm
// input) and normalised without ever being read back. An id outside its table poisons the row with NaN,
// like embedding_fwd_kernel. Algorithmic bytes per token: H*4*(tables [+1 emb]) + H*(sizeof(y) [+ sizeof(y2)]).
struct EmbedLnTables {
  const long long* ids[3];
  const float* w[3];
  long long vocab[3];
  int n;
};

template <int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32)
    embed_ln_fwd_kernel(EmbedLnTables tb, const float* __restrict__ gamma, const float* __restrict__ beta,
                        float* __restrict__ emb, void* __restrict__ y, int y_dtype, void* __restrict__ y2,
                        int y2_dtype, float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows,
                        float eps) {
  constexpr int COLS = VPL * 128;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * LN_WARPS;
  float4 g[VPL], b[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    g[i] = Vec4<float>::load(gamma + i * 128 + lane * 4);
    b[i] = Vec4<float>::load(beta + i * 128 + lane * 4);
  }
  for (int64_t row = warp; row < rows; row += nwarps) {
    float4 v[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    bool bad = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (k < tb.n) {
        const long long id = tb.ids[k][row];
        if (id < 0 || id >= tb.vocab[k]) {
          bad = true;
        } else {
          const float* src = tb.w[k] + id * COLS + lane * 4;
#pragma unroll
          for (int i = 0; i < VPL; ++i) {
            const float4 t = Vec4<float>::load(src + i * 128);
            v[i].x += t.x; v[i].y += t.y; v[i].z += t.z; v[i].w += t.w;
          }
        }
      }
    }
    if (bad) {
      const float nan = __int_as_float(0x7fc00000);
#pragma unroll
      for (int i = 0; i < VPL; ++i) v[i] = make_float4(nan, nan, nan, nan);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (emb) Vec4<float>::store(emb + row * COLS + i * 128 + lane * 4, v[i]);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.f / COLS);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float var = warp_sum(q) * (1.f / COLS) + eps;
    const float rstd = rsqrtf(var);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float4 o;
      o.x = fmaf(v[i].x * rstd, g[i].x, b[i].x);
      o.y = fmaf(v[i].y * rstd, g[i].y, b[i].y);
      o.z = fmaf(v[i].z * rstd, g[i].z, b[i].z);
      o.w = fmaf(v[i].w * rstd, g[i].w, b[i].w);
      const int64_t idx = row * COLS + i * 128 + lane * 4;
      if (y) store4_dyn(y, y_dtype, idx, o);
      if (y2) store4_dyn(y2, y2_dtype, idx, o);
    }
  }
}

// ---- forward, arbitrary cols (row re-read from L1/L2) -------------------------------------------
__global__ void __launch_bounds__(LN_WARPS * 32)
    ln_fwd_generic_kernel(const void* __restrict__ x, int x_dtype, const float* __restrict__ gamma,
                          const float* __restrict__ beta, void* __restrict__ y, int y_dtype,
                          void* __restrict__ y2, int y2_dtype, float* __restrict__ mean_out,
                          float* __restrict__ rstd_out, int64_t rows, int64_t cols, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * LN_WARPS;
  for (int64_t row = warp; row < rows; row += nwarps) {
    float s = 0.f;
    for (int64_t c = lane; c < cols; c += 32) s += load1_dyn(x, x_dtype, row * cols + c);
    const float mean = warp_sum(s) / (float)cols;
    float q = 0.f;
    for (int64_t c = lane; c < cols; c += 32) {
      float d = load1_dyn(x, x_dtype, row * cols + c) - mean;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)cols + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
    for (int64_t c = lane; c < cols; c += 32) {
      float o = fmaf((load1_dyn(x, x_dtype, row * cols + c) - mean) * rstd, gamma[c], beta[c]);
      if (y) store1_dyn(y, y_dtype, row * cols + c, o);
      if (y2) store1_dyn(y2, y2_dtype, row * cols + c, o);
    }
  }
}

// ---- backward -----------------------------------------------------------------------------------
// dx = rstd * (gy - mean(gy) - xhat * mean(gy * xhat)),  gy = gamma * dy,  xhat = (x - mean) * rstd
// dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy   (per-warp register partials -> smem ->
// one atomicAdd per column per CTA)
// dgamma/dbeta partials live in shared memory (one private [2][COLS] slice per warp, each lane
// only ever touches its own columns, so no synchronisation is needed until the final reduction).
// Keeping them in registers cost 64 registers per thread (194 total) and limited the kernel to 8
// warps per SM; with ~100 registers two CTAs are resident.
template <int VPL>
__global__ void __launch_bounds__(LN_WARPS * 32, 2)
    ln_bwd_vec_kernel(const void* __restrict__ dy, int dy_dtype, const void* __restrict__ dy2,
                      int dy2_dtype, const void* __restrict__ x, int x_dtype,
                      const float* __restrict__ gamma, const float* __restrict__ mean_in,
                      const float* __restrict__ rstd_in, const void* __restrict__ dx_add,
                      int dx_add_dtype, void* __restrict__ dx, int dx_dtype,
                      float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows,
                      float* __restrict__ partial /* [gridDim.x][2][COLS] or null (-> atomics) */) {
  constexpr int COLS = VPL * 128;
  extern __shared__ float acc_s[];  // [LN_WARPS][2][COLS]
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * LN_WARPS + wib;
  const int64_t nwarps = (int64_t)gridDim.x * LN_WARPS;
  float* my_dg = acc_s + (size_t)wib * 2 * COLS;
  float* my_db = my_dg + COLS;
  const bool need_gb = (dgamma != nullptr) || (dbeta != nullptr);

  float4 g[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    g[i] = Vec4<float>::load(gamma + i * 128 + lane * 4);
    if (need_gb) {
      Vec4<float>::store(my_dg + i * 128 + lane * 4, make_float4(0.f, 0.f, 0.f, 0.f));
      Vec4<float>::store(my_db + i * 128 + lane * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
  for (int64_t row = warp; row < rows; row += nwarps) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    float4 xh[VPL], d[VPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int64_t idx = row * COLS + i * 128 + lane * 4;
      float4 xv = load4_dyn(x, x_dtype, idx);
      if (dy) {
        d[i] = load4_dyn(dy, dy_dtype, idx);
        if (dy2) {
          float4 e = load4_dyn(dy2, dy2_dtype, idx);
          d[i].x += e.x; d[i].y += e.y; d[i].z += e.z; d[i].w += e.w;
        }
      } else {
        d[i] = load4_dyn(dy2, dy2_dtype, idx);
      }
      xh[i].x = (xv.x - mean) * rstd; xh[i].y = (xv.y - mean) * rstd;
      xh[i].z = (xv.z - mean) * rstd; xh[i].w = (xv.w - mean) * rstd;
    }
    if (need_gb) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        float4 a = Vec4<float>::load(my_dg + i * 128 + lane * 4);
        float4 b = Vec4<float>::load(my_db + i * 128 + lane * 4);
        a.x = fmaf(d[i].x, xh[i].x, a.x); a.y = fmaf(d[i].y, xh[i].y, a.y);
        a.z = fmaf(d[i].z, xh[i].z, a.z); a.w = fmaf(d[i].w, xh[i].w, a.w);
        b.x += d[i].x; b.y += d[i].y; b.z += d[i].z; b.w += d[i].w;
        Vec4<float>::store(my_dg + i * 128 + lane * 4, a);
        Vec4<float>::store(my_db + i * 128 + lane * 4, b);
      }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      d[i].x *= g[i].x; d[i].y *= g[i].y; d[i].z *= g[i].z; d[i].w *= g[i].w;
      s1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      s2 += (d[i].x * xh[i].x + d[i].y * xh[i].y) + (d[i].z * xh[i].z + d[i].w * xh[i].w);
    }
    s1 = warp_sum(s1) * (1.f / COLS);
    s2 = warp_sum(s2) * (1.f / COLS);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int64_t idx = row * COLS + i * 128 + lane * 4;
      float4 o;
      o.x = rstd * (d[i].x - s1 - xh[i].x * s2);
      o.y = rstd * (d[i].y - s1 - xh[i].y * s2);
      o.z = rstd * (d[i].z - s1 - xh[i].z * s2);
      o.w = rstd * (d[i].w - s1 - xh[i].w * s2);
      if (dx_add) {
        float4 a = load4_dyn(dx_add, dx_add_dtype, idx);
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
      }
      store4_dyn(dx, dx_dtype, idx, o);
    }
  }
  if (!need_gb) return;
  __syncthreads();
  // cross-warp reduction: thread t sums column t, t + 256, ... over the LN_WARPS slices
  for (int c = threadIdx.x; c < COLS; c += LN_WARPS * 32) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) {
      sg += acc_s[(size_t)w * 2 * COLS + c];
      sb += acc_s[(size_t)w * 2 * COLS + COLS + c];
    }
    if (partial) {  // deterministic two-stage reduction: no same-address atomics at the tail
      partial[((size_t)blockIdx.x * 2 + 0) * COLS + c] = sg;
      partial[((size_t)blockIdx.x * 2 + 1) * COLS + c] = sb;
    } else {
      if (dgamma) atomicAdd(dgamma + c, sg);
      if (dbeta) atomicAdd(dbeta + c, sb);
    }
  }
}

// second stage: out[c] (+)= sum_b partial[b][which][c]
__global__ void __launch_bounds__(256)
    ln_bwd_reduce_kernel(const float* __restrict__ partial, int nblocks, int cols, float* __restrict__ dgamma,
                         float* __restrict__ dbeta, int accumulate) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int which = blockIdx.y;
  float* out = which == 0 ? dgamma : dbeta;
  if (c >= cols || out == nullptr) return;
  float acc = 0.f;
  for (int b = 0; b < nblocks; ++b) acc += partial[((size_t)b * 2 + which) * cols + c];
  out[c] = accumulate ? out[c] + acc : acc;
}

// ---- backward v2: a row is spread over cols/4 threads (one float4 per thread and tensor) -----------
// The warp-per-row kernel above keeps 3 x 32 values per lane in registers (128 registers, 16 warps per
// SM) and measured 54 us at [8192,1024] against a 21 us HBM floor: too few loads in flight. Here a
// CTA is two independent halves of cols/4 threads; each half handles LNB_R rows per iteration, so a
// thread holds LNB_R float4 per tensor (<= 64 registers, 32 warps per SM) and OWNS four fixed
// columns: the dgamma / dbeta / column-sum(dx) partials stay in registers for the whole kernel. The
// two row reductions go warp-shuffle -> smem -> one named barrier per iteration (slots double-buffered
// by iteration parity). Per-CTA partials are combined by ln_bwd_reduce3_kernel (deterministic).
constexpr int LNB_R = 2;

struct LnBwdP {
  const void* dy; int dy_dtype;
  const void* dy2; int dy2_dtype;
  const void* x; int x_dtype;
  const float* gamma; const float* mean; const float* rstd;
  const void* dx_add; int dx_add_dtype;
  void* dx; int dx_dtype;
  void* dx2; int dx2_dtype;
  float* dgamma; float* dbeta; float* dxsum;
  float* partial;  // [gridDim.x][3][cols] or null (-> atomics into dgamma / dbeta / dxsum)
  int64_t rows;
  int cols;
};

__global__ void __launch_bounds__(512, 2) ln_bwd_cta_kernel(const LnBwdP p) {
  __shared__ float red[2][2][2 * LNB_R][8];
  __shared__ __align__(16) float comb[3 * 1024];
  const int tpr = blockDim.x >> 1;  // threads per row = cols / 4
  const int half = threadIdx.x >= tpr ? 1 : 0;
  const int t = threadIdx.x - half * tpr;
  const int lane = t & 31, wih = t >> 5, nwh = tpr >> 5;
  const int col0 = 4 * t;
  const float inv_cols = 1.f / (float)p.cols;
  const bool need_part = p.dgamma || p.dbeta || p.dxsum;

  const float4 g4 = Vec4<float>::load(p.gamma + col0);
  float4 adg = make_float4(0.f, 0.f, 0.f, 0.f), adb = adg, adc = adg;
  int it = 0;
  for (int64_t grp = (int64_t)blockIdx.x * 2 + half; grp * LNB_R < p.rows; grp += (int64_t)gridDim.x * 2, ++it) {
    float4 xh[LNB_R], d[LNB_R], av[LNB_R];
    float rs[LNB_R];
    float part[2 * LNB_R];
#pragma unroll
    for (int r = 0; r < LNB_R; ++r) {
      const int64_t row = grp * LNB_R + r;
      const bool ok = row < p.rows;
      const int64_t idx = row * p.cols + col0;
      float mean = 0.f;
      rs[r] = 0.f;
      xh[r] = make_float4(0.f, 0.f, 0.f, 0.f); d[r] = xh[r]; av[r] = xh[r];
      if (ok) {
        xh[r] = load4_dyn(p.x, p.x_dtype, idx);
        if (p.dy) {
          d[r] = load4_dyn(p.dy, p.dy_dtype, idx);
          if (p.dy2) {
            const float4 e = load4_dyn(p.dy2, p.dy2_dtype, idx);
            d[r].x += e.x; d[r].y += e.y; d[r].z += e.z; d[r].w += e.w;
          }
        } else {
          d[r] = load4_dyn(p.dy2, p.dy2_dtype, idx);
        }
        if (p.dx_add) av[r] = load4_dyn(p.dx_add, p.dx_add_dtype, idx);
        mean = __ldg(p.mean + row);
        rs[r] = __ldg(p.rstd + row);
      }
      xh[r].x = (xh[r].x - mean) * rs[r]; xh[r].y = (xh[r].y - mean) * rs[r];
      xh[r].z = (xh[r].z - mean) * rs[r]; xh[r].w = (xh[r].w - mean) * rs[r];
    }
#pragma unroll
    for (int r = 0; r < LNB_R; ++r) {
      adg.x = fmaf(d[r].x, xh[r].x, adg.x); adg.y = fmaf(d[r].y, xh[r].y, adg.y);
      adg.z = fmaf(d[r].z, xh[r].z, adg.z); adg.w = fmaf(d[r].w, xh[r].w, adg.w);
      adb.x += d[r].x; adb.y += d[r].y; adb.z += d[r].z; adb.w += d[r].w;
      d[r].x *= g4.x; d[r].y *= g4.y; d[r].z *= g4.z; d[r].w *= g4.w;
      part[2 * r] = (d[r].x + d[r].y) + (d[r].z + d[r].w);
      part[2 * r + 1] = (d[r].x * xh[r].x + d[r].y * xh[r].y) + (d[r].z * xh[r].z + d[r].w * xh[r].w);
    }
#pragma unroll
    for (int q = 0; q < 2 * LNB_R; ++q) part[q] = warp_sum(part[q]);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < 2 * LNB_R; ++q) red[half][it & 1][q][wih] = part[q];
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "r"(tpr) : "memory");
#pragma unroll
    for (int q = 0; q < 2 * LNB_R; ++q) {
      float s = 0.f;
      for (int w = 0; w < nwh; ++w) s += red[half][it & 1][q][w];
      part[q] = s * inv_cols;
    }
#pragma unroll
    for (int r = 0; r < LNB_R; ++r) {
      const int64_t row = grp * LNB_R + r;
      if (row >= p.rows) continue;
      const int64_t idx = row * p.cols + col0;
      const float s1 = part[2 * r], s2 = part[2 * r + 1];
      float4 o;
      o.x = fmaf(rs[r], d[r].x - s1 - xh[r].x * s2, av[r].x);
      o.y = fmaf(rs[r], d[r].y - s1 - xh[r].y * s2, av[r].y);
      o.z = fmaf(rs[r], d[r].z - s1 - xh[r].z * s2, av[r].z);
      o.w = fmaf(rs[r], d[r].w - s1 - xh[r].w * s2, av[r].w);
      store4_dyn(p.dx, p.dx_dtype, idx, o);
      if (p.dx2) store4_dyn(p.dx2, p.dx2_dtype, idx, o);
      adc.x += o.x; adc.y += o.y; adc.z += o.z; adc.w += o.w;
    }
  }
  if (!need_part) return;
  // combine the two halves, then one partial row set per CTA
  if (half == 1) {
    Vec4<float>::store(comb + col0, adg);
    Vec4<float>::store(comb + p.cols + col0, adb);
    Vec4<float>::store(comb + 2 * p.cols + col0, adc);
  }
  __syncthreads();
  if (half == 0) {
    const float4 a = Vec4<float>::load(comb + col0), b = Vec4<float>::load(comb + p.cols + col0),
                 c = Vec4<float>::load(comb + 2 * p.cols + col0);
    adg.x += a.x; adg.y += a.y; adg.z += a.z; adg.w += a.w;
    adb.x += b.x; adb.y += b.y; adb.z += b.z; adb.w += b.w;
    adc.x += c.x; adc.y += c.y; adc.z += c.z; adc.w += c.w;
    if (p.partial) {
      float* dst = p.partial + (size_t)blockIdx.x * 3 * p.cols + col0;
      if (p.dgamma) Vec4<float>::store(dst, adg);
      if (p.dbeta) Vec4<float>::store(dst + p.cols, adb);
      if (p.dxsum) Vec4<float>::store(dst + 2 * p.cols, adc);
    } else {
      const float vg[4] = {adg.x, adg.y, adg.z, adg.w}, vb[4] = {adb.x, adb.y, adb.z, adb.w},
                  vc[4] = {adc.x, adc.y, adc.z, adc.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (p.dgamma) atomicAdd(p.dgamma + col0 + k, vg[k]);
        if (p.dbeta) atomicAdd(p.dbeta + col0 + k, vb[k]);
        if (p.dxsum) atomicAdd(p.dxsum + col0 + k, vc[k]);
      }
    }
  }
}

// Compile-time variant of ln_bwd_cta_kernel for the training hot path (x f32, dy bf16, optional f32 dx_add,
// f32 dx, optional bf16 dx2): ncu on the run-time-typed kernel above showed it ISSUE-bound (60 % issue
// utilisation, ~1100 instructions per two-row iteration, mostly dtype dispatch and 64-bit address math).
template <int NW, bool HAS_ADD, bool HAS_DX2>
__global__ void __launch_bounds__(NW * 64, 2)
    ln_bwd_fast_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                       const float* __restrict__ gamma, const float* __restrict__ mean_in,
                       const float* __restrict__ rstd_in, const float* __restrict__ dx_add, float* __restrict__ dx,
                       __nv_bfloat16* __restrict__ dx2, float* __restrict__ dgamma, float* __restrict__ dbeta,
                       float* __restrict__ dxsum, float* __restrict__ partial, int64_t rows) {
  constexpr int COLS = NW * 128, TPR = NW * 32;
  __shared__ __align__(16) float red[2][2][4][8];  // [half][iteration parity][quantity][warp (zero padded)]
  __shared__ __align__(16) float comb[3 * COLS];
  pdl_wait();               // programmatic dependent launch: see ct_common.cuh
  pdl_launch_dependents();
  const int half = threadIdx.x >= TPR ? 1 : 0;
  const int t = threadIdx.x - half * TPR;
  const int lane = t & 31, wih = t >> 5;
  const int col0 = 4 * t;
  constexpr float inv_cols = 1.f / (float)COLS;
  if (threadIdx.x < 2 * 2 * 4 * 8) (&red[0][0][0][0])[threadIdx.x] = 0.f;
  __syncthreads();

  const float4 g4 = Vec4<float>::load(gamma + col0);
  float4 adg = make_float4(0.f, 0.f, 0.f, 0.f), adb = adg, adc = adg;
  const int64_t n_grp = (rows + 1) / 2;
  int par = 0;
  for (int64_t grp = (int64_t)blockIdx.x * 2 + half; grp < n_grp; grp += (int64_t)gridDim.x * 2, par ^= 1) {
    const int64_t row0 = grp * 2;
    const bool ok1 = row0 + 1 < rows;
    const int64_t i0 = row0 * COLS + col0, i1 = i0 + COLS;
    float4 xh0 = Vec4<float>::load(x + i0);
    float4 d0 = Vec4<__nv_bfloat16>::load(dy + i0);
    float4 xh1 = make_float4(0.f, 0.f, 0.f, 0.f), d1 = xh1, a0 = xh1, a1 = xh1;
    if (ok1) { xh1 = Vec4<float>::load(x + i1); d1 = Vec4<__nv_bfloat16>::load(dy + i1); }
    if constexpr (HAS_ADD) {
      a0 = Vec4<float>::load(dx_add + i0);
      if (ok1) a1 = Vec4<float>::load(dx_add + i1);
    }
    const float m0 = __ldg(mean_in + row0), r0 = __ldg(rstd_in + row0);
    const float m1 = ok1 ? __ldg(mean_in + row0 + 1) : 0.f, r1 = ok1 ? __ldg(rstd_in + row0 + 1) : 0.f;
    xh0.x = (xh0.x - m0) * r0; xh0.y = (xh0.y - m0) * r0; xh0.z = (xh0.z - m0) * r0; xh0.w = (xh0.w - m0) * r0;
    xh1.x = (xh1.x - m1) * r1; xh1.y = (xh1.y - m1) * r1; xh1.z = (xh1.z - m1) * r1; xh1.w = (xh1.w - m1) * r1;
    adg.x = fmaf(d0.x, xh0.x, adg.x); adg.y = fmaf(d0.y, xh0.y, adg.y);
    adg.z = fmaf(d0.z, xh0.z, adg.z); adg.w = fmaf(d0.w, xh0.w, adg.w);
    adg.x = fmaf(d1.x, xh1.x, adg.x); adg.y = fmaf(d1.y, xh1.y, adg.y);
    adg.z = fmaf(d1.z, xh1.z, adg.z); adg.w = fmaf(d1.w, xh1.w, adg.w);
    adb.x += d0.x + d1.x; adb.y += d0.y + d1.y; adb.z += d0.z + d1.z; adb.w += d0.w + d1.w;
    d0.x *= g4.x; d0.y *= g4.y; d0.z *= g4.z; d0.w *= g4.w;
    d1.x *= g4.x; d1.y *= g4.y; d1.z *= g4.z; d1.w *= g4.w;
    float p0 = (d0.x + d0.y) + (d0.z + d0.w);
    float p1 = (d0.x * xh0.x + d0.y * xh0.y) + (d0.z * xh0.z + d0.w * xh0.w);
    float p2 = (d1.x + d1.y) + (d1.z + d1.w);
    float p3 = (d1.x * xh1.x + d1.y * xh1.y) + (d1.z * xh1.z + d1.w * xh1.w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      p0 += __shfl_xor_sync(0xffffffffu, p0, o); p1 += __shfl_xor_sync(0xffffffffu, p1, o);
      p2 += __shfl_xor_sync(0xffffffffu, p2, o); p3 += __shfl_xor_sync(0xffffffffu, p3, o);
    }
    if (lane == 0) {
      red[half][par][0][wih] = p0; red[half][par][1][wih] = p1;
      red[half][par][2][wih] = p2; red[half][par][3][wih] = p3;
    }
    asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "n"(TPR) : "memory");
    float tot[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 u = *reinterpret_cast<const float4*>(&red[half][par][q][0]);
      const float4 v = *reinterpret_cast<const float4*>(&red[half][par][q][4]);
      tot[q] = ((u.x + u.y) + (u.z + u.w) + (v.x + v.y) + (v.z + v.w)) * inv_cols;
    }
    float4 o0, o1;
    o0.x = fmaf(r0, d0.x - tot[0] - xh0.x * tot[1], a0.x); o0.y = fmaf(r0, d0.y - tot[0] - xh0.y * tot[1], a0.y);
    o0.z = fmaf(r0, d0.z - tot[0] - xh0.z * tot[1], a0.z); o0.w = fmaf(r0, d0.w - tot[0] - xh0.w * tot[1], a0.w);
    o1.x = fmaf(r1, d1.x - tot[2] - xh1.x * tot[3], a1.x); o1.y = fmaf(r1, d1.y - tot[2] - xh1.y * tot[3], a1.y);
    o1.z = fmaf(r1, d1.z - tot[2] - xh1.z * tot[3], a1.z); o1.w = fmaf(r1, d1.w - tot[2] - xh1.w * tot[3], a1.w);
    Vec4<float>::store(dx + i0, o0);
    if constexpr (HAS_DX2) Vec4<__nv_bfloat16>::store(dx2 + i0, o0);
    adc.x += o0.x; adc.y += o0.y; adc.z += o0.z; adc.w += o0.w;
    if (ok1) {
      Vec4<float>::store(dx + i1, o1);
      if constexpr (HAS_DX2) Vec4<__nv_bfloat16>::store(dx2 + i1, o1);
      adc.x += o1.x; adc.y += o1.y; adc.z += o1.z; adc.w += o1.w;
    }
  }
  if (!(dgamma || dbeta || dxsum)) return;
  if (half == 1) {
    Vec4<float>::store(comb + col0, adg);
    Vec4<float>::store(comb + COLS + col0, adb);
    Vec4<float>::store(comb + 2 * COLS + col0, adc);
  }
  __syncthreads();
  if (half == 0) {
    const float4 a = Vec4<float>::load(comb + col0), b = Vec4<float>::load(comb + COLS + col0),
                 c = Vec4<float>::load(comb + 2 * COLS + col0);
    adg.x += a.x; adg.y += a.y; adg.z += a.z; adg.w += a.w;
    adb.x += b.x; adb.y += b.y; adb.z += b.z; adb.w += b.w;
    adc.x += c.x; adc.y += c.y; adc.z += c.z; adc.w += c.w;
    float* dst = partial + (size_t)blockIdx.x * 3 * COLS + col0;
    if (dgamma) Vec4<float>::store(dst, adg);
    if (dbeta) Vec4<float>::store(dst + COLS, adb);
    if (dxsum) Vec4<float>::store(dst + 2 * COLS, adc);
  }
}

template <int NW>
static int launch_ln_bwd_fast(const ct_ln_bwd_args& a, int grid, cudaStream_t st) {
  const float* x = reinterpret_cast<const float*>(a.x);
  const __nv_bfloat16* dy = reinterpret_cast<const __nv_bfloat16*>(a.dy);
  const float* add = reinterpret_cast<const float*>(a.dx_add);
  float* dx = reinterpret_cast<float*>(a.dx);
  __nv_bfloat16* dx2 = reinterpret_cast<__nv_bfloat16*>(a.dx2);
#define CT_LNF(A, D)                                                                                         \
  CT_CUDA_OK(launch_k(ln_bwd_fast_kernel<NW, A, D>, dim3(grid), dim3(NW * 64), 0, st, x, dy, a.gamma, a.mean, a.rstd, add, \
                      dx, dx2, a.dgamma, a.dbeta, a.dxsum, a.workspace, a.rows))
  if (add && dx2) CT_LNF(true, true);
  else if (add) CT_LNF(true, false);
  else if (dx2) CT_LNF(false, true);
  else CT_LNF(false, false);
#undef CT_LNF
  return 0;
}

// second stage of v2: out_which[c] (+)= sum_b partial[b][which][c]; CTA = 32 columns x 8 row groups
__global__ void __launch_bounds__(256)
    ln_bwd_reduce3_kernel(const float* __restrict__ partial, int nblocks, int cols, float* __restrict__ dgamma,
                          float* __restrict__ dbeta, float* __restrict__ dxsum, int acc_gb, int acc_sum) {
  __shared__ float sm[8][33];
  pdl_wait();               // programmatic dependent launch: see ct_common.cuh
  pdl_launch_dependents();
  const int which = blockIdx.y;
  float* out = which == 0 ? dgamma : (which == 1 ? dbeta : dxsum);
  if (out == nullptr) return;  // CTA-uniform
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < cols) {
    const float* src = partial + (size_t)which * cols + c;
    const size_t stride = (size_t)3 * cols;
    int b = ty;
    for (; b + 24 < nblocks; b += 32) {
      a0 += src[(size_t)b * stride]; a1 += src[(size_t)(b + 8) * stride];
      a2 += src[(size_t)(b + 16) * stride]; a3 += src[(size_t)(b + 24) * stride];
    }
    for (; b < nblocks; b += 8) a0 += src[(size_t)b * stride];
  }
  sm[ty][tx] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (ty == 0 && c < cols) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sm[k][tx];
    const int acc = which == 2 ? acc_sum : acc_gb;
    out[c] = acc ? out[c] + s : s;
  }
}

__global__ void __launch_bounds__(LN_WARPS * 32)
    ln_bwd_generic_kernel(const void* __restrict__ dy, int dy_dtype, const void* __restrict__ dy2,
                          int dy2_dtype, const void* __restrict__ x, int x_dtype,
                          const float* __restrict__ gamma, const float* __restrict__ mean_in,
                          const float* __restrict__ rstd_in, const void* __restrict__ dx_add,
                          int dx_add_dtype, void* __restrict__ dx, int dx_dtype,
                          float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows,
                          int64_t cols) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * LN_WARPS;
  for (int64_t row = warp; row < rows; row += nwarps) {
    const float mean = mean_in[row], rstd = rstd_in[row];
    float s1 = 0.f, s2 = 0.f;
    for (int64_t c = lane; c < cols; c += 32) {
      const int64_t idx = row * cols + c;
      float d = (dy ? load1_dyn(dy, dy_dtype, idx) : 0.f) + (dy2 ? load1_dyn(dy2, dy2_dtype, idx) : 0.f);
      float xh = (load1_dyn(x, x_dtype, idx) - mean) * rstd;
      if (dgamma) atomicAdd(dgamma + c, d * xh);
      if (dbeta) atomicAdd(dbeta + c, d);
      d *= gamma[c];
      s1 += d;
      s2 += d * xh;
    }
    s1 = warp_sum(s1) / (float)cols;
    s2 = warp_sum(s2) / (float)cols;
    for (int64_t c = lane; c < cols; c += 32) {
      const int64_t idx = row * cols + c;
      float d = (dy ? load1_dyn(dy, dy_dtype, idx) : 0.f) + (dy2 ? load1_dyn(dy2, dy2_dtype, idx) : 0.f);
      d *= gamma[c];
      float xh = (load1_dyn(x, x_dtype, idx) - mean) * rstd;
      float o = rstd * (d - s1 - xh * s2);
      if (dx_add) o += load1_dyn(dx_add, dx_add_dtype, idx);
      store1_dyn(dx, dx_dtype, idx, o);
    }
  }
}

static bool dt_ok(int dt) { return dt == DT_F32 || dt == DT_BF16 || dt == DT_F16; }
static bool aligned16(const void* p) { return p == nullptr || ((uintptr_t)p & 15) == 0; }

static int ln_grid(int64_t rows) {
  int64_t need = (rows + LN_WARPS - 1) / LN_WARPS;
  int64_t cap = (int64_t)sm_count() * 8;  // 8 CTAs x 8 warps per SM = full occupancy
  if (need > cap) need = cap;
  if (need < 1) need = 1;
  return (int)need;
}

}  // namespace ct

using namespace ct;

extern "C" int ct_layernorm_fwd(const void* x, int x_dtype, const float* gamma, const float* beta,
                                void* y, int y_dtype, void* y2, int y2_dtype, float* mean,
                                float* rstd, int64_t rows, int64_t cols, float eps, void* stream) {
  CT_REQUIRE(rows >= 0 && cols > 0, CT_ERR_BAD_ARG, "ct_layernorm_fwd: bad shape %lld x %lld",
             (long long)rows, (long long)cols);
  if (rows == 0) return 0;  // empty input: nothing to do (pointers may legitimately be null)
  CT_REQUIRE(x && gamma && beta && (y || y2), CT_ERR_BAD_ARG, "ct_layernorm_fwd: null pointer");
  CT_REQUIRE(dt_ok(x_dtype) && (!y || dt_ok(y_dtype)) && (!y2 || dt_ok(y2_dtype)),
             CT_ERR_UNSUPPORTED, "ct_layernorm_fwd: dtype must be f32, bf16 or f16");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(rows);
  const bool vec = (cols % 128 == 0) && cols <= 1024 && aligned16(x) && aligned16(y) &&
                   aligned16(y2) && aligned16(gamma) && aligned16(beta);
#define CT_LN_FWD(V)                                                                           \
  case V:                                                                                      \
    CT_CUDA_OK(launch_k(ln_fwd_vec_kernel<V>, dim3(grid), dim3(LN_WARPS * 32), 0, st, x, x_dtype, gamma, beta, y, \
                        y_dtype, y2, y2_dtype, mean, rstd, rows, eps));                                         \
    break;
  if (vec) {
    switch ((int)(cols / 128)) {
      CT_LN_FWD(1) CT_LN_FWD(2) CT_LN_FWD(3) CT_LN_FWD(4) CT_LN_FWD(5) CT_LN_FWD(6) CT_LN_FWD(7)
      CT_LN_FWD(8)
    }
  } else {
    ln_fwd_generic_kernel<<<grid, LN_WARPS * 32, 0, st>>>(x, x_dtype, gamma, beta, y, y_dtype, y2,
                                                          y2_dtype, mean, rstd, rows, cols, eps);
  }
#undef CT_LN_FWD
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_embedding_layernorm_fwd(const int64_t* ids0, const float* table0, int64_t vocab0,
                                          const int64_t* ids1, const float* table1, int64_t vocab1,
                                          const int64_t* ids2, const float* table2, int64_t vocab2,
                                          const float* gamma, const float* beta, float* emb, void* y, int y_dtype,
                                          void* y2, int y2_dtype, float* mean, float* rstd, int64_t rows,
                                          int64_t cols, float eps, void* stream) {
  CT_REQUIRE(rows >= 0 && cols > 0, CT_ERR_BAD_ARG, "ct_embedding_layernorm_fwd: bad shape %lld x %lld",
             (long long)rows, (long long)cols);
  if (rows == 0) return 0;
  CT_REQUIRE(ids0 && table0 && gamma && beta && (y || y2), CT_ERR_BAD_ARG,
             "ct_embedding_layernorm_fwd: null pointer");
  CT_REQUIRE((ids1 == nullptr) == (table1 == nullptr) && (ids2 == nullptr) == (table2 == nullptr) &&
                 (ids1 != nullptr || ids2 == nullptr),
             CT_ERR_BAD_ARG, "ct_embedding_layernorm_fwd: tables are given in order, each with its ids");
  CT_REQUIRE((!y || dt_ok(y_dtype)) && (!y2 || dt_ok(y2_dtype)), CT_ERR_UNSUPPORTED,
             "ct_embedding_layernorm_fwd: dtype must be f32, bf16 or f16");
  CT_REQUIRE((cols % 128 == 0) && cols <= 1024 && aligned16(table0) && aligned16(table1) && aligned16(table2) &&
                 aligned16(emb) && aligned16(y) && aligned16(y2) && aligned16(gamma) && aligned16(beta),
             CT_ERR_UNSUPPORTED,
             "ct_embedding_layernorm_fwd: needs cols %% 128 == 0, cols <= 1024 and 16-byte aligned pointers "
             "(use ct_embedding_fwd + ct_layernorm_fwd otherwise)");
  EmbedLnTables tb;
  tb.ids[0] = reinterpret_cast<const long long*>(ids0); tb.w[0] = table0; tb.vocab[0] = vocab0;
  tb.ids[1] = reinterpret_cast<const long long*>(ids1); tb.w[1] = table1; tb.vocab[1] = vocab1;
  tb.ids[2] = reinterpret_cast<const long long*>(ids2); tb.w[2] = table2; tb.vocab[2] = vocab2;
  tb.n = 1 + (ids1 != nullptr) + (ids2 != nullptr);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ln_grid(rows);
#define CT_EMBED_LN(V)                                                                                   \
  case V:                                                                                                \
    embed_ln_fwd_kernel<V><<<grid, LN_WARPS * 32, 0, st>>>(tb, gamma, beta, emb, y, y_dtype, y2, y2_dtype, \
                                                           mean, rstd, rows, eps);                       \
    break;
  switch ((int)(cols / 128)) {
    CT_EMBED_LN(1) CT_EMBED_LN(2) CT_EMBED_LN(3) CT_EMBED_LN(4) CT_EMBED_LN(5) CT_EMBED_LN(6) CT_EMBED_LN(7)
    CT_EMBED_LN(8)
  }
#undef CT_EMBED_LN
  CT_LAUNCH_OK();
  return 0;
}

static int ln_bwd_v1(const void* dy, int dy_dtype, const void* dy2, int dy2_dtype,
                                const void* x, int x_dtype, const float* gamma, const float* mean,
                                const float* rstd, const void* dx_add, int dx_add_dtype, void* dx,
                                int dx_dtype, float* dgamma, float* dbeta, int dgb_accumulate,
                                float* workspace, size_t workspace_bytes, int64_t rows, int64_t cols,
                                void* stream) {
  CT_REQUIRE((dy || dy2) && x && gamma && mean && rstd && dx, CT_ERR_BAD_ARG,
             "ct_layernorm_bwd: null pointer");
  CT_REQUIRE(rows >= 0 && cols > 0, CT_ERR_BAD_ARG, "ct_layernorm_bwd: bad shape");
  CT_REQUIRE(dt_ok(x_dtype) && dt_ok(dx_dtype) && (!dy || dt_ok(dy_dtype)) &&
                 (!dy2 || dt_ok(dy2_dtype)) && (!dx_add || dt_ok(dx_add_dtype)),
             CT_ERR_UNSUPPORTED, "ct_layernorm_bwd: dtype must be f32, bf16 or f16");
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec_shape = (cols % 128 == 0) && cols <= 1024;
  int grid_probe = ln_grid(rows);
  if (grid_probe > sm_count() * 2) grid_probe = sm_count() * 2;
  // two-stage (workspace) reduction of dgamma/dbeta when the caller provides scratch space
  const bool use_ws = vec_shape && workspace != nullptr && (dgamma || dbeta) &&
                      workspace_bytes >= (size_t)grid_probe * 2 * cols * sizeof(float) && rows > 0;
  if (!dgb_accumulate && !use_ws) {
    if (dgamma) CT_CUDA_OK(cudaMemsetAsync(dgamma, 0, sizeof(float) * cols, st));
    if (dbeta) CT_CUDA_OK(cudaMemsetAsync(dbeta, 0, sizeof(float) * cols, st));
  }
  if (rows == 0) return 0;
  // fewer, fatter CTAs than forward: each CTA ends with 2*cols atomics
  int grid = ln_grid(rows);
  const int cap = sm_count() * 2;  // 2 resident CTAs per SM, each ends with 2*cols atomics
  if (grid > cap) grid = cap;
  const bool vec = (cols % 128 == 0) && cols <= 1024 && aligned16(x) && aligned16(dy) &&
                   aligned16(dy2) && aligned16(dx) && aligned16(dx_add) && aligned16(gamma);
#define CT_LN_BWD(V)                                                                             \
  case V: {                                                                                      \
    const int smem = LN_WARPS * 2 * V * 128 * (int)sizeof(float);                                \
    static bool attr = false;                                                                    \
    if (!attr) {                                                                                 \
      CT_CUDA_OK(cudaFuncSetAttribute(ln_bwd_vec_kernel<V>,                                      \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, smem));       \
      attr = true;                                                                               \
    }                                                                                            \
    ln_bwd_vec_kernel<V><<<grid, LN_WARPS * 32, smem, st>>>(dy, dy_dtype, dy2, dy2_dtype, x,     \
                                                            x_dtype, gamma, mean, rstd, dx_add,  \
                                                            dx_add_dtype, dx, dx_dtype, dgamma,  \
                                                            dbeta, rows, use_ws ? workspace : nullptr); \
  } break;
  if (vec) {
    switch ((int)(cols / 128)) {
      CT_LN_BWD(1) CT_LN_BWD(2) CT_LN_BWD(3) CT_LN_BWD(4) CT_LN_BWD(5) CT_LN_BWD(6) CT_LN_BWD(7)
      CT_LN_BWD(8)
    }
  } else {
    ln_bwd_generic_kernel<<<grid, LN_WARPS * 32, 0, st>>>(dy, dy_dtype, dy2, dy2_dtype, x, x_dtype,
                                                          gamma, mean, rstd, dx_add, dx_add_dtype,
                                                          dx, dx_dtype, dgamma, dbeta, rows, cols);
  }
#undef CT_LN_BWD
  CT_LAUNCH_OK();
  if (use_ws) {
    dim3 g2((unsigned)((cols + 255) / 256), 2);
    ln_bwd_reduce_kernel<<<g2, 256, 0, st>>>(workspace, grid, (int)cols, dgamma, dbeta, dgb_accumulate);
    CT_LAUNCH_OK();
  }
  return 0;
}

extern "C" int ct_layernorm_bwd_ex(const ct_ln_bwd_args* args, void* stream) {
  CT_REQUIRE(args != nullptr, CT_ERR_BAD_ARG, "ct_layernorm_bwd_ex: null args");
  const ct_ln_bwd_args& a = *args;
  CT_REQUIRE((a.dy || a.dy2) && a.x && a.gamma && a.mean && a.rstd && a.dx, CT_ERR_BAD_ARG,
             "ct_layernorm_bwd: null pointer");
  CT_REQUIRE(a.rows >= 0 && a.cols > 0, CT_ERR_BAD_ARG, "ct_layernorm_bwd: bad shape");
  CT_REQUIRE(dt_ok(a.x_dtype) && dt_ok(a.dx_dtype) && (!a.dy || dt_ok(a.dy_dtype)) &&
                 (!a.dy2 || dt_ok(a.dy2_dtype)) && (!a.dx_add || dt_ok(a.dx_add_dtype)) &&
                 (!a.dx2 || dt_ok(a.dx2_dtype)),
             CT_ERR_UNSUPPORTED, "ct_layernorm_bwd: dtype must be f32, bf16 or f16");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = a.rows, cols = a.cols;
  const bool vec_shape = (cols % 128 == 0) && cols <= 1024;
  const bool vec = vec_shape && aligned16(a.x) && aligned16(a.dy) && aligned16(a.dy2) && aligned16(a.dx) &&
                   aligned16(a.dx2) && aligned16(a.dx_add) && aligned16(a.gamma);
  const bool want_red = a.dgamma || a.dbeta || a.dxsum;

  if (vec && option(OPT_LN_BWD_IMPL) != 1) {  // LN_BWD_IMPL: 0 = auto (v2), 1 = warp-per-row kernel
    // ---- v2: cols/4 threads per row, two halves per CTA ----
    int64_t groups = (rows + 2 * LNB_R - 1) / (2 * LNB_R);
    int grid = (int)(groups < (int64_t)sm_count() * 2 ? groups : (int64_t)sm_count() * 2);
    if (grid < 1) grid = 1;
    const bool use_ws = want_red && a.workspace != nullptr &&
                        a.workspace_bytes >= (size_t)grid * 3 * cols * sizeof(float) && rows > 0;
    if (want_red && !use_ws) {
      if (a.dgamma && !a.dgb_accumulate) CT_CUDA_OK(cudaMemsetAsync(a.dgamma, 0, sizeof(float) * cols, st));
      if (a.dbeta && !a.dgb_accumulate) CT_CUDA_OK(cudaMemsetAsync(a.dbeta, 0, sizeof(float) * cols, st));
      if (a.dxsum && !a.dxsum_accumulate) CT_CUDA_OK(cudaMemsetAsync(a.dxsum, 0, sizeof(float) * cols, st));
    }
    if (rows == 0) return 0;
    // compile-time hot variant (LN_BWD_IMPL 2 forces the run-time-typed kernel of the same design)
    const bool fast = option(OPT_LN_BWD_IMPL) != 2 && (cols == 1024 || cols == 768) && a.x_dtype == DT_F32 && a.dy &&
                      a.dy_dtype == DT_BF16 && !a.dy2 && (!a.dx_add || a.dx_add_dtype == DT_F32) &&
                      a.dx_dtype == DT_F32 && (!a.dx2 || a.dx2_dtype == DT_BF16) && (!want_red || use_ws);
    if (fast) {
      const int lrc = cols == 1024 ? launch_ln_bwd_fast<8>(a, grid, st) : launch_ln_bwd_fast<6>(a, grid, st);
      if (lrc) return lrc;
      CT_LAUNCH_OK();
      if (use_ws) {
        dim3 g2((unsigned)((cols + 31) / 32), 3);
        CT_CUDA_OK(launch_k(ln_bwd_reduce3_kernel, g2, dim3(256), 0, st, a.workspace, (int)grid, (int)cols, a.dgamma, a.dbeta,
                            a.dxsum, a.dgb_accumulate, a.dxsum_accumulate));
        CT_LAUNCH_OK();
      }
      return 0;
    }
    LnBwdP p;
    p.dy = a.dy; p.dy_dtype = a.dy_dtype; p.dy2 = a.dy2; p.dy2_dtype = a.dy2_dtype;
    p.x = a.x; p.x_dtype = a.x_dtype; p.gamma = a.gamma; p.mean = a.mean; p.rstd = a.rstd;
    p.dx_add = a.dx_add; p.dx_add_dtype = a.dx_add_dtype; p.dx = a.dx; p.dx_dtype = a.dx_dtype;
    p.dx2 = a.dx2; p.dx2_dtype = a.dx2_dtype;
    p.dgamma = a.dgamma; p.dbeta = a.dbeta; p.dxsum = a.dxsum;
    p.partial = use_ws ? a.workspace : nullptr;
    p.rows = rows; p.cols = (int)cols;
    ln_bwd_cta_kernel<<<grid, (int)(cols / 2), 0, st>>>(p);
    CT_LAUNCH_OK();
    if (use_ws) {
      dim3 g2((unsigned)((cols + 31) / 32), 3);
      ln_bwd_reduce3_kernel<<<g2, 256, 0, st>>>(a.workspace, grid, (int)cols, a.dgamma, a.dbeta, a.dxsum,
                                                a.dgb_accumulate, a.dxsum_accumulate);
      CT_LAUNCH_OK();
    }
    return 0;
  }

  // ---- v1 (warp per row) / generic: no fused second output or column sum -> separate passes ----
  int r = ln_bwd_v1(a.dy, a.dy_dtype, a.dy2, a.dy2_dtype, a.x, a.x_dtype, a.gamma, a.mean, a.rstd,
                           a.dx_add, a.dx_add_dtype, a.dx, a.dx_dtype, a.dgamma, a.dbeta, a.dgb_accumulate,
                           a.workspace, a.workspace_bytes, rows, cols, stream);
  if (r) return r;
  if (a.dx2) {
    r = ct_cast(a.dx, a.dx_dtype, a.dx2, a.dx2_dtype, rows * cols, stream);
    if (r) return r;
  }
  if (a.dxsum) return ct_colsum(a.dx, a.dx_dtype, cols, a.dxsum, a.dxsum_accumulate, rows, cols, stream);
  return 0;
}


extern "C" int ct_layernorm_bwd(const void* dy, int dy_dtype, const void* dy2, int dy2_dtype,
                                const void* x, int x_dtype, const float* gamma, const float* mean,
                                const float* rstd, const void* dx_add, int dx_add_dtype, void* dx,
                                int dx_dtype, float* dgamma, float* dbeta, int dgb_accumulate,
                                float* workspace, size_t workspace_bytes, int64_t rows, int64_t cols,
                                void* stream) {
  ct_ln_bwd_args a;
  memset(&a, 0, sizeof(a));
  a.rows = rows; a.cols = cols;
  a.dy = dy; a.dy_dtype = dy_dtype; a.dy2 = dy2; a.dy2_dtype = dy2_dtype;
  a.x = x; a.x_dtype = x_dtype; a.gamma = gamma; a.mean = mean; a.rstd = rstd;
  a.dx_add = dx_add; a.dx_add_dtype = dx_add_dtype; a.dx = dx; a.dx_dtype = dx_dtype;
  a.dgamma = dgamma; a.dbeta = dbeta; a.dgb_accumulate = dgb_accumulate;
  a.workspace = workspace; a.workspace_bytes = workspace_bytes;
  return ct_layernorm_bwd_ex(&a, stream);
}
