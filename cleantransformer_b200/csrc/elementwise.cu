// elementwise.cu — casts (autocast's fp32->bf16 weight/activation casts), bias-gradient column
// sums, and stand-alone activation forward/backward (modeling_bloom.py:335-363 GeLUFunction,
// modeling_gpt.py:112-122 NewGELUActivation, torch.nn.GELU/ReLU used by modeling_bert.py:229 and
// transformer.py:100). All HBM-bound, 128-bit accesses where alignment allows.
#include "ct_common.cuh"
#include "../../include/ct_b200.h"

namespace ct {

__device__ __forceinline__ float ld_any(const void* p, int dt, int64_t i) {
  if (dt == DT_F32) return reinterpret_cast<const float*>(p)[i];
  if (dt == DT_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, int dt, int64_t i, float v) {
  if (dt == DT_F32) reinterpret_cast<float*>(p)[i] = v;
  else if (dt == DT_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
}

// f32 -> bf16, 8 elements per thread per iteration (2x16B loads, 1x16B store)
__global__ void __launch_bounds__(256)
    cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t nv = n >> 3;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float4 a = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i);
    float4 b = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
    reinterpret_cast<uint4*>(dst)[i] = o;
  }
  const int64_t t = (nv << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) dst[t] = __float2bfloat16_rn(src[t]);
}

__global__ void __launch_bounds__(256)
    cast_generic_kernel(const void* __restrict__ src, int sdt, void* __restrict__ dst, int ddt,
                        int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    st_any(dst, ddt, i, ld_any(src, sdt, i));
}

// Column sums: block (32 x 8) handles a 32-column strip over a slab of rows; partials reduced in
// smem then one atomicAdd per column per block.
__global__ void __launch_bounds__(256)
    colsum_kernel(const void* __restrict__ x, int dt, int64_t ld, float* __restrict__ out,
                  int64_t rows, int64_t cols, int64_t rows_per_block) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t col = (int64_t)blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float acc = 0.f;
  if (col < cols)
    for (int64_t r = r0 + ty; r < r1; r += 8) acc += ld_any(x, dt, r * ld + col);
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && col < cols) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][tx];
    atomicAdd(out + col, s);
  }
}

// bf16 fast path: each thread owns 8 consecutive columns (one 16-byte load per row), a block covers
// 256 columns x a slab of rows; 8 row-groups reduced through smem.
__global__ void __launch_bounds__(256)
    colsum_bf16x8_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, float* __restrict__ out,
                         int64_t rows, int64_t cols, int64_t rows_per_block) {
  __shared__ float red[8][256 + 8];
  pdl_wait();               // programmatic dependent launch: see ct_common.cuh
  pdl_launch_dependents();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * 256 + tx * 8;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c0 < cols) {
#pragma unroll 4
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const uint4 u = __ldcs(reinterpret_cast<const uint4*>(x + r * ld + c0));
      float2 f;
      f = unpack_bf16x2(u.x); acc[0] += f.x; acc[1] += f.y;
      f = unpack_bf16x2(u.y); acc[2] += f.x; acc[3] += f.y;
      f = unpack_bf16x2(u.z); acc[4] += f.x; acc[5] += f.y;
      f = unpack_bf16x2(u.w); acc[6] += f.x; acc[7] += f.y;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[ty][tx * 8 + i] = acc[i];
  __syncthreads();
  const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (c < cols) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    atomicAdd(out + c, s);
  }
}

__global__ void __launch_bounds__(256)
    act_fwd_kernel(const void* __restrict__ x, int xdt, void* __restrict__ y, int ydt, int act,
                   int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    st_any(y, ydt, i, act_apply(ld_any(x, xdt, i), act));
}
__global__ void __launch_bounds__(256)
    act_bwd_kernel(const void* __restrict__ dy, int gdt, const void* __restrict__ x, int xdt,
                   void* __restrict__ dx, int ddt, int act, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    st_any(dx, ddt, i, ld_any(dy, gdt, i) * act_grad(ld_any(x, xdt, i), act));
}
// all-bf16 fast paths: 16-byte accesses
__global__ void __launch_bounds__(256)
    act_fwd_bf16x8_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int act, int64_t nvec) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 u = __ldcs(x + i);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    uint4 o;
    o.x = pack_bf16x2(act_apply(a.x, act), act_apply(a.y, act));
    o.y = pack_bf16x2(act_apply(b.x, act), act_apply(b.y, act));
    o.z = pack_bf16x2(act_apply(c.x, act), act_apply(c.y, act));
    o.w = pack_bf16x2(act_apply(d.x, act), act_apply(d.y, act));
    y[i] = o;
  }
}
__global__ void __launch_bounds__(256)
    act_bwd_bf16x8_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, uint4* __restrict__ dx,
                          int act, int64_t nvec) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 g = __ldcs(dy + i), u = __ldcs(x + i);
    float2 ga = unpack_bf16x2(g.x), gb = unpack_bf16x2(g.y), gc = unpack_bf16x2(g.z), gd = unpack_bf16x2(g.w);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    uint4 o;
    o.x = pack_bf16x2(ga.x * act_grad(a.x, act), ga.y * act_grad(a.y, act));
    o.y = pack_bf16x2(gb.x * act_grad(b.x, act), gb.y * act_grad(b.y, act));
    o.z = pack_bf16x2(gc.x * act_grad(c.x, act), gc.y * act_grad(c.y, act));
    o.w = pack_bf16x2(gd.x * act_grad(d.x, act), gd.y * act_grad(d.y, act));
    dx[i] = o;
  }
}

static int ew_blocks(int64_t work) {
  int64_t b = (work + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
// torch.nn.Dropout on a hidden-state tensor, optionally fused with the residual add that follows it
// (include/ct_b200.h: ct_dropout). 4 elements per thread and iteration: the streams are read as consecutive elements.
__global__ void __launch_bounds__(256)
    dropout_kernel(const void* __restrict__ x, int xdt, const void* __restrict__ res, int rdt, void* __restrict__ out,
                   int odt, int64_t n, const DropKey k) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = drop_keep(k, (uint32_t)((uint64_t)i >> 32), (uint32_t)i) ? ld_any(x, xdt, i) * k.rscale : 0.f;
    if (res) v += ld_any(res, rdt, i);
    st_any(out, odt, i, v);
  }
}

static bool dt_any_ok(int dt) { return dt == DT_F32 || dt == DT_BF16 || dt == DT_F16; }

}  // namespace ct

using namespace ct;

extern "C" int ct_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n,
                       void* stream) {
  CT_REQUIRE(src && dst, CT_ERR_BAD_ARG, "ct_cast: null pointer");
  CT_REQUIRE(dt_any_ok(src_dtype) && dt_any_ok(dst_dtype), CT_ERR_UNSUPPORTED, "ct_cast: dtype");
  if (n <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (src_dtype == DT_F32 && dst_dtype == DT_BF16 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0)
    cast_f32_bf16_kernel<<<ew_blocks(n >> 3), 256, 0, st>>>((const float*)src,
                                                            (__nv_bfloat16*)dst, n);
  else
    cast_generic_kernel<<<ew_blocks(n), 256, 0, st>>>(src, src_dtype, dst, dst_dtype, n);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_colsum(const void* x, int x_dtype, int64_t ld, float* out, int accumulate,
                         int64_t rows, int64_t cols, void* stream) {
  CT_REQUIRE(x && out, CT_ERR_BAD_ARG, "ct_colsum: null pointer");
  CT_REQUIRE(dt_any_ok(x_dtype), CT_ERR_UNSUPPORTED, "ct_colsum: dtype");
  CT_REQUIRE(rows >= 0 && cols > 0 && ld >= cols, CT_ERR_BAD_ARG, "ct_colsum: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) CT_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * cols, st));
  if (rows == 0) return 0;
  const int64_t strips = (cols + 31) / 32;
  int64_t slabs = ((int64_t)sm_count() * 8 + strips - 1) / strips;
  if (slabs < 1) slabs = 1;
  int64_t rpb = (rows + slabs - 1) / slabs;
  if (rpb < 64) rpb = 64;
  slabs = (rows + rpb - 1) / rpb;
  if (x_dtype == DT_BF16 && (cols % 8) == 0 && (ld % 8) == 0 && (((uintptr_t)x) & 15) == 0) {
    const int64_t strips8 = (cols + 255) / 256;
    int64_t slabs8 = ((int64_t)sm_count() * 8 + strips8 - 1) / strips8;
    int64_t rpb8 = (rows + slabs8 - 1) / slabs8;
    if (rpb8 < 64) rpb8 = 64;
    slabs8 = (rows + rpb8 - 1) / rpb8;
    dim3 grid8((unsigned)strips8, (unsigned)slabs8);
    CT_CUDA_OK(launch_k(colsum_bf16x8_kernel, grid8, dim3(256), 0, st, (const __nv_bfloat16*)x, ld, out, rows, cols, rpb8));
    CT_LAUNCH_OK();
    return 0;
  }
  dim3 grid((unsigned)strips, (unsigned)slabs);
  colsum_kernel<<<grid, 256, 0, st>>>(x, x_dtype, ld, out, rows, cols, rpb);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_act_fwd(const void* x, int x_dtype, void* y, int y_dtype, int act, int64_t n,
                          void* stream) {
  CT_REQUIRE(x && y, CT_ERR_BAD_ARG, "ct_act_fwd: null pointer");
  CT_REQUIRE(dt_any_ok(x_dtype) && dt_any_ok(y_dtype) && act >= 0 && act <= 4, CT_ERR_UNSUPPORTED,
             "ct_act_fwd: dtype/act");
  if (n <= 0) return 0;
  if (x_dtype == DT_BF16 && y_dtype == DT_BF16 && (n % 8) == 0 && ((((uintptr_t)x) | ((uintptr_t)y)) & 15) == 0)
    act_fwd_bf16x8_kernel<<<ew_blocks(n / 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, act, n / 8);
  else
    act_fwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, y, y_dtype, act, n);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_act_bwd(const void* dy, int dy_dtype, const void* x, int x_dtype, void* dx,
                          int dx_dtype, int act, int64_t n, void* stream) {
  CT_REQUIRE(dy && x && dx, CT_ERR_BAD_ARG, "ct_act_bwd: null pointer");
  CT_REQUIRE(dt_any_ok(dy_dtype) && dt_any_ok(x_dtype) && dt_any_ok(dx_dtype) && act >= 0 &&
                 act <= 4,
             CT_ERR_UNSUPPORTED, "ct_act_bwd: dtype/act");
  if (n <= 0) return 0;
  if (dy_dtype == DT_BF16 && x_dtype == DT_BF16 && dx_dtype == DT_BF16 && (n % 8) == 0 &&
      ((((uintptr_t)dy) | ((uintptr_t)x) | ((uintptr_t)dx)) & 15) == 0)
    act_bwd_bf16x8_kernel<<<ew_blocks(n / 8), 256, 0, (cudaStream_t)stream>>>((const uint4*)dy, (const uint4*)x,
                                                                              (uint4*)dx, act, n / 8);
  else
    act_bwd_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(dy, dy_dtype, x, x_dtype, dx,
                                                                   dx_dtype, act, n);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_dropout(const void* x, int x_dtype, const void* residual, int res_dtype, void* out, int out_dtype,
                          int64_t n, float p, uint64_t seed, uint32_t rng_stream, void* stream) {
  CT_REQUIRE(x && out, CT_ERR_BAD_ARG, "ct_dropout: null pointer");
  CT_REQUIRE(dt_any_ok(x_dtype) && dt_any_ok(out_dtype) && (!residual || dt_any_ok(res_dtype)), CT_ERR_UNSUPPORTED,
             "ct_dropout: dtype");
  CT_REQUIRE(p >= 0.f && p < 1.f, CT_ERR_BAD_ARG, "ct_dropout: p must be in [0, 1)");
  if (n <= 0) return 0;
  dropout_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(x, x_dtype, residual, res_dtype, out, out_dtype, n,
                                                                 make_drop_key(p, seed, rng_stream));
  CT_LAUNCH_OK();
  return 0;
}
