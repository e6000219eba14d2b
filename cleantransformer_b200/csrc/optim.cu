// optim.cu — AdamW / SGD as single vectorised HBM-streaming kernels.
//
// Replaces ~12 eager elementwise ops x 293 tensors of CleanTransformer/optimizer.py:71-97
// (AdamW.step) and the torch.optim.AdamW call of examples/ft_bloom.py:70,90.
// Algorithmic bytes: 28 B/param (read p,g,m,v; write p,m,v) (+2 B/param with the bf16 shadow,
// +4 B/param in mode 1 with weight decay, where the reference also rewrites g).
#include "ct_common.cuh"
#include "../../include/ct_b200.h"
#include <cmath>

namespace ct {

struct AdamConsts {
  float lr, beta1, beta2, eps, wd;
  float omb1, omb2;     // 1 - beta, rounded from the double-precision difference (as Python/torch do)
  float decay;          // 1 - lr*wd
  float bc1, bc2;       // 1 - beta^t
  float rsqrt_bc2;      // 1/sqrt(bc2)   (torch path)
  float step_size;      // lr / bc1       (torch path)
  float grad_scale;
  int mode;
};

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v,
                                         const AdamConsts& c) {
  g *= c.grad_scale;
  if (c.mode == 0) {
    // torch.optim.AdamW (decoupled): p *= 1 - lr*wd; m = lerp(m, g, 1-b1); v = b2 v + (1-b2) g^2;
    // denom = sqrt(v)/sqrt(bc2) + eps; p -= (lr/bc1) * m / denom
    p *= c.decay;
    m = m + (g - m) * c.omb1;
    v = c.beta2 * v + c.omb2 * g * g;
    const float denom = sqrtf(v) * c.rsqrt_bc2 + c.eps;
    p -= c.step_size * (m / denom);
  } else {
    // optimizer.py:80-95 — coupled L2, bias-corrected moments, eps outside the sqrt
    if (c.wd != 0.f) g += c.wd * p;
    m = c.beta1 * m + c.omb1 * g;
    v = c.beta2 * v + c.omb2 * g * g;
    const float mh = m / c.bc1;
    const float vh = v / c.bc2;
    p -= c.lr * mh / (sqrtf(vh) + c.eps);
  }
}

__global__ void __launch_bounds__(256)
    adamw_flat_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                      float* __restrict__ v, __nv_bfloat16* __restrict__ shadow, int64_t n,
                      AdamConsts c, int write_g) {
  const int64_t nvec = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mm = __ldcs(reinterpret_cast<const float4*>(m) + i);
    float4 vv = __ldcs(reinterpret_cast<const float4*>(v) + i);
    adam_one(pp.x, gg.x, mm.x, vv.x, c);
    adam_one(pp.y, gg.y, mm.y, vv.y, c);
    adam_one(pp.z, gg.z, mm.z, vv.z, c);
    adam_one(pp.w, gg.w, mm.w, vv.w, c);
    reinterpret_cast<float4*>(p)[i] = pp;
    __stcs(reinterpret_cast<float4*>(m) + i, mm);
    __stcs(reinterpret_cast<float4*>(v) + i, vv);
    if (write_g) __stcs(reinterpret_cast<float4*>(g) + i, gg);
    if (shadow) {
      uint2 u;
      u.x = pack_bf16x2(pp.x, pp.y);
      u.y = pack_bf16x2(pp.z, pp.w);
      reinterpret_cast<uint2*>(shadow)[i] = u;
    }
  }
  // tail (n % 4)
  const int64_t t = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    float pp = p[t], gg = g[t], mm = m[t], vv = v[t];
    adam_one(pp, gg, mm, vv, c);
    p[t] = pp; m[t] = mm; v[t] = vv;
    if (write_g) g[t] = gg;
    if (shadow) shadow[t] = __float2bfloat16_rn(pp);
  }
}

constexpr int MT_MAX = 40;
struct MultiTable {
  float* p[MT_MAX];
  float* g[MT_MAX];
  float* m[MT_MAX];
  float* v[MT_MAX];
  __nv_bfloat16* s[MT_MAX];
  int64_t n[MT_MAX];
  int count;
};

// blockIdx.y selects the tensor; scalar accesses keep arbitrary (unaligned) views legal.
__global__ void __launch_bounds__(256)
    adamw_multi_kernel(const __grid_constant__ MultiTable t, AdamConsts c, int write_g) {
  const int k = blockIdx.y;
  float* __restrict__ p = t.p[k];
  float* __restrict__ g = t.g[k];
  float* __restrict__ m = t.m[k];
  float* __restrict__ v = t.v[k];
  __nv_bfloat16* __restrict__ s = t.s[k];
  const int64_t n = t.n[k];
  const bool vec_ok = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0) &&
                      (s == nullptr || ((uintptr_t)s & 7) == 0);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t start = 0;
  if (vec_ok) {
    const int64_t nvec = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
      float4 pp = reinterpret_cast<float4*>(p)[i];
      float4 gg = reinterpret_cast<const float4*>(g)[i];
      float4 mm = reinterpret_cast<const float4*>(m)[i];
      float4 vv = reinterpret_cast<const float4*>(v)[i];
      adam_one(pp.x, gg.x, mm.x, vv.x, c);
      adam_one(pp.y, gg.y, mm.y, vv.y, c);
      adam_one(pp.z, gg.z, mm.z, vv.z, c);
      adam_one(pp.w, gg.w, mm.w, vv.w, c);
      reinterpret_cast<float4*>(p)[i] = pp;
      reinterpret_cast<float4*>(m)[i] = mm;
      reinterpret_cast<float4*>(v)[i] = vv;
      if (write_g) reinterpret_cast<float4*>(g)[i] = gg;
      if (s) {
        uint2 u;
        u.x = pack_bf16x2(pp.x, pp.y);
        u.y = pack_bf16x2(pp.z, pp.w);
        reinterpret_cast<uint2*>(s)[i] = u;
      }
    }
    start = nvec << 2;
  }
  for (int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    adam_one(pp, gg, mm, vv, c);
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (write_g) g[i] = gg;
    if (s) s[i] = __float2bfloat16_rn(pp);
  }
}

__global__ void __launch_bounds__(256)
    sgd_flat_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ buf,
                    int64_t n, float lr, float momentum, float dampening, float wd, int first) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float pp = p[i], gg = g[i];
    if (wd != 0.f) gg += wd * pp;                                   // optimizer.py:37-38
    if (momentum != 0.f) {                                           // optimizer.py:40-47
      float b = first ? gg : fmaf(momentum, buf[i], (1.f - dampening) * gg);
      buf[i] = b;
      gg = b;
    }
    g[i] = gg;
    p[i] = pp - lr * gg;                                             // optimizer.py:48
  }
}

// Hyper-parameters arrive as doubles (Python floats) and every derived constant is formed in double
// before rounding to fp32 once, which is what `(1 - beta) * tensor` does in the reference / torch.
static AdamConsts make_consts(double lr, double b1, double b2, double eps, double wd, int64_t step,
                              int mode, float grad_scale) {
  AdamConsts c;
  c.lr = (float)lr; c.beta1 = (float)b1; c.beta2 = (float)b2; c.eps = (float)eps; c.wd = (float)wd;
  c.omb1 = (float)(1.0 - b1); c.omb2 = (float)(1.0 - b2);
  c.decay = (float)(1.0 - lr * wd);
  const double bc1 = 1.0 - std::pow(b1, (double)step);
  const double bc2 = 1.0 - std::pow(b2, (double)step);
  c.bc1 = (float)bc1;
  c.bc2 = (float)bc2;
  c.rsqrt_bc2 = (float)(1.0 / std::sqrt(bc2));
  c.step_size = (float)(lr / bc1);
  c.grad_scale = grad_scale;
  c.mode = mode;
  return c;
}

}  // namespace ct

using namespace ct;

extern "C" int ct_adamw_step(float* p, float* g, float* m, float* v, void* p_shadow_bf16,
                             int64_t n, double lr, double beta1, double beta2, double eps,
                             double weight_decay, int64_t step, int mode, float grad_scale,
                             void* stream) {
  CT_REQUIRE(p && g && m && v, CT_ERR_BAD_ARG, "ct_adamw_step: null pointer");
  CT_REQUIRE(n >= 0 && step >= 1 && (mode == 0 || mode == 1), CT_ERR_BAD_ARG,
             "ct_adamw_step: bad n/step/mode");
  CT_REQUIRE(((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0) &&
                 (((uintptr_t)p_shadow_bf16 & 7) == 0),
             CT_ERR_BAD_ARG, "ct_adamw_step: arena pointers must be 16-byte aligned");
  if (n == 0) return 0;
  AdamConsts c = make_consts(lr, beta1, beta2, eps, weight_decay, step, mode, grad_scale);
  const int write_g = (mode == 1 && weight_decay != 0.0) ? 1 : 0;
  int64_t blocks = ((n >> 2) + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adamw_flat_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      p, g, m, v, (__nv_bfloat16*)p_shadow_bf16, n, c, write_g);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_adamw_multi(int ntensors, float* const* p, float* const* g, float* const* m,
                              float* const* v, void* const* p_shadow_bf16, const int64_t* sizes,
                              double lr, double beta1, double beta2, double eps, double weight_decay,
                              int64_t step, int mode, float grad_scale, void* stream) {
  CT_REQUIRE(ntensors >= 0 && p && g && m && v && sizes, CT_ERR_BAD_ARG,
             "ct_adamw_multi: null pointer");
  CT_REQUIRE(step >= 1 && (mode == 0 || mode == 1), CT_ERR_BAD_ARG, "ct_adamw_multi: bad step/mode");
  AdamConsts c = make_consts(lr, beta1, beta2, eps, weight_decay, step, mode, grad_scale);
  const int write_g = (mode == 1 && weight_decay != 0.0) ? 1 : 0;
  int done = 0;
  while (done < ntensors) {
    MultiTable t;
    t.count = 0;
    int64_t maxn = 0;
    while (done < ntensors && t.count < MT_MAX) {
      if (sizes[done] > 0) {
        CT_REQUIRE(p[done] && g[done] && m[done] && v[done], CT_ERR_BAD_ARG,
                   "ct_adamw_multi: null tensor %d", done);
        const int k = t.count++;
        t.p[k] = p[done]; t.g[k] = g[done]; t.m[k] = m[done]; t.v[k] = v[done];
        t.s[k] = p_shadow_bf16 ? (__nv_bfloat16*)p_shadow_bf16[done] : nullptr;
        t.n[k] = sizes[done];
        if (sizes[done] > maxn) maxn = sizes[done];
      }
      ++done;
    }
    if (t.count == 0) break;
    int64_t bx = ((maxn >> 2) + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 4;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)t.count);
    adamw_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(t, c, write_g);
    CT_LAUNCH_OK();
  }
  return 0;
}

extern "C" int ct_sgd_step(float* p, float* g, float* buf, int64_t n, float lr, float momentum,
                           float dampening, float weight_decay, int first_step, void* stream) {
  CT_REQUIRE(p && g && (buf || momentum == 0.f), CT_ERR_BAD_ARG, "ct_sgd_step: null pointer");
  if (n <= 0) return 0;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  sgd_flat_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, buf, n, lr, momentum,
                                                                 dampening, weight_decay, first_step);
  CT_LAUNCH_OK();
  return 0;
}

// Same arithmetic over `ntensors` separate tensors (host arrays of device pointers; `buf` entries may be NULL when
// momentum == 0): one launch of the flat kernel per tensor.
extern "C" int ct_sgd_multi(int ntensors, float* const* p, float* const* g, float* const* buf, const int64_t* sizes,
                            float lr, float momentum, float dampening, float weight_decay, int first_step,
                            void* stream) {
  CT_REQUIRE(ntensors >= 0 && (ntensors == 0 || (p && g && sizes)), CT_ERR_BAD_ARG, "ct_sgd_multi: null table");
  for (int i = 0; i < ntensors; ++i) {
    const int rc = ct_sgd_step(p[i], g[i], buf ? buf[i] : nullptr, sizes[i], lr, momentum, dampening, weight_decay,
                               first_step, stream);
    if (rc) return rc;
  }
  return 0;
}
