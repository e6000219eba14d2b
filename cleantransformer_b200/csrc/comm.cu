// comm.cu — gradient-bucket all-reduce as plain CUDA kernels over NVLink/NVSwitch peer memory.
//
// Replaces the ncclAllReduce calls that torch's DDP reducer issues for the reference
// (examples/ft_bloom_DDP.py:99,126,135 `DDP(model, device_ids=[local_rank])`; README.md:46-52
// describes the hand-rolled version: param sync, gradient buckets, overlapped reduction).
//
// Memory: every rank cudaMalloc()s one symmetric gradient buffer + one small signal buffer and
// publishes cudaIpcMemHandles; Python (torch.distributed, used only as bootstrap) exchanges the 64-byte
// handles and each rank maps every peer buffer (cudaIpcOpenMemHandle). The gradient arena
// (arena.py) lives inside the symmetric buffer, so wgrad kernels write straight into NVLink-visible
// memory and the reduction is in place.
//
// ct_allreduce(offset, count): two-shot, one kernel per rank
//   phase 0  signal "my data for epoch e is ready" to every peer, wait for all peers
//   phase 1  rank r owns slice r of the range: 128-bit loads of that slice from all W ranks
//            (peer reads over NVLink), fp32 sum in rank order (deterministic), * scale, 128-bit
//            stores of the result to all W ranks (peer writes)
//   phase 2  last CTA signals "my slice is written everywhere", waits for all peers' signals
// One-shot variant (each rank reads everything from every peer, writes only locally) for small,
// latency-bound buckets. Bytes over NVLink per rank: two-shot 2*(W-1)/W * bytes, one-shot
// (W-1) * bytes inbound.
#include "ct_common.cuh"
#include "../../include/ct_b200.h"
#include <cstdlib>
#include <mutex>
#include <vector>

namespace ct {

constexpr int MAX_WORLD = 16;

struct CommCtx {
  int rank = -1, world = 0, device = 0;
  bool ready = false;
  float* data[MAX_WORLD] = {nullptr};        // symmetric gradient buffers (local at [rank])
  uint32_t* sig[MAX_WORLD] = {nullptr};      // signal buffers: [0..W) ready flags, [W..2W) done flags,
                                             // [2W] CTA counter (local use only)
  size_t data_bytes = 0;
  uint32_t epoch = 0;
  std::vector<void*> opened;
};
static CommCtx g_comm;
static std::mutex g_comm_mu;

struct ArParams {
  float* data[MAX_WORLD];
  uint32_t* sig[MAX_WORLD];
  int rank, world;
  uint32_t epoch;
  int64_t offset, count;  // elements (count % 4 == 0, offset % 4 == 0)
  float scale;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_peer_f4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// Upper bound on one peer-flag wait, in SM clock cycles; 0 = wait for ever (what NCCL does). Set once by
// ct_comm_init from CT_COMM_TIMEOUT_S (default 1800 s: rank skew from a slow checkpoint write, a stalled data
// loader or a debugger must not kill the job; a peer that is really gone still ends in a trap, not a hung node).
__device__ long long g_wait_timeout_cycles = 0;

__device__ __forceinline__ void wait_flags(const uint32_t* flags, int world, uint32_t epoch) {
  // threads 0..world-1 poll one peer flag each
  if ((int)threadIdx.x < world) {
    const long long limit = g_wait_timeout_cycles;
    long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
      if (limit > 0 && clock64() - t0 > limit) {
        printf("ct_b200 comm: peer %d did not reach epoch %u within CT_COMM_TIMEOUT_S\n", threadIdx.x, epoch);
        __trap();
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
}

constexpr int AR_THREADS = 256;

template <bool ONE_SHOT>
__global__ void __launch_bounds__(AR_THREADS, 6)  // <= 40 registers: fits beside a resident GEMM CTA
    allreduce_kernel(const ArParams p) {
  const int W = p.world, r = p.rank;
  uint32_t* my_sig = p.sig[r];
  // ---- phase 0: publish readiness, wait for everyone ----
  if (blockIdx.x == 0 && (int)threadIdx.x < W) st_release_sys(p.sig[threadIdx.x] + r, p.epoch);
  wait_flags(my_sig, W, p.epoch);

  const int64_t nvec = p.count >> 2;
  if (ONE_SHOT) {
    // every rank reduces the whole range into its own buffer
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 mine;
      for (int q = 0; q < W; ++q) {
        const float4 v = ld_peer_f4(p.data[q] + p.offset + 4 * i);
        if (q == r) mine = v;
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      (void)mine;
      acc.x *= p.scale; acc.y *= p.scale; acc.z *= p.scale; acc.w *= p.scale;
      // results are staged: nobody may overwrite its input while peers still read it
      // -> written after phase 2's barrier below (kept in registers is impossible for big ranges),
      // so one-shot writes to a shadow half of the range instead: see host side (count doubled)
      st_peer_f4(p.data[r] + p.offset + p.count + 4 * i, acc);
    }
  } else {
    // slice owned by this rank (multiple of 4 elements). Four independent 16-byte vectors per
    // thread per iteration: W x 4 peer loads in flight per thread (NVLink latency ~2 us needs
    // ~1.5 MB in flight per direction), small CTAs / few registers so they co-reside with the
    // persistent GEMM CTAs of the backward pass.
    const int64_t per = ((nvec + W - 1) / W);
    const int64_t v0 = min(nvec, per * r), v1 = min(nvec, per * (r + 1));
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = v0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v1; i += 4 * stride) {
      float4 acc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
      for (int q = 0; q < W; ++q) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t idx = i + u * stride;
          v[u] = idx < v1 ? ld_peer_f4(p.data[q] + p.offset + 4 * idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[u].x *= p.scale; acc[u].y *= p.scale; acc[u].z *= p.scale; acc[u].w *= p.scale;
      }
      for (int q = 0; q < W; ++q) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t idx = i + u * stride;
          if (idx < v1) st_peer_f4(p.data[q] + p.offset + 4 * idx, acc[u]);
        }
      }
    }
  }
  // ---- phase 2: last CTA publishes completion and waits for the peers' completion ----
  __threadfence_system();
  __syncthreads();
  __shared__ int last;
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(my_sig + 2 * MAX_WORLD, 1u);
    last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) my_sig[2 * MAX_WORLD] = 0;  // reset for the next call (stream-ordered)
  if ((int)threadIdx.x < W) st_release_sys(p.sig[threadIdx.x] + MAX_WORLD + r, p.epoch);
  wait_flags(my_sig + MAX_WORLD, W, p.epoch);
}

// one-shot epilogue: copy the staged result back over the input (local, after the barrier)
__global__ void __launch_bounds__(256)
    copy_f4_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t nvec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
}

// broadcast: every rank copies [offset, offset+count) from the root's buffer (after a barrier)
__global__ void __launch_bounds__(512)
    broadcast_kernel(const ArParams p, int root) {
  const int W = p.world, r = p.rank;
  uint32_t* my_sig = p.sig[r];
  if (blockIdx.x == 0 && (int)threadIdx.x < W) st_release_sys(p.sig[threadIdx.x] + r, p.epoch);
  wait_flags(my_sig, W, p.epoch);
  if (r != root) {
    const int64_t nvec = p.count >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x)
      st_peer_f4(p.data[r] + p.offset + 4 * i, ld_peer_f4(p.data[root] + p.offset + 4 * i));
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int last;
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(my_sig + 2 * MAX_WORLD, 1u);
    last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) my_sig[2 * MAX_WORLD] = 0;
  if ((int)threadIdx.x < W) st_release_sys(p.sig[threadIdx.x] + MAX_WORLD + r, p.epoch);
  wait_flags(my_sig + MAX_WORLD, W, p.epoch);
}

// ---------------------------------------------------------------------------------------------------
// Tied-embedding gradient, sparse half. The LM-head wgrad of a tied table is dense and ready at the START of
// backward; the embedding scatter-add arrives at the very END. Reducing the table once both are there
// (what a bucketed reducer does) leaves a ~1 GB all-reduce with nothing to hide behind. Instead the dense
// half is all-reduced as soon as the LM-head wgrad is enqueued, and the sparse half is exchanged as what it
// is: every rank publishes its T x H token gradients + token ids in the symmetric buffer, and every rank
// scatter-adds ALL ranks' rows (scaled by 1/W) into its own, already averaged, table gradient:
// (W-1) * T * H * 4 bytes inbound per rank (235 MB at W = 8) instead of 2 (W-1)/W x 1 GB.
struct EsParams {
  float* data[MAX_WORLD];
  uint32_t* sig[MAX_WORLD];
  int rank, world;
  uint32_t epoch;
  int64_t hdr_off;    // float offset of the staging header: int64 T (tokens this rank staged)
  int64_t rows_off;   // float offset of the staged rows [cap, H] f32
  int64_t ids_off;    // float offset of the staged ids [cap] int64
  int64_t grad_off;   // float offset of the table gradient [V, H] in the LOCAL buffer
  int64_t H, V;
  long long padding_idx;
  float scale;
};

__device__ __forceinline__ long long ld_peer_s64(const long long* p) {
  long long v;
  asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
    embed_scatter_allranks_kernel(const EsParams p) {
  const int W = p.world, r = p.rank;
  uint32_t* my_sig = p.sig[r];
  if (blockIdx.x == 0 && (int)threadIdx.x < W) st_release_sys(p.sig[threadIdx.x] + r, p.epoch);
  wait_flags(my_sig, W, p.epoch);

  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float* grad = p.data[r] + p.grad_off;
  for (int k = 0; k < W; ++k) {
    const int q = (r + k) % W;  // own rows first, then the peers round-robin (spreads the NVLink load)
    const long long Tq = ld_peer_s64(reinterpret_cast<const long long*>(p.data[q] + p.hdr_off));
    const long long* ids = reinterpret_cast<const long long*>(p.data[q] + p.ids_off);
    const float* rows = p.data[q] + p.rows_off;
    for (int64_t t = warp0; t < Tq; t += nwarps) {
      const long long id = ld_peer_s64(ids + t);
      if (id < 0 || id >= p.V || id == p.padding_idx) continue;
      const float* src = rows + t * p.H;
      float* dst = grad + id * p.H;
      for (int64_t c = lane * 4; c < p.H; c += 128) {
        const float4 v = ld_peer_f4(src + c);
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(v.x * p.scale),
                     "f"(v.y * p.scale), "f"(v.z * p.scale), "f"(v.w * p.scale)
                     : "memory");
      }
    }
  }
  // completion barrier: nobody restages while a peer still reads
  __threadfence_system();
  __syncthreads();
  __shared__ int last;
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(my_sig + 2 * MAX_WORLD, 1u);
    last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) my_sig[2 * MAX_WORLD] = 0;
  if ((int)threadIdx.x < W) st_release_sys(p.sig[threadIdx.x] + MAX_WORLD + r, p.epoch);
  wait_flags(my_sig + MAX_WORLD, W, p.epoch);
}

// ---------------------------------------------------------------------------------------------------
// Copy-engine transport (comm mode "ce"). r01i showed what the register-staged kernel above costs when it shares
// SMs with the persistent GEMMs of the backward pass: 48 CTAs hide the exchange but slow the GEMMs by 3.3 ms per
// step, fewer CTAs leave it exposed. The alternative keeps the SMs out of the transport altogether: the two
// NVLink legs of the two-shot all-reduce are cudaMemcpyAsync peer copies (DMA engines), and the only kernels are a
// one-CTA flag barrier and an HBM-bound local reduction of the W staged slices:
//   barrier | pull slice r of every peer into local staging | reduce (rank order) | push the result into every
//   peer's slice r | barrier
// Every slice is reduced by exactly one rank and then copied, so all ranks end up bit-identical.
__global__ void __launch_bounds__(32) comm_barrier_kernel(const ArParams p) {
  const int W = p.world, r = p.rank;
  if ((int)threadIdx.x < W) st_release_sys(p.sig[threadIdx.x] + r, p.epoch);
  wait_flags(p.sig[r], W, p.epoch);
}

// dst[i] = scale * sum over ranks q in rank order of (q == rank ? dst[i] : staged[slot(q)][i]); slots are the peers
// in ascending rank order, `stride` floats apart
__global__ void __launch_bounds__(256)
    reduce_slices_kernel(float* __restrict__ dst, const float* __restrict__ staged, int64_t stride, int world,
                         int rank, int64_t nvec, float scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int slot = 0;
    for (int q = 0; q < world; ++q) {
      const float4 v = q == rank ? reinterpret_cast<const float4*>(dst)[i]
                                 : __ldcs(reinterpret_cast<const float4*>(staged + (int64_t)(slot++) * stride) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
    reinterpret_cast<float4*>(dst)[i] = acc;
  }
}

static void fill_params(ArParams& p, int64_t offset, int64_t count, float scale) {
  for (int q = 0; q < g_comm.world; ++q) { p.data[q] = g_comm.data[q]; p.sig[q] = g_comm.sig[q]; }
  p.rank = g_comm.rank; p.world = g_comm.world;
  p.epoch = ++g_comm.epoch;
  p.offset = offset; p.count = count; p.scale = scale;
}

}  // namespace ct

using namespace ct;

extern "C" int ct_comm_init(int rank, int world, int device, size_t data_bytes, void** local_data,
                            void* data_handle_out, void* sig_handle_out) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  CT_REQUIRE(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world, CT_ERR_BAD_ARG,
             "ct_comm_init: bad rank/world %d/%d", rank, world);
  CT_REQUIRE(!g_comm.ready && g_comm.rank < 0, CT_ERR_COMM, "ct_comm_init: already initialised");
  CT_REQUIRE(local_data && data_handle_out && sig_handle_out, CT_ERR_BAD_ARG, "ct_comm_init: null out");
  CT_CUDA_OK(cudaSetDevice(device));
  data_bytes = (data_bytes + 255) & ~(size_t)255;
  float* d = nullptr;
  uint32_t* s = nullptr;
  CT_CUDA_OK(cudaMalloc(&d, data_bytes));
  CT_CUDA_OK(cudaMalloc(&s, sizeof(uint32_t) * (2 * MAX_WORLD + 8)));
  CT_CUDA_OK(cudaMemset(d, 0, data_bytes));
  CT_CUDA_OK(cudaMemset(s, 0, sizeof(uint32_t) * (2 * MAX_WORLD + 8)));
  CT_CUDA_OK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)data_handle_out, d));
  CT_CUDA_OK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)sig_handle_out, s));
  {
    const char* e = getenv("CT_COMM_TIMEOUT_S");
    const double secs = e ? atof(e) : 1800.0;
    int khz = 2000000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    const long long cycles = secs > 0 ? (long long)(secs * 1e3 * (double)khz) : 0;
    CT_CUDA_OK(cudaMemcpyToSymbol(g_wait_timeout_cycles, &cycles, sizeof(cycles)));
  }
  g_comm.rank = rank; g_comm.world = world; g_comm.device = device;
  g_comm.data[rank] = d; g_comm.sig[rank] = s;
  g_comm.data_bytes = data_bytes;
  g_comm.epoch = 0;
  *local_data = d;
  return 0;
}

extern "C" int ct_comm_connect(const void* data_handles, const void* sig_handles) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  CT_REQUIRE(g_comm.rank >= 0, CT_ERR_COMM, "ct_comm_connect: ct_comm_init first");
  CT_REQUIRE(data_handles && sig_handles, CT_ERR_BAD_ARG, "ct_comm_connect: null handles");
  const cudaIpcMemHandle_t* dh = (const cudaIpcMemHandle_t*)data_handles;
  const cudaIpcMemHandle_t* sh = (const cudaIpcMemHandle_t*)sig_handles;
  for (int q = 0; q < g_comm.world; ++q) {
    if (q == g_comm.rank) continue;
    void* pd = nullptr;
    void* ps = nullptr;
    CT_CUDA_OK(cudaIpcOpenMemHandle(&pd, dh[q], cudaIpcMemLazyEnablePeerAccess));
    CT_CUDA_OK(cudaIpcOpenMemHandle(&ps, sh[q], cudaIpcMemLazyEnablePeerAccess));
    g_comm.data[q] = (float*)pd;
    g_comm.sig[q] = (uint32_t*)ps;
    g_comm.opened.push_back(pd);
    g_comm.opened.push_back(ps);
  }
  g_comm.ready = true;
  return 0;
}

extern "C" int ct_allreduce_bucket(int64_t offset, int64_t count, float scale, int mode, int max_ctas,
                                   void* stream) {
  CT_REQUIRE(g_comm.ready, CT_ERR_COMM, "ct_allreduce_bucket: comm not initialised");
  CT_REQUIRE(offset >= 0 && count >= 0 && (offset % 4) == 0 && (count % 4) == 0 &&
                 (size_t)(offset + count) * 4 <= g_comm.data_bytes,
             CT_ERR_BAD_ARG, "ct_allreduce_bucket: range [%lld,+%lld) must be 16-byte aligned and inside the buffer",
             (long long)offset, (long long)count);
  CT_REQUIRE(mode == 0 || mode == 1, CT_ERR_BAD_ARG, "ct_allreduce_bucket: mode 0 (auto/two-shot) or 1 (one-shot)");
  if (count == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  ArParams p;
  {
    std::lock_guard<std::mutex> lk(g_comm_mu);
    fill_params(p, offset, count, scale);
  }
  if (max_ctas <= 0) max_ctas = 64;
  {
    // same shared-memory carve-out as the big GEMM / attention kernels so that an all-reduce CTA can
    // be co-resident with them (an SM cannot host CTAs that want different L1/shared splits)
    static bool carve = false;
    if (!carve) {
      cudaFuncSetAttribute(allreduce_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
      cudaFuncSetAttribute(allreduce_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
      carve = true;
    }
  }
  const int64_t nvec = count >> 2;
  if (mode == 1) {
    // one-shot needs a staging area of `count` floats right after the range
    CT_REQUIRE((size_t)(offset + 2 * count) * 4 <= g_comm.data_bytes, CT_ERR_WORKSPACE,
               "ct_allreduce_bucket: one-shot needs a staging area after the range");
    int64_t ctas = (nvec + AR_THREADS - 1) / AR_THREADS;
    if (ctas > max_ctas) ctas = max_ctas;
    allreduce_kernel<true><<<(unsigned)ctas, AR_THREADS, 0, st>>>(p);
    CT_LAUNCH_OK();
    copy_f4_kernel<<<(unsigned)ctas, 256, 0, st>>>(g_comm.data[g_comm.rank] + offset,
                                                    g_comm.data[g_comm.rank] + offset + count, nvec);
    CT_LAUNCH_OK();
    return 0;
  }
  int64_t per = (nvec + g_comm.world - 1) / g_comm.world;
  int64_t ctas = (per + 4 * AR_THREADS - 1) / (4 * AR_THREADS);
  if (ctas > max_ctas) ctas = max_ctas;
  if (ctas < 1) ctas = 1;
  allreduce_kernel<false><<<(unsigned)ctas, AR_THREADS, 0, st>>>(p);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_embedding_bwd_allranks(int64_t hdr_offset, int64_t rows_offset, int64_t ids_offset,
                                         int64_t grad_offset, int64_t H, int64_t V, int64_t padding_idx,
                                         float scale, int max_ctas, void* stream) {
  CT_REQUIRE(g_comm.ready, CT_ERR_COMM, "ct_embedding_bwd_allranks: comm not initialised");
  const size_t nfl = g_comm.data_bytes / 4;
  CT_REQUIRE(H > 0 && (H % 4) == 0 && V > 0 && hdr_offset >= 0 && rows_offset >= 0 && ids_offset >= 0 &&
                 grad_offset >= 0 && (hdr_offset % 2) == 0 && (rows_offset % 4) == 0 && (ids_offset % 2) == 0 &&
                 (grad_offset % 4) == 0 && (size_t)hdr_offset + 2 <= nfl && (size_t)rows_offset <= nfl &&
                 (size_t)ids_offset <= nfl && (size_t)(grad_offset + V * H) <= nfl,
             CT_ERR_BAD_ARG, "ct_embedding_bwd_allranks: offsets must be aligned and inside the symmetric buffer");
  EsParams p;
  {
    std::lock_guard<std::mutex> lk(g_comm_mu);
    for (int q = 0; q < g_comm.world; ++q) { p.data[q] = g_comm.data[q]; p.sig[q] = g_comm.sig[q]; }
    p.rank = g_comm.rank; p.world = g_comm.world;
    p.epoch = ++g_comm.epoch;
  }
  p.hdr_off = hdr_offset; p.rows_off = rows_offset; p.ids_off = ids_offset; p.grad_off = grad_offset;
  p.H = H; p.V = V; p.padding_idx = (long long)padding_idx; p.scale = scale;
  if (max_ctas <= 0) max_ctas = sm_count() * 2;
  embed_scatter_allranks_kernel<<<(unsigned)max_ctas, 256, 0, (cudaStream_t)stream>>>(p);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_comm_barrier(void* stream) {
  CT_REQUIRE(g_comm.ready, CT_ERR_COMM, "ct_comm_barrier: comm not initialised");
  ArParams p;
  {
    std::lock_guard<std::mutex> lk(g_comm_mu);
    fill_params(p, 0, 0, 1.f);
  }
  comm_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p);
  CT_LAUNCH_OK();
  return 0;
}

// copy-engine legs: `count` floats between the symmetric buffer of `peer` (offset in floats) and local memory
extern "C" int ct_comm_pull(int peer, int64_t peer_offset, void* dst_local, int64_t count, void* stream) {
  CT_REQUIRE(g_comm.ready, CT_ERR_COMM, "ct_comm_pull: comm not initialised");
  CT_REQUIRE(peer >= 0 && peer < g_comm.world && dst_local && peer_offset >= 0 && count >= 0 &&
                 (size_t)(peer_offset + count) * 4 <= g_comm.data_bytes,
             CT_ERR_BAD_ARG, "ct_comm_pull: bad peer / range");
  if (count == 0) return 0;
  CT_CUDA_OK(cudaMemcpyAsync(dst_local, g_comm.data[peer] + peer_offset, (size_t)count * 4, cudaMemcpyDeviceToDevice,
                             (cudaStream_t)stream));
  return 0;
}

extern "C" int ct_comm_push(int peer, int64_t peer_offset, int64_t local_offset, int64_t count, void* stream) {
  CT_REQUIRE(g_comm.ready, CT_ERR_COMM, "ct_comm_push: comm not initialised");
  CT_REQUIRE(peer >= 0 && peer < g_comm.world && peer_offset >= 0 && local_offset >= 0 && count >= 0 &&
                 (size_t)(peer_offset + count) * 4 <= g_comm.data_bytes &&
                 (size_t)(local_offset + count) * 4 <= g_comm.data_bytes,
             CT_ERR_BAD_ARG, "ct_comm_push: bad peer / range");
  if (count == 0) return 0;
  CT_CUDA_OK(cudaMemcpyAsync(g_comm.data[peer] + peer_offset, g_comm.data[g_comm.rank] + local_offset, (size_t)count * 4,
                             cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

// local_offset: the slice this rank owns (inside its symmetric buffer); staged: (world-1) slices, `stride` floats
// apart, peers in ascending rank order. count % 4 == 0, everything 16-byte aligned.
extern "C" int ct_comm_reduce_slices(int64_t local_offset, const float* staged, int64_t stride, int64_t count,
                                     float scale, int max_ctas, void* stream) {
  CT_REQUIRE(g_comm.ready, CT_ERR_COMM, "ct_comm_reduce_slices: comm not initialised");
  CT_REQUIRE(staged && local_offset >= 0 && count >= 0 && (count % 4) == 0 && (local_offset % 4) == 0 &&
                 (stride % 4) == 0 && stride >= count && ((uintptr_t)staged & 15) == 0 &&
                 (size_t)(local_offset + count) * 4 <= g_comm.data_bytes,
             CT_ERR_BAD_ARG, "ct_comm_reduce_slices: bad range / alignment");
  if (count == 0) return 0;
  const int64_t nvec = count >> 2;
  int64_t ctas = (nvec + 255) / 256;
  if (max_ctas <= 0) max_ctas = sm_count() * 4;
  if (ctas > max_ctas) ctas = max_ctas;
  reduce_slices_kernel<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(
      g_comm.data[g_comm.rank] + local_offset, staged, stride, g_comm.world, g_comm.rank, nvec, scale);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_broadcast(int64_t offset, int64_t count, int root, void* stream) {
  CT_REQUIRE(g_comm.ready, CT_ERR_COMM, "ct_broadcast: comm not initialised");
  CT_REQUIRE(offset >= 0 && count >= 0 && (offset % 4) == 0 && (count % 4) == 0 &&
                 (size_t)(offset + count) * 4 <= g_comm.data_bytes && root >= 0 && root < g_comm.world,
             CT_ERR_BAD_ARG, "ct_broadcast: bad range/root");
  if (count == 0) return 0;
  ArParams p;
  {
    std::lock_guard<std::mutex> lk(g_comm_mu);
    fill_params(p, offset, count, 1.f);
  }
  broadcast_kernel<<<32, 512, 0, (cudaStream_t)stream>>>(p, root);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_comm_finalize(void) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  if (g_comm.rank < 0) return 0;
  cudaDeviceSynchronize();  // teardown only
  for (void* p : g_comm.opened) cudaIpcCloseMemHandle(p);
  g_comm.opened.clear();
  if (g_comm.data[g_comm.rank]) cudaFree(g_comm.data[g_comm.rank]);
  if (g_comm.sig[g_comm.rank]) cudaFree(g_comm.sig[g_comm.rank]);
  g_comm = CommCtx();
  return 0;
}
