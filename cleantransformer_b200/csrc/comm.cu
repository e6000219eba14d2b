// comm.cu — gradient-bucket all-reduce as plain CUDA kernels over NVLink 5 / NVSwitch peer memory.
//
// Replaces the ncclAllReduce calls that torch's DDP reducer issues for the reference
// (examples/ft_bloom_DDP.py:99,126,135 `DDP(model, device_ids=[local_rank])`; README.md:46-52
// describes the hand-rolled version: param sync, gradient buckets, overlapped reduction).
//
// Memory. Every rank owns one symmetric buffer (gradient arena + staging + a page of signal words). Two ways to
// make it visible to the peers:
//   * VMM + multicast (default where the driver offers it): cuMemCreate / cuMemMap, the allocation exported as a
//     POSIX file descriptor that Python passes to the peers over a Unix socket (SCM_RIGHTS), every peer maps it
//     (unicast pointers), and all ranks bind their allocation to ONE multicast object (cuMulticastCreate /
//     cuMulticastBindMem): a store to the multicast address lands in every GPU's buffer, a `multimem.ld_reduce`
//     returns the sum over all GPUs, added inside the NVSwitch (NVLS);
//   * cudaMalloc + cudaIpcMemHandle (fallback): unicast peer pointers only.
// The gradient arena (arena.py) lives inside the symmetric buffer, so wgrad kernels write straight into
// NVLink-visible memory and the reduction is in place.
//
// All-reduce of a bucket [offset, offset+count), one kernel per rank:
//   phase 0  signal "my data for epoch e is ready" to every peer, wait for all peers
//   phase 1  rank r owns slice r of the range
//            NVLS:     v = multimem.ld_reduce.add(slice element)  (the switch reads all W copies and adds);
//                      multimem.st(v * scale)                      (the switch writes all W copies)
//                      -> per rank 1/W of the bucket crosses the SM twice, instead of (W-1)/W of it W-fold: a
//                      handful of small CTAs keep NVLink busy, and the persistent GEMMs of backward keep their SMs;
//            unicast:  128-bit loads of the slice from all W ranks, fp32 sum in rank order, * scale, 128-bit
//                      stores of the result to all W ranks
//   phase 2  last CTA signals "my slice is written everywhere", waits for all peers' signals
// One-shot variant (each rank reads everything from every peer, writes only locally) for small,
// latency-bound buckets. Every rank ends up with bit-identical bucket contents in all modes (one owner adds,
// everybody receives its bits).
//
// Epochs live in DEVICE memory (a word of the local signal page, bumped by the last CTA of every collective):
// the kernels take no per-call host counter, so a captured CUDA graph replays them correctly.
#include "ct_common.cuh"
#include "../../include/ct_b200.h"
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unistd.h>
#include <vector>

namespace ct {

constexpr int MAX_WORLD = 16;
constexpr size_t SIG_BYTES = 4096;          // signal page behind the data (VMM mode) / own cudaMalloc (IPC mode)
constexpr int SIG_COUNTER = 2 * MAX_WORLD;  // CTA counter (local use only)
constexpr int SIG_EPOCH = 2 * MAX_WORLD + 1;  // epoch of the last finished collective (local use only)

struct Drv {
  bool ok = false;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemExport)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImport)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MemGetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*McCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*McAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*McBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                        unsigned long long) = nullptr;
  CUresult (*McGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
};

static Drv& drv() {
  static Drv d;
  static std::once_flag once;
  std::call_once(once, [] {
    bool ok = true;
    auto get = [&](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess ||
          *fn == nullptr)
        ok = false;
    };
    get("cuMemCreate", (void**)&d.MemCreate);
    get("cuMemRelease", (void**)&d.MemRelease);
    get("cuMemAddressReserve", (void**)&d.MemAddressReserve);
    get("cuMemAddressFree", (void**)&d.MemAddressFree);
    get("cuMemMap", (void**)&d.MemMap);
    get("cuMemUnmap", (void**)&d.MemUnmap);
    get("cuMemSetAccess", (void**)&d.MemSetAccess);
    get("cuMemExportToShareableHandle", (void**)&d.MemExport);
    get("cuMemImportFromShareableHandle", (void**)&d.MemImport);
    get("cuMemGetAllocationGranularity", (void**)&d.MemGetGranularity);
    get("cuMulticastCreate", (void**)&d.McCreate);
    get("cuMulticastAddDevice", (void**)&d.McAddDevice);
    get("cuMulticastBindMem", (void**)&d.McBindMem);
    get("cuMulticastGetGranularity", (void**)&d.McGetGranularity);
    get("cuDeviceGet", (void**)&d.DeviceGet);
    get("cuDeviceGetAttribute", (void**)&d.DeviceGetAttribute);
    d.ok = ok;
  });
  return d;
}

#define CT_CU_OK(expr)                                                                    \
  do {                                                                                    \
    CUresult _r = (expr);                                                                 \
    if (_r != CUDA_SUCCESS) {                                                             \
      ct::set_error("%s:%d: %s -> CUresult %d", __FILE__, __LINE__, #expr, (int)_r);      \
      return CT_ERR_COMM;                                                                 \
    }                                                                                     \
  } while (0)

struct CommCtx {
  int rank = -1, world = 0, device = 0;
  bool ready = false;
  bool vmm = false;                           // buffers come from cuMemCreate (else cudaMalloc + IPC)
  float* data[MAX_WORLD] = {nullptr};         // symmetric buffers (local at [rank])
  uint32_t* sig[MAX_WORLD] = {nullptr};       // signal words: [0..W) ready flags, [W..2W) done flags, counter, epoch
  float* mc_data = nullptr;                   // multicast mapping of the same bytes on every rank (NVLS), or null
  size_t data_bytes = 0;                      // usable floats * 4
  size_t alloc_bytes = 0;                     // VMM: granularity-rounded size of the allocation (data + signal page)
  CUmemGenericAllocationHandle mem = 0, mc = 0;
  CUmemGenericAllocationHandle peer_mem[MAX_WORLD] = {0};
  std::vector<void*> opened;                  // IPC mappings
};
static CommCtx g_comm;
static std::mutex g_comm_mu;

struct ArParams {
  float* data[MAX_WORLD];
  uint32_t* sig[MAX_WORLD];
  float* mc;
  int rank, world;
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
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
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
// NVLS: one load returns the fp32 sum of the addressed 16 bytes over every GPU bound to the multicast object (the
// NVSwitch reads the W copies and adds); one store writes all W copies.
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y),
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

// Every collective kernel: epoch = (last finished epoch of this rank) + 1, read from device memory. All ranks run
// the same sequence of collectives on their communication stream, so the counters agree without any host state.
__device__ __forceinline__ uint32_t begin_collective(const ArParams& p) {
  const uint32_t epoch = ld_volatile_u32(p.sig[p.rank] + SIG_EPOCH) + 1u;
  if (blockIdx.x == 0 && (int)threadIdx.x < p.world) st_release_sys(p.sig[threadIdx.x] + p.rank, epoch);
  wait_flags(p.sig[p.rank], p.world, epoch);
  return epoch;
}
// The last CTA to get here publishes completion to the peers, waits for theirs and closes the epoch.
__device__ __forceinline__ void end_collective(const ArParams& p, uint32_t epoch) {
  uint32_t* my_sig = p.sig[p.rank];
  __threadfence_system();
  __syncthreads();
  __shared__ int last;
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(my_sig + SIG_COUNTER, 1u);
    last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) my_sig[SIG_COUNTER] = 0;  // reset for the next call (stream-ordered)
  if ((int)threadIdx.x < p.world) st_release_sys(p.sig[threadIdx.x] + MAX_WORLD + p.rank, epoch);
  wait_flags(my_sig + MAX_WORLD, p.world, epoch);
  if (threadIdx.x == 0) {
    my_sig[SIG_EPOCH] = epoch;
    __threadfence();
  }
}

constexpr int AR_THREADS = 256;

// MODE 0 two-shot over unicast peer pointers, 1 one-shot, 2 two-shot through the multicast mapping (NVLS)
template <int MODE>
__global__ void __launch_bounds__(AR_THREADS, 6)  // <= 40 registers: fits beside a resident GEMM CTA
    allreduce_kernel(const ArParams p) {
  const int W = p.world, r = p.rank;
  const uint32_t epoch = begin_collective(p);
  const int64_t nvec = p.count >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (MODE == 1) {
    // every rank reduces the whole range; results are staged behind the range (nobody may overwrite its input
    // while peers still read it) and copied back by a second kernel
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int q = 0; q < W; ++q) {
        const float4 v = ld_peer_f4(p.data[q] + p.offset + 4 * i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      acc.x *= p.scale; acc.y *= p.scale; acc.z *= p.scale; acc.w *= p.scale;
      st_peer_f4(p.data[r] + p.offset + p.count + 4 * i, acc);
    }
  } else {
    // slice owned by this rank (multiple of 4 elements); four independent 16-byte vectors per thread per iteration
    const int64_t per = ((nvec + W - 1) / W);
    const int64_t v0 = min(nvec, per * r), v1 = min(nvec, per * (r + 1));
    for (int64_t i = v0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v1; i += 4 * stride) {
      float4 acc[4];
      if (MODE == 2) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t idx = i + u * stride;
          acc[u] = idx < v1 ? multimem_ld_reduce_f4(p.mc + p.offset + 4 * idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
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
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc[u].x *= p.scale; acc[u].y *= p.scale; acc[u].z *= p.scale; acc[u].w *= p.scale;
      }
      if (MODE == 2) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t idx = i + u * stride;
          if (idx < v1) multimem_st_f4(p.mc + p.offset + 4 * idx, acc[u]);
        }
      } else {
        for (int q = 0; q < W; ++q) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int64_t idx = i + u * stride;
            if (idx < v1) st_peer_f4(p.data[q] + p.offset + 4 * idx, acc[u]);
          }
        }
      }
    }
  }
  end_collective(p, epoch);
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
  const uint32_t epoch = begin_collective(p);
  if (p.rank != root) {
    const int64_t nvec = p.count >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x)
      st_peer_f4(p.data[p.rank] + p.offset + 4 * i, ld_peer_f4(p.data[root] + p.offset + 4 * i));
  }
  end_collective(p, epoch);
}

__global__ void __launch_bounds__(32) comm_barrier_kernel(const ArParams p) {
  const uint32_t epoch = begin_collective(p);
  end_collective(p, epoch);
}

// ---------------------------------------------------------------------------------------------------
// Tied-embedding gradient, sparse half. The LM-head wgrad of a tied table is dense and ready at the START of
// backward; the embedding scatter-add arrives at the very END. Reducing the table once both are there
// (what a bucketed reducer does) leaves a ~1 GB all-reduce with nothing to hide behind. Instead the dense
// half is all-reduced as soon as the LM-head wgrad is enqueued, and the sparse half is exchanged as what it
// is: every rank publishes its T x H token gradients + token ids in the symmetric buffer, and every rank
// scatter-adds ALL ranks' rows (scaled by 1/W) into its own, already averaged, table gradient:
// (W-1) * T * H * 4 bytes inbound per rank (235 MB at W = 8) instead of 2 (W-1)/W x 1 GB.
// (fp32 atomics: the table gradients of different ranks agree to rounding, not bit for bit.)
struct EsParams {
  ArParams a;
  int64_t hdr_off;    // float offset of the staging header: int64 T (tokens this rank staged)
  int64_t rows_off;   // float offset of the staged rows [cap, H] f32
  int64_t ids_off;    // float offset of the staged ids [cap] int64
  int64_t grad_off;   // float offset of the table gradient [V, H] in the LOCAL buffer
  int64_t H, V;
  long long padding_idx;
};

__device__ __forceinline__ long long ld_peer_s64(const long long* p) {
  long long v;
  asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
    embed_scatter_allranks_kernel(const EsParams p) {
  const int W = p.a.world, r = p.a.rank;
  const uint32_t epoch = begin_collective(p.a);
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float* grad = p.a.data[r] + p.grad_off;
  const float scale = p.a.scale;
  for (int k = 0; k < W; ++k) {
    const int q = (r + k) % W;  // own rows first, then the peers round-robin (spreads the NVLink load)
    const long long Tq = ld_peer_s64(reinterpret_cast<const long long*>(p.a.data[q] + p.hdr_off));
    const long long* ids = reinterpret_cast<const long long*>(p.a.data[q] + p.ids_off);
    const float* rows = p.a.data[q] + p.rows_off;
    for (int64_t t = warp0; t < Tq; t += nwarps) {
      const long long id = ld_peer_s64(ids + t);
      if (id < 0 || id >= p.V || id == p.padding_idx) continue;
      const float* src = rows + t * p.H;
      float* dst = grad + id * p.H;
      // eight 16-byte peer loads in flight per lane (a 1024-wide row in one go) BEFORE the first reduction: a
      // volatile red after every load serialised them on the NVLink round trip
      for (int64_t c0 = 0; c0 < p.H; c0 += 1024) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int64_t c = c0 + u * 128 + lane * 4;
          v[u] = c < p.H ? ld_peer_f4(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int64_t c = c0 + u * 128 + lane * 4;
          if (c < p.H)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c), "f"(v[u].x * scale),
                         "f"(v[u].y * scale), "f"(v[u].z * scale), "f"(v[u].w * scale)
                         : "memory");
        }
      }
    }
  }
  end_collective(p.a, epoch);  // nobody restages while a peer still reads
}

static void fill_params(ArParams& p, int64_t offset, int64_t count, float scale) {
  for (int q = 0; q < g_comm.world; ++q) { p.data[q] = g_comm.data[q]; p.sig[q] = g_comm.sig[q]; }
  p.mc = g_comm.mc_data;
  p.rank = g_comm.rank; p.world = g_comm.world;
  p.offset = offset; p.count = count; p.scale = scale;
}

static int set_timeout(int device) {
  const char* e = getenv("CT_COMM_TIMEOUT_S");
  const double secs = e ? atof(e) : 1800.0;
  int khz = 2000000;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
  const long long cycles = secs > 0 ? (long long)(secs * 1e3 * (double)khz) : 0;
  CT_CUDA_OK(cudaMemcpyToSymbol(g_wait_timeout_cycles, &cycles, sizeof(cycles)));
  return 0;
}

static size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int map_rw(CUmemGenericAllocationHandle h, size_t bytes, size_t gran, int device, CUdeviceptr* out) {
  Drv& D = drv();
  CUdeviceptr va = 0;
  CT_CU_OK(D.MemAddressReserve(&va, bytes, gran, 0, 0));
  CT_CU_OK(D.MemMap(va, bytes, 0, h, 0));
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  CT_CU_OK(D.MemSetAccess(va, bytes, &acc, 1));
  *out = va;
  return 0;
}

}  // namespace ct

using namespace ct;

// ---------------------------------------------------------------------------------------------------
// set-up, IPC flavour
// ---------------------------------------------------------------------------------------------------
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
  CT_CUDA_OK(cudaMalloc(&s, SIG_BYTES));
  CT_CUDA_OK(cudaMemset(d, 0, data_bytes));
  CT_CUDA_OK(cudaMemset(s, 0, SIG_BYTES));
  CT_CUDA_OK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)data_handle_out, d));
  CT_CUDA_OK(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)sig_handle_out, s));
  int rc = set_timeout(device);
  if (rc) return rc;
  g_comm.rank = rank; g_comm.world = world; g_comm.device = device;
  g_comm.vmm = false;
  g_comm.data[rank] = d; g_comm.sig[rank] = s;
  g_comm.data_bytes = data_bytes;
  *local_data = d;
  return 0;
}

extern "C" int ct_comm_connect(const void* data_handles, const void* sig_handles) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  CT_REQUIRE(g_comm.rank >= 0 && !g_comm.vmm, CT_ERR_COMM, "ct_comm_connect: ct_comm_init first");
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

// ---------------------------------------------------------------------------------------------------
// set-up, VMM + multicast flavour
// ---------------------------------------------------------------------------------------------------
extern "C" int ct_comm_vmm_supported(int device, int* vmm_ok, int* multicast_ok) {
  CT_REQUIRE(vmm_ok && multicast_ok, CT_ERR_BAD_ARG, "ct_comm_vmm_supported: null out");
  *vmm_ok = *multicast_ok = 0;
  Drv& D = drv();
  if (!D.ok) return 0;
  CT_CUDA_OK(cudaSetDevice(device));
  CT_CUDA_OK(cudaFree(0));
  CUdevice dev;
  CT_CU_OK(D.DeviceGet(&dev, device));
  int a = 0, b = 0, c = 0;
  D.DeviceGetAttribute(&a, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev);
  D.DeviceGetAttribute(&b, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
  D.DeviceGetAttribute(&c, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
  *vmm_ok = (a && b) ? 1 : 0;
  *multicast_ok = (a && b && c) ? 1 : 0;
  return 0;
}

extern "C" int ct_comm_vmm_init(int rank, int world, int device, size_t data_bytes, void** local_data,
                                int* local_fd_out) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  CT_REQUIRE(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world, CT_ERR_BAD_ARG,
             "ct_comm_vmm_init: bad rank/world %d/%d", rank, world);
  CT_REQUIRE(!g_comm.ready && g_comm.rank < 0, CT_ERR_COMM, "ct_comm_vmm_init: already initialised");
  CT_REQUIRE(local_data && local_fd_out, CT_ERR_BAD_ARG, "ct_comm_vmm_init: null out");
  Drv& D = drv();
  CT_REQUIRE(D.ok, CT_ERR_UNSUPPORTED, "ct_comm_vmm_init: driver VMM / multicast entry points not available");
  CT_CUDA_OK(cudaSetDevice(device));
  CT_CUDA_OK(cudaFree(0));  // primary context current on this thread
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  CT_CU_OK(D.MemGetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  data_bytes = round_up(data_bytes, 4096);
  {
    // the multicast object wants its own (possibly larger) granularity for size and binding offsets
    CUmulticastObjectProp mp;
    memset(&mp, 0, sizeof(mp));
    mp.numDevices = (unsigned)world;
    mp.size = round_up(data_bytes + SIG_BYTES, gran);
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (D.McGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > gran) gran = mg;
  }
  const size_t total = round_up(data_bytes + SIG_BYTES, gran);
  CUmemGenericAllocationHandle h = 0;
  CT_CU_OK(D.MemCreate(&h, total, &prop, 0));
  CUdeviceptr va = 0;
  int rc = map_rw(h, total, gran, device, &va);
  if (rc) return rc;
  CT_CUDA_OK(cudaMemset((void*)va, 0, total));
  CT_CUDA_OK(cudaDeviceSynchronize());
  int fd = -1;
  CT_CU_OK(D.MemExport(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  rc = set_timeout(device);
  if (rc) return rc;
  g_comm.rank = rank; g_comm.world = world; g_comm.device = device;
  g_comm.vmm = true;
  g_comm.mem = h;
  g_comm.alloc_bytes = total;
  g_comm.data_bytes = data_bytes;
  g_comm.data[rank] = (float*)va;
  g_comm.sig[rank] = (uint32_t*)((char*)va + data_bytes);
  *local_data = (void*)va;
  *local_fd_out = fd;
  return 0;
}

// peer_fds: [world] file descriptors received from the peers (entry [rank] ignored); the caller closes them
extern "C" int ct_comm_vmm_connect(const int* peer_fds) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  CT_REQUIRE(g_comm.rank >= 0 && g_comm.vmm, CT_ERR_COMM, "ct_comm_vmm_connect: ct_comm_vmm_init first");
  CT_REQUIRE(peer_fds, CT_ERR_BAD_ARG, "ct_comm_vmm_connect: null fds");
  Drv& D = drv();
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = g_comm.device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  CT_CU_OK(D.MemGetGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  for (int q = 0; q < g_comm.world; ++q) {
    if (q == g_comm.rank) continue;
    CUmemGenericAllocationHandle ph = 0;
    CT_CU_OK(D.MemImport(&ph, (void*)(uintptr_t)peer_fds[q], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
    CUdeviceptr va = 0;
    int rc = map_rw(ph, g_comm.alloc_bytes, gran, g_comm.device, &va);
    if (rc) return rc;
    g_comm.peer_mem[q] = ph;
    g_comm.data[q] = (float*)va;
    g_comm.sig[q] = (uint32_t*)((char*)va + g_comm.data_bytes);
  }
  g_comm.ready = true;
  return 0;
}

// Multicast object over the symmetric buffers: rank 0 creates it (and exports a descriptor for the peers), everyone
// else imports; then ALL ranks add their device (barrier), bind their allocation and map the object (barrier).
extern "C" int ct_comm_mc_create(int* mc_fd_out) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  CT_REQUIRE(g_comm.ready && g_comm.vmm && mc_fd_out, CT_ERR_COMM, "ct_comm_mc_create: VMM comm not connected");
  Drv& D = drv();
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof(mp));
  mp.numDevices = (unsigned)g_comm.world;
  mp.size = g_comm.alloc_bytes;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  CT_CU_OK(D.McCreate(&g_comm.mc, &mp));
  int fd = -1;
  CT_CU_OK(D.MemExport(&fd, g_comm.mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  *mc_fd_out = fd;
  return 0;
}

extern "C" int ct_comm_mc_import(int mc_fd) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  CT_REQUIRE(g_comm.ready && g_comm.vmm, CT_ERR_COMM, "ct_comm_mc_import: VMM comm not connected");
  CT_CU_OK(drv().MemImport(&g_comm.mc, (void*)(uintptr_t)mc_fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  return 0;
}

extern "C" int ct_comm_mc_add_device(void) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  CT_REQUIRE(g_comm.ready && g_comm.vmm && g_comm.mc != 0, CT_ERR_COMM, "ct_comm_mc_add_device: no multicast object");
  Drv& D = drv();
  CUdevice dev;
  CT_CU_OK(D.DeviceGet(&dev, g_comm.device));
  CT_CU_OK(D.McAddDevice(g_comm.mc, dev));
  return 0;
}

extern "C" int ct_comm_mc_bind(void) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  CT_REQUIRE(g_comm.ready && g_comm.vmm && g_comm.mc != 0, CT_ERR_COMM, "ct_comm_mc_bind: no multicast object");
  Drv& D = drv();
  CT_CU_OK(D.McBindMem(g_comm.mc, 0, g_comm.mem, 0, g_comm.alloc_bytes, 0));
  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof(mp));
  mp.numDevices = (unsigned)g_comm.world;
  mp.size = g_comm.alloc_bytes;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t mg = 0;
  CT_CU_OK(D.McGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
  CUdeviceptr va = 0;
  int rc = map_rw(g_comm.mc, g_comm.alloc_bytes, mg, g_comm.device, &va);
  if (rc) return rc;
  g_comm.mc_data = (float*)va;
  return 0;
}

// flags[0] = buffers come from the VMM API, flags[1] = multicast (NVLS) mapping present
extern "C" int ct_comm_info(int* flags) {
  CT_REQUIRE(flags, CT_ERR_BAD_ARG, "ct_comm_info: null out");
  flags[0] = g_comm.ready && g_comm.vmm ? 1 : 0;
  flags[1] = g_comm.ready && g_comm.mc_data != nullptr ? 1 : 0;
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// collectives
// ---------------------------------------------------------------------------------------------------
extern "C" int ct_allreduce_bucket(int64_t offset, int64_t count, float scale, int mode, int max_ctas,
                                   void* stream) {
  CT_REQUIRE(g_comm.ready, CT_ERR_COMM, "ct_allreduce_bucket: comm not initialised");
  CT_REQUIRE(offset >= 0 && count >= 0 && (offset % 4) == 0 && (count % 4) == 0 &&
                 (size_t)(offset + count) * 4 <= g_comm.data_bytes,
             CT_ERR_BAD_ARG, "ct_allreduce_bucket: range [%lld,+%lld) must be 16-byte aligned and inside the buffer",
             (long long)offset, (long long)count);
  CT_REQUIRE(mode >= 0 && mode <= 3, CT_ERR_BAD_ARG,
             "ct_allreduce_bucket: mode 0 (auto), 1 (one-shot), 2 (NVLS multimem), 3 (two-shot unicast)");
  CT_REQUIRE(mode != 2 || g_comm.mc_data != nullptr, CT_ERR_UNSUPPORTED, "ct_allreduce_bucket: no multicast mapping");
  if (count == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  ArParams p;
  fill_params(p, offset, count, scale);
  if (mode == 0) mode = g_comm.mc_data != nullptr ? 2 : 3;
  {
    // same shared-memory carve-out as the big GEMM / attention kernels so that an all-reduce CTA can
    // be co-resident with them (an SM cannot host CTAs that want different L1/shared splits)
    static bool carve = false;
    if (!carve) {
      cudaFuncSetAttribute(allreduce_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      cudaFuncSetAttribute(allreduce_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      cudaFuncSetAttribute(allreduce_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      carve = true;
    }
  }
  const int64_t nvec = count >> 2;
  if (mode == 1) {
    // one-shot needs a staging area of `count` floats right after the range
    CT_REQUIRE((size_t)(offset + 2 * count) * 4 <= g_comm.data_bytes, CT_ERR_WORKSPACE,
               "ct_allreduce_bucket: one-shot needs a staging area after the range");
    if (max_ctas <= 0) max_ctas = 64;
    int64_t ctas = (nvec + AR_THREADS - 1) / AR_THREADS;
    if (ctas > max_ctas) ctas = max_ctas;
    allreduce_kernel<1><<<(unsigned)ctas, AR_THREADS, 0, st>>>(p);
    CT_LAUNCH_OK();
    copy_f4_kernel<<<(unsigned)ctas, 256, 0, st>>>(g_comm.data[g_comm.rank] + offset,
                                                    g_comm.data[g_comm.rank] + offset + count, nvec);
    CT_LAUNCH_OK();
    return 0;
  }
  if (max_ctas <= 0) max_ctas = mode == 2 ? 16 : 64;
  int64_t per = (nvec + g_comm.world - 1) / g_comm.world;
  int64_t ctas = (per + 4 * AR_THREADS - 1) / (4 * AR_THREADS);
  if (ctas > max_ctas) ctas = max_ctas;
  if (ctas < 1) ctas = 1;
  if (mode == 2) allreduce_kernel<2><<<(unsigned)ctas, AR_THREADS, 0, st>>>(p);
  else allreduce_kernel<0><<<(unsigned)ctas, AR_THREADS, 0, st>>>(p);
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
  fill_params(p.a, 0, 0, scale);
  p.hdr_off = hdr_offset; p.rows_off = rows_offset; p.ids_off = ids_offset; p.grad_off = grad_offset;
  p.H = H; p.V = V; p.padding_idx = (long long)padding_idx;
  if (max_ctas <= 0) max_ctas = sm_count() * 2;
  embed_scatter_allranks_kernel<<<(unsigned)max_ctas, 256, 0, (cudaStream_t)stream>>>(p);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_comm_barrier(void* stream) {
  CT_REQUIRE(g_comm.ready, CT_ERR_COMM, "ct_comm_barrier: comm not initialised");
  ArParams p;
  fill_params(p, 0, 0, 1.f);
  comm_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(p);
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
  fill_params(p, offset, count, 1.f);
  broadcast_kernel<<<32, 512, 0, (cudaStream_t)stream>>>(p, root);
  CT_LAUNCH_OK();
  return 0;
}

extern "C" int ct_comm_finalize(void) {
  std::lock_guard<std::mutex> lk(g_comm_mu);
  if (g_comm.rank < 0) return 0;
  cudaDeviceSynchronize();  // teardown only
  if (g_comm.vmm) {
    Drv& D = drv();
    if (g_comm.mc_data) {
      D.MemUnmap((CUdeviceptr)g_comm.mc_data, g_comm.alloc_bytes);
      D.MemAddressFree((CUdeviceptr)g_comm.mc_data, g_comm.alloc_bytes);
    }
    for (int q = 0; q < g_comm.world; ++q) {
      if (!g_comm.data[q]) continue;
      D.MemUnmap((CUdeviceptr)g_comm.data[q], g_comm.alloc_bytes);
      D.MemAddressFree((CUdeviceptr)g_comm.data[q], g_comm.alloc_bytes);
      if (q != g_comm.rank && g_comm.peer_mem[q]) D.MemRelease(g_comm.peer_mem[q]);
    }
    if (g_comm.mc) D.MemRelease(g_comm.mc);
    if (g_comm.mem) D.MemRelease(g_comm.mem);
  } else {
    for (void* p : g_comm.opened) cudaIpcCloseMemHandle(p);
    g_comm.opened.clear();
    if (g_comm.data[g_comm.rank]) cudaFree(g_comm.data[g_comm.rank]);
    if (g_comm.sig[g_comm.rank]) cudaFree(g_comm.sig[g_comm.rank]);
  }
  g_comm = CommCtx();
  return 0;
}
