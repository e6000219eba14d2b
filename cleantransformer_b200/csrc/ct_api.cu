// ct_api.cu — library-level entry points of libct_b200.so: version, error retrieval, device
// check, TMA descriptor construction. See include/ct_b200.h for the C ABI contract.
#include "ct_common.cuh"
#include "../../include/ct_b200.h"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace ct {

static thread_local char g_err[512] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, int swizzle_128b) {
  PFN_encodeTiled enc = get_encode();
  CT_REQUIRE(enc != nullptr, CT_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available");
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                           : (elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                              : CU_TENSOR_MAP_DATA_TYPE_UINT8);
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_128b ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult=%d (base=%p rank=%d dims=%llu,%llu "
              "stride1=%llu box=%u,%u)",
              (int)r, base, rank, (unsigned long long)dims[0],
              (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 1 ? strides_bytes[1] : 0), box[0], rank > 1 ? box[1] : 0);
    return CT_ERR_BAD_ARG;
  }
  return 0;
}

// ---- tuning knobs (kernel-variant selection for A/B runs and tests) ------------------------------
// Initial values come from the environment (CT_<NAME>), ct_set_option overrides at run time.
static int g_opt[OPT_COUNT];
static std::once_flag g_opt_once;
static const char* const kOptNames[OPT_COUNT] = {"LN_BWD_IMPL", "GEMM_EPI_IMPL",
                                                 "GEMM_2CTA", "CE_IMPL", "GEMM_SPLITK", "ATTN_BWD_IMPL", "PDL"};
static const int kOptDefaults[OPT_COUNT] = {0, 0, 1, 0, 0, 0, 0};
static void opt_init() {
  std::call_once(g_opt_once, [] {
    for (int i = 0; i < OPT_COUNT; ++i) {
      char name[64];
      snprintf(name, sizeof(name), "CT_%s", kOptNames[i]);
      const char* e = getenv(name);
      g_opt[i] = e ? atoi(e) : kOptDefaults[i];
    }
  });
}
int option(int which) {
  opt_init();
  return (which >= 0 && which < OPT_COUNT) ? __atomic_load_n(&g_opt[which], __ATOMIC_RELAXED) : 0;
}

}  // namespace ct

extern "C" {

int ct_set_option(const char* name, int value) {
  ct::opt_init();
  CT_REQUIRE(name != nullptr, CT_ERR_BAD_ARG, "ct_set_option: null name");
  for (int i = 0; i < ct::OPT_COUNT; ++i)
    if (strcmp(name, ct::kOptNames[i]) == 0) {
      __atomic_store_n(&ct::g_opt[i], value, __ATOMIC_RELAXED);
      return 0;
    }
  ct::set_error("ct_set_option: unknown option '%s'", name);
  return CT_ERR_BAD_ARG;
}

int ct_get_option(const char* name, int* value) {
  ct::opt_init();
  CT_REQUIRE(name != nullptr && value != nullptr, CT_ERR_BAD_ARG, "ct_get_option: null argument");
  for (int i = 0; i < ct::OPT_COUNT; ++i)
    if (strcmp(name, ct::kOptNames[i]) == 0) {
      *value = ct::option(i);
      return 0;
    }
  ct::set_error("ct_get_option: unknown option '%s'", name);
  return CT_ERR_BAD_ARG;
}

int ct_version(void) { return CT_B200_VERSION; }

int ct_last_error(char* buf, size_t n) {
  if (!buf || n == 0) return CT_ERR_BAD_ARG;
  strncpy(buf, ct::get_error(), n - 1);
  buf[n - 1] = 0;
  return 0;
}

int ct_device_check(int device) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    ct::set_error("no CUDA device visible (%s)", cudaGetErrorString(e));
    return CT_ERR_UNSUPPORTED;
  }
  CT_REQUIRE(device >= 0 && device < ndev, CT_ERR_BAD_ARG, "device %d out of range [0,%d)", device,
             ndev);
  int major = 0, minor = 0;
  CT_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  CT_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  CT_REQUIRE(major == 10 && minor == 0, CT_ERR_UNSUPPORTED,
             "device %d is sm_%d%d; libct_b200 contains sm_100a code only", device, major, minor);
  return 0;
}

}  // extern "C"
