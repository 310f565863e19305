// Library core: error string, capability probe, tensor-map encoding through the runtime-resolved driver entry point.
#include <stdarg.h>

#include <mutex>

#include "../../include/ekl_b200.h"
#include "ekl_common.cuh"

thread_local char g_ekl_err[512] = "";

int ekl_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_ekl_err, sizeof(g_ekl_err), fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

int ekl_make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int swizzle, int elem_bytes) {
  PFN_encodeTiled enc = get_encode();
  EKL_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  // The encode call is a DRIVER API call: unlike runtime calls it does not bind the device's primary context to the
  // calling thread.  A thread whose first CUDA work is this library's (e.g. an autograd worker thread) gets it bound by
  // one runtime call, once per thread.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    typedef CUresult (*PFN_ctxGetCurrent)(CUcontext*);
    static PFN_ctxGetCurrent get_ctx = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
      void* p = nullptr;
      cudaDriverEntryPointQueryResult qres;
      if (cudaGetDriverEntryPoint("cuCtxGetCurrent", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
        get_ctx = (PFN_ctxGetCurrent)p;
    });
    CUcontext cur = nullptr;
    if (get_ctx == nullptr || get_ctx(&cur) != CUDA_SUCCESS || cur == nullptr) EKL_CHECK_CUDA(cudaFree(nullptr));
    ctx_bound = true;
  }
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  static const CUtensorMapSwizzle sw[4] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_32B,
                                           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw[swizzle], CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return ekl_fail(1000 + (int)r,
                    "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] stride0 %llu swz %d base %p",
                    (int)r, rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
                    (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0), bx[0],
                    rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0, (unsigned long long)gs[0], swizzle, base);
  return 0;
}

extern "C" const char* ekl_last_error(void) { return g_ekl_err; }

extern "C" int ekl_version(void) { return EKL_B200_VERSION; }

extern "C" int ekl_require_sm100(void) {
  int dev = 0, major = 0, minor = 0;
  EKL_CHECK_CUDA(cudaGetDevice(&dev));
  EKL_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  EKL_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  EKL_REQUIRE(major == 10, "ekl_b200 needs an sm_100a device (B200); found sm_%d%d. There is no fallback path.", major, minor);
  return 0;
}
