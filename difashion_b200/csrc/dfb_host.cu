#include "dfb_host.h"

#include <mutex>
#include <stdio.h>
#include <string.h>

#include "../../include/dfb200.h"

namespace dfb {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* what, cudaError_t e) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s (%s)", what, cudaGetErrorName(e),
           cudaGetErrorString(e));
}
void set_last_error_msg(const char* what) { snprintf(g_last_error, sizeof(g_last_error), "%s", what); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error_msg("cuTensorMapEncodeTiled driver entry point not available");
    return DFB_ERR_NO_DRIVER;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_last_error, sizeof(g_last_error),
             "cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,%llu,%llu] "
             "stride1=%llu box=[%u,%u,%u,%u] base=%p",
             (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
             (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
             (unsigned long long)(rank > 1 ? gstr[0] : 0), bdim[0], rank > 1 ? bdim[1] : 0,
             rank > 2 ? bdim[2] : 0, rank > 3 ? bdim[3] : 0, base);
    return DFB_ERR_CUDA;
  }
  return DFB_OK;
}

}  // namespace dfb

extern "C" {

const char* dfb_strerror(int rc) {
  switch (rc) {
    case DFB_OK: return "ok";
    case DFB_ERR_INVALID: return "invalid argument";
    case DFB_ERR_CUDA: return "CUDA error";
    case DFB_ERR_UNSUPPORTED: return "unsupported shape";
    case DFB_ERR_NO_DRIVER: return "tensor-map driver entry point unavailable";
    default: return "unknown error";
  }
}

const char* dfb_last_error(void) { return dfb::g_last_error; }

int dfb_abi_version(void) { return DFB200_ABI_VERSION; }

int dfb_num_sms(void) { return dfb::num_sms(); }

size_t dfb_sizeof_gemm_params(void) { return sizeof(dfb_gemm_params); }
size_t dfb_sizeof_attn_params(void) { return sizeof(dfb_attn_params); }

}  // extern "C"
