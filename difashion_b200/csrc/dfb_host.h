// Host-side helpers shared by the C-ABI entry points: error plumbing, device properties and
// CUtensorMap construction (driver entry point resolved at run time, so the library links only
// against libcudart).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dfb_common.cuh"

namespace dfb {

void set_last_error(const char* what, cudaError_t e);
void set_last_error_msg(const char* what);
int num_sms();

// rank-N (<=5) bf16 / fp32 tiled tensor map.  dims/box innermost-first; strides in BYTES for
// dims 1..rank-1 (dim 0 is contiguous).  Returns DFB_OK or an error code.
int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle);

#define DFB_CHECK_CUDA(expr)                                  \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) {                                  \
      ::dfb::set_last_error(#expr, _e);                       \
      return DFB_ERR_CUDA;                             \
    }                                                         \
  } while (0)

#define DFB_REQUIRE(cond, msg)                                \
  do {                                                        \
    if (!(cond)) {                                            \
      ::dfb::set_last_error_msg(msg " [" #cond "]");          \
      return DFB_ERR_INVALID;                          \
    }                                                         \
  } while (0)

}  // namespace dfb
