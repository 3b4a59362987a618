// Host-side helpers shared by the C-ABI translation units: error reporting, driver entry points.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/dyk_b200.h"

namespace dyk {

// Thread-local last-error message (dyk_last_error()).
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define DYK_CUDA_OK(expr)                                                                    \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ::dyk::fail(DYK_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                         __FILE__, __LINE__);                                                \
  } while (0)

#define DYK_REQUIRE(cond, ...)                              \
  do {                                                      \
    if (!(cond)) return ::dyk::fail(DYK_EINVAL, __VA_ARGS__); \
  } while (0)

// Checks the launch that was just issued (does not synchronise).
#define DYK_LAUNCH_OK(name)                                                                         \
  do {                                                                                              \
    cudaError_t _e = cudaPeekAtLastError();                                                         \
    if (_e != cudaSuccess) {                                                                        \
      cudaGetLastError();                                                                           \
      return ::dyk::fail(DYK_ECUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e));       \
    }                                                                                               \
  } while (0)

// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_tiled();

int num_sms();
// SMs a convolution launch may occupy (dyk_conv_params.sm_limit; even, >= 2)
static inline int sm_budget(int sm_limit) {
  const int all = num_sms();
  if (sm_limit <= 0 || sm_limit >= all) return all;
  return sm_limit < 2 ? 2 : (sm_limit & ~1);
}

// Launch with programmatic stream serialization: the kernel may begin (prologue only — it calls griddepcontrol.wait before
// touching global data) while the tail of the previous kernel in the stream is still running.  DYK_PDL=0 disables it.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace dyk
