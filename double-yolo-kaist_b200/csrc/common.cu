// Error reporting, device checks and driver entry points shared by all C-ABI functions.
#include "common.h"
#include <cstdlib>

#include <cstring>

namespace dyk {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

bool pdl_enabled() {
  static const bool on = !(getenv("DYK_PDL") != nullptr && getenv("DYK_PDL")[0] == '0');
  return on;
}

}  // namespace dyk

extern "C" __attribute__((visibility("default"))) int dyk_abi_version(void) { return DYK_ABI_VERSION; }
extern "C" __attribute__((visibility("default"))) int dyk_conv_params_size(void) { return (int)sizeof(dyk_conv_params); }

extern "C" __attribute__((visibility("default"))) const char* dyk_last_error(void) { return dyk::g_err; }

extern "C" __attribute__((visibility("default"))) int dyk_check_device(void) {
  int dev = 0;
  DYK_CUDA_OK(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  DYK_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DYK_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10)
    return dyk::fail(DYK_EARCH, "device %d has compute capability %d.%d; libdyk_b200 is built for sm_100a only",
                     dev, major, minor);
  if (!dyk::get_encode_tiled()) return dyk::fail(DYK_ECUDA, "cuTensorMapEncodeTiled not resolvable");
  return DYK_OK;
}
