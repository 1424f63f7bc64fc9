// Error plumbing + device check for the C ABI (include/mphsir.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mphsir {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return MPHSIR_ERR_CUDA;
  }
  return MPHSIR_OK;
}

}  // namespace mphsir

extern "C" int mphsir_version(void) { return MPHSIR_VERSION; }

extern "C" const char* mphsir_last_error(void) { return mphsir::g_err; }

extern "C" int mphsir_device_check(int device, int* sm_count) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    mphsir::set_error("cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
    return MPHSIR_ERR_CUDA;
  }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (prop.major != 10) {
    mphsir::set_error("libmphsir is built for sm_100a only; device %d is sm_%d%d", device, prop.major,
                      prop.minor);
    return MPHSIR_ERR_CUDA;
  }
  return MPHSIR_OK;
}
