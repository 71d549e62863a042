// dvc_api.cu -- version / error / device plumbing of libdvc_b200.so.
#include <stdarg.h>
#include <string.h>

#include "dvc_common.cuh"

namespace dvc {

static thread_local char g_err[512] = "";

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

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    (void)cudaGetLastError();  // clear the sticky launch error
    return fail(DVC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }
  return DVC_OK;
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace dvc

extern "C" {

int dvc_version(void) { return DVC_VERSION_NUMBER; }

#ifndef DVC_SRC_HASH
#define DVC_SRC_HASH "unknown"
#endif
// sha256 (first 16 hex digits) of the sources and compiler flags this binary was built from
const char* dvc_build_info(void) { return "src=" DVC_SRC_HASH " arch=sm_100a"; }

const char* dvc_last_error_string(void) { return dvc::g_err; }

int dvc_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return dvc::fail(DVC_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  int sm = 0, ma = 0, mi = 0;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev);
  if (sm_count) *sm_count = sm;
  if (cc_major) *cc_major = ma;
  if (cc_minor) *cc_minor = mi;
  return DVC_OK;
}

}  // extern "C"
