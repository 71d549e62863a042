// dvc_common.cuh -- shared helpers of libdvc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "dvc_b200.h"

#define DVC_VERSION_NUMBER 100  // 0.1.0

namespace dvc {

// ---- error plumbing (thread local; the ABI never throws) -------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);   // cudaPeekAtLastError -> status

#define DVC_REQUIRE(cond, ...)                                        \
  do {                                                                \
    if (!(cond)) return ::dvc::fail(DVC_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
  } while (0)

int sm_count();

// ---- strides ---------------------------------------------------------------
struct Strides4 {
  int64_t n, c, h, w;
};
static inline Strides4 make_strides(const int64_t s[4]) {
  Strides4 r;
  r.n = s[0]; r.c = s[1]; r.h = s[2]; r.w = s[3];
  return r;
}
static inline bool aligned16(const void* p) {
  return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}
// channels_last fast path: channel stride 1, vectorisable by 4.
static inline bool nhwc_vec4_ok(const void* p, const Strides4& s, int64_t C) {
  return s.c == 1 && (C % 4) == 0 && aligned16(p) && (s.n % 4) == 0 &&
         (s.h % 4) == 0 && (s.w % 4) == 0;
}
static inline bool fits_int32(int64_t N, int64_t C, int64_t H, int64_t W,
                              const Strides4& s) {
  // largest element offset must fit in int32 for the fast 32-bit index paths
  long double m = (long double)(N - 1) * s.n + (long double)(C - 1) * s.c +
                  (long double)(H - 1) * s.h + (long double)(W - 1) * s.w;
  return m < 2147483647.0L;
}

// ---- arithmetic that must NOT be contracted into FMAs ----------------------
// PyTorch eager evaluates every tensor op in its own kernel, i.e. every
// product and sum is rounded separately.  nvcc contracts a*b+c by default, so
// wherever we replay an eager op sequence we spell the rounding out.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }

// ---- warp / block reductions ----------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- cache-hinted 128-bit accesses ----------------------------------------
__device__ __forceinline__ void st_streaming(float4* p, const float4& v) {
  // write-once output: evict-first in L2, do not keep in L1
  __stcs(p, v);
}

}  // namespace dvc
