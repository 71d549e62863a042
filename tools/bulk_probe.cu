// Probe: how many small 1-D bulk copies (cp.async.bulk, UBLKCP) per cycle does one SM sustain?
// The NCHW warp's tap box is ~30 rows of <= 288 bytes per channel plane; staging it with one
// bulk copy per row would take the copies off the LSU pipe (profiles/r02_planar.md) -- if the
// copy engine keeps up.  One CTA per SM, W warps; lane 0 of every warp issues its share of
// the R rows of a "plane", 4 planes in flight, mbarrier completion.  Prints cycles per row copy.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bulk_probe tools/bulk_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int kStages = 4;
constexpr int kMaxRows = 48, kMaxRowBytes = 320;

__global__ void __launch_bounds__(256) probe(const float* __restrict__ src, long long plane_floats,
                                             int pitch_floats, int rows, int row_bytes, int planes,
                                             int per_lane, long long* out, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ unsigned long long bar[kStages];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(&bar[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int per_col = 1088 / rows;                                         // a box per CTA
  const float* base = src + (long long)(blockIdx.x % per_col) * rows * pitch_floats + (blockIdx.x / per_col) * 128;
  auto issue = [&](int c) {
    const int s = c % kStages;
    const uint32_t b = smem_u32(&bar[s]);
    if (threadIdx.x == 0) mbar_expect_tx(b, (uint32_t)(rows * row_bytes));
    __syncthreads();   // the expect precedes the copies (as a per-plane barrier would in the kernel)
    if (per_lane) {    // every lane of every warp issues rows (the compiler serialises UBLKCP)
      for (int r = threadIdx.x; r < rows; r += blockDim.x)
        bulk_g2s(smem_u32(smem + (s * kMaxRows + r) * kMaxRowBytes),
                 base + (long long)c * plane_floats + (long long)r * pitch_floats, row_bytes, b);
    } else if (lane == 0) {
      for (int r = wid; r < rows; r += nw)
        bulk_g2s(smem_u32(smem + (s * kMaxRows + r) * kMaxRowBytes),
                 base + (long long)c * plane_floats + (long long)r * pitch_floats, row_bytes, b);
    }
  };
  float acc = 0.f;
  const long long t0 = clock64();
  for (int c = 0; c < kStages - 1 && c < planes; ++c) issue(c);
  for (int c = 0; c < planes; ++c) {
    const int s = c % kStages;
    while (!mbar_try(smem_u32(&bar[s]), (uint32_t)((c / kStages) & 1))) {}
    acc += reinterpret_cast<const float*>(smem + (s * kMaxRows + (threadIdx.x % rows)) * kMaxRowBytes)[threadIdx.x & 31];
    __syncthreads();
    if (c + kStages - 1 < planes) issue(c + kStages - 1);
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 12345.678f) sink[0] = acc;
}

int main() {
  const int H = 1088, W = 1920, C = 64;
  float* src; long long* out; float* sink;
  cudaMalloc(&src, sizeof(float) * (size_t)H * W * C);
  cudaMemset(src, 0, sizeof(float) * (size_t)H * W * C);
  cudaMalloc(&out, sizeof(long long) * 1024);
  cudaMalloc(&sink, 4);
  const int smem = kStages * kMaxRows * kMaxRowBytes;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int sms = 148;
  for (int per_lane = 0; per_lane <= 1; ++per_lane)
    for (int ctas_per_sm = 1; ctas_per_sm <= 3; ctas_per_sm += 2)
      for (int rows : {24, 32, 40})
        for (int row_bytes : {160, 256, 288}) {
          const int grid = sms * ctas_per_sm;
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          probe<<<grid, 256, smem>>>(src, (long long)H * W, W, rows, row_bytes, C, per_lane, out, sink);
          cudaEventRecord(e0);
          probe<<<grid, 256, smem>>>(src, (long long)H * W, W, rows, row_bytes, C, per_lane, out, sink);
          cudaEventRecord(e1);
          cudaError_t err = cudaDeviceSynchronize();
          if (err != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(err)); return 1; }
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          long long h[1024]; cudaMemcpy(h, out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
          double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
          const double copies = (double)C * rows;
          printf("per_lane=%d ctas/sm=%d rows=%2d row_bytes=%3d: %.1f cycles/plane, %.2f cycles per row copy per CTA, "
                 "%.2f us, %.1f GB/s aggregate\n", per_lane, ctas_per_sm, rows, row_bytes, avg / C, avg / copies,
                 ms * 1e3, (double)grid * copies * row_bytes / (ms * 1e-3) * 1e-9);
        }
  return 0;
}
