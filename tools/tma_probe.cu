#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#ifndef MODE
#define MODE 0
#endif
#ifndef SWZ
#define SWZ CU_TENSOR_MAP_SWIZZLE_NONE
#endif
#ifndef RANK
#define RANK 4
#endif
#ifndef BOXW
#define BOXW 64
#endif
#ifndef BOXH
#define BOXH 24
#endif
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#ifdef MAPGLOBAL
__global__ void k(const CUtensorMap* gmap, float* out,
#else
__global__ void k(const __grid_constant__ CUtensorMap map1, float* out,
#endif
 int x, int y, int c, int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  uint32_t sb = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#ifdef MAPGLOBAL
    const CUtensorMap* mp = gmap;
#else
    const CUtensorMap* mp = &map1;
#endif
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(BOXW * 4 * BOXH) : "memory");
#if RANK == 4
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(sb), "l"(mp), "r"(x), "r"(y), "r"(c), "r"(0), "r"(b) : "memory");
#elif RANK == 3
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(sb), "l"(mp), "r"(x), "r"(y), "r"(c), "r"(b) : "memory");
#else
#if MODE == 0
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(sb), "l"(mp), "r"(x), "r"(y + 64 * c), "r"(b) : "memory");
#elif MODE == 1
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(sb), "l"(mp), "r"(x), "r"(y + 64 * c), "r"(b), "l"(0x1000000000000000ull) : "memory");
#elif MODE == 3
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(mp), "r"(x), "r"(y + 64 * c) : "memory");
    asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" :: "r"(b), "r"(BOXW * 4 * BOXH) : "memory");
#elif MODE == 4
    asm volatile("prefetch.tensormap [%0];" :: "l"(mp) : "memory");
    asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" :: "r"(b), "r"(BOXW * 4 * BOXH) : "memory");
#endif
#endif
  }
  uint32_t ok = 0; int spins = 0;
  while (!ok && spins < 200000) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(b), "r"(0), "r"(2000u) : "memory");
    ++spins;
  }
  if (threadIdx.x == 0) status[0] = ok ? spins : -1;
  const float* s = reinterpret_cast<const float*>(smem + (sb - (uint32_t)__cvta_generic_to_shared(smem)));
  for (int i = threadIdx.x; i < BOXW * BOXH; i += blockDim.x) out[i] = ok ? s[i] : -7.f;
}
int main() {
  const int N = 1, C = 4, Hh = 64, W = 128;
  std::vector<float> h(N * C * Hh * W);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o; int* st;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 64 * 40 * 4); cudaMalloc(&st, 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
#ifdef LINKED
  enc = &cuTensorMapEncodeTiled;
#endif
  alignas(64) CUtensorMap map;
  const cuuint64_t dims4[4] = {W, Hh, C, N};
  const cuuint64_t dims2[2] = {W, (cuuint64_t)Hh * C};
  const cuuint64_t strides[3] = {W * 4ull, (cuuint64_t)W * Hh * 4ull, (cuuint64_t)W * Hh * C * 4ull};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const cuuint32_t box[4] = {BOXW, BOXH, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, RANK, d, RANK == 2 ? dims2 : dims4, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, SWZ,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  { const unsigned long long* w = reinterpret_cast<const unsigned long long*>(&map);
    for (int i = 0; i < 16; ++i) printf("%016llx%c", w[i], (i % 4 == 3) ? '\n' : ' '); printf("d=%p\n", (void*)d); }
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 40 * 4 + 1024);
#ifdef MAPGLOBAL
  CUtensorMap* gm; cudaMalloc(&gm, 128); cudaMemcpy(gm, &map, 128, cudaMemcpyHostToDevice);
  k<<<1, 256, 64 * 40 * 4 + 1024>>>(gm, o, 5, 3, 2, st);
#else
  k<<<1, 256, 64 * 40 * 4 + 1024>>>(map, o, 5, 3, 2, st);
#endif
  cudaError_t e = cudaDeviceSynchronize();
  int s = 0; float v[3] = {0, 0, 0};
  cudaMemcpy(&s, st, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(v, o, 12, cudaMemcpyDeviceToHost);
  float expect = (float)(2 * Hh * W + 3 * W + 5);
  printf("rank %d box %dx%d: encode=%d sync=%s status=%d out[0..2]=%g %g %g expect %g\n", RANK, BOXW, BOXH, (int)r, cudaGetErrorString(e), s, v[0], v[1], v[2], expect);
  return 0;
}
