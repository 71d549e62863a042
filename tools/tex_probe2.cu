// Probe 2: do the texture pipe (tex2Dgather) and the LSU/shared-memory pipe (conflicted LDS +
// LDGSTS staging) of an SM run side by side?  Same thread does, per plane pair, (a) one texture
// gather + blend + store and/or (b) a staged-plane emulation: 3 x 16-byte cp.async into a ring of
// buffers and 4 LDS at flow-dependent addresses + blend + store.  Times: (a) only, (b) only, both.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <math.h>
constexpr int C = 64, H = 1088, W = 1920;
constexpr int kBuf = 64 * 40;   // floats per staged buffer
template <int MODE>   // 1 = tex, 2 = lsu, 3 = both
__global__ void __launch_bounds__(256, 4) k(cudaTextureObject_t t, const float* __restrict__ im,
                                           const float* __restrict__ flow, float* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int w = blockIdx.x * 32 + lane;
  const int h0 = blockIdx.y * 16 + wid;
  float wgt[2][4]; float gx[2], gy[2]; int idx[2]; float* po[2];
#pragma unroll
  for (int k2 = 0; k2 < 2; ++k2) {
    const int h = min(h0 + 8 * k2, H - 1), ww = min(w, W - 1);
    const float fx = flow[h * W + ww], fy = flow[H * W + h * W + ww];
    float ix = fminf(fmaxf(ww + fx, 0.f), W - 1.f), iy = fminf(fmaxf(h + fy, 0.f), H - 1.f);
    const float x0 = floorf(ix), y0 = floorf(iy);
    const float bx = ix - x0, by = iy - y0, ax = 1.f - bx, ay = 1.f - by;
    wgt[k2][0] = ax * ay; wgt[k2][1] = bx * ay; wgt[k2][2] = ax * by; wgt[k2][3] = bx * by;
    gx[k2] = x0 + 1.f; gy[k2] = y0 + 1.f;
    idx[k2] = (((int)y0 - (int)blockIdx.y * 16 + 12) & 31) * 64 + (((int)x0 - (int)blockIdx.x * 32 + 12) & 62);
    po[k2] = out + (size_t)h * W + ww;
  }
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
  const float* src = im + (size_t)(blockIdx.y * 16) * W + min((int)blockIdx.x * 32, W - 64) + (threadIdx.x & 7) * 4 + (size_t)(threadIdx.x >> 3) * W;
  const unsigned dst = ((threadIdx.x >> 3) * 64 + (threadIdx.x & 7) * 4) * 4;
#pragma unroll 1
  for (int c = 0; c < C / 2; ++c) {
    if (MODE & 2) {
      const unsigned b = sbase + (c & 3) * kBuf * 4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(b + dst), "l"(src) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(b + dst + 32u * 4), "l"(src + 32) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(b + dst + 8u * 64u * 4), "l"(src + 8 * W) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 2;" ::: "memory");
      __syncthreads();
      const float* bb = sm + ((c + 1) & 3) * kBuf;
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        float acc = bb[idx[k2]] * wgt[k2][0];
        acc = fmaf(bb[idx[k2] + 1], wgt[k2][1], acc);
        acc = fmaf(bb[idx[k2] + 64], wgt[k2][2], acc);
        acc = fmaf(bb[idx[k2] + 65], wgt[k2][3], acc);
        __stcs(po[k2] + (size_t)c * H * W, acc);
      }
      src += (size_t)H * W;
    }
    if (MODE & 1) {
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        const float4 q = tex2Dgather<float4>(t, gx[k2], gy[k2] + (float)(c * H), 0);
        float acc = q.w * wgt[k2][0];
        acc = fmaf(q.z, wgt[k2][1], acc);
        acc = fmaf(q.x, wgt[k2][2], acc);
        acc = fmaf(q.y, wgt[k2][3], acc);
        __stcs(po[k2] + (size_t)(C / 2 + c) * H * W, acc);
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
template <int MODE>
float run(cudaTextureObject_t t, const float* im, const float* flow, float* out) {
  dim3 grid((W + 31) / 32, (H + 15) / 16);
  const int smem = 4 * kBuf * 4;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) k<MODE><<<grid, 256, smem>>>(t, im, flow, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) printf("mode %d: %s\n", MODE, cudaGetErrorString(e));
  e = cudaGetLastError();
  if (e != cudaSuccess) printf("mode %d launch: %s\n", MODE, cudaGetErrorString(e));
  cudaEventRecord(a);
  for (int i = 0; i < 10; ++i) k<MODE><<<grid, 256, smem>>>(t, im, flow, out);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / 10 * 1e3f;
}
int main(int argc, char** argv) {
  const size_t n = (size_t)C * H * W;
  float *im, *out, *flow;
  cudaMalloc(&im, n * 4); cudaMalloc(&out, n * 4); cudaMalloc(&flow, 2ull * H * W * 4);
  cudaMemset(im, 0, n * 4);
  std::vector<float> hf(2ull * H * W);
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  for (int h = 0; h < H; ++h) for (int w = 0; w < W; ++w) {
    float fx, fy;
    if (mode == 0) { fx = 4.f * sinf(0.25f * w + 0.11f * h) + 2.f * sinf(0.05f * w - 0.11f * h);
                     fy = 4.f * cosf(0.22f * h - 0.09f * w) + 2.f * sinf(0.07f * h + 0.03f * w); }
    else { fx = 4.f * sinf(0.02f * w + 0.01f * h); fy = 4.f * cosf(0.015f * h - 0.01f * w); }
    hf[(size_t)h * W + w] = fx; hf[(size_t)H * W + (size_t)h * W + w] = fy;
  }
  cudaMemcpy(flow, hf.data(), hf.size() * 4, cudaMemcpyHostToDevice);
  cudaResourceDesc rd = {};
  rd.resType = cudaResourceTypePitch2D;
  rd.res.pitch2D.devPtr = im + (size_t)(C / 2) * H * W;
  rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
  rd.res.pitch2D.width = W; rd.res.pitch2D.height = (size_t)(C / 2) * H; rd.res.pitch2D.pitchInBytes = (size_t)W * 4;
  cudaTextureDesc td = {};
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
  cudaTextureObject_t t;
  printf("tex: %s\n", cudaGetErrorString(cudaCreateTextureObject(&t, &rd, &td, nullptr)));
  const float t1 = run<1>(t, im, flow, out), t2 = run<2>(t, im, flow, out), t3 = run<3>(t, im, flow, out);
  printf("flow %d: tex-only (32 planes) %.1f us | lsu-only (32 planes) %.1f us | both (64 planes) %.1f us | sum %.1f\n",
         mode, t1, t2, t3, t1 + t2);
  return 0;
}
