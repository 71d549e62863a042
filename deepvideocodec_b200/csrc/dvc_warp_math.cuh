// dvc_warp_math.cuh -- the coordinate pipeline of flow_warp, shared by the
// stand-alone warp kernels (dvc_warp.cu) and the fused warp + 3x3 conv
// (dvc_warp_conv.cu).  Bit-faithful replay of
//   /root/reference/dmc/models/layers.py:175-198 (torch_warp / flow_warp)
// as PyTorch-CUDA eager executes it.
#pragma once
#include "dvc_common.cuh"

namespace dvc {

// ---------------------------------------------------------------------------
// coordinate pipeline (bit-faithful replay of linspace + div + add +
// grid_sampler_compute_source_index + clip), see header comment of each step
// ---------------------------------------------------------------------------
struct WarpGeom {
  int H, W;
  float step_x, step_y;  // fl32(2 / (S-1)): torch.linspace step
  float norm_x, norm_y;  // default: fl32(1 / fl32((S-1)/2)); IEEE mode: fl32((S-1)/2)
  float wm1, hm1;        // (float)(S-1)
  int ieee_div;
};

// torch.linspace(-1, 1, S)[j]: start + step*j below the midpoint, end -
// step*(S-1-j) above it, each contracted to ONE fused multiply-add by both
// ATen back ends (RangeFactories); the unfused form mismatches ~40% of entries.
__device__ __forceinline__ float linspace_pm1(int j, int S, float step) {
  return (j < (S >> 1)) ? fmaf(step, (float)j, -1.0f)
                        : fmaf(-step, (float)(S - 1 - j), 1.0f);
}

struct Taps {
  int x0, y0;            // north-west tap
  int dx, dy;            // 1 if the east / south neighbour is inside, else 0
  float nw, ne, sw, se;  // bilinear weights
};

__device__ __forceinline__ float source_index(float base, float f, float norm,
                                              float sm1, int ieee_div) {
  // layers.py:185-186  flow / ((S-1)/2): CUDA eager multiplies by the fp32
  // reciprocal of the python scalar, CPU eager divides.
  float fn = ieee_div ? div_rn(f, norm) : mul_rn(f, norm);
  float c = add_rn(base, fn);                                   // layers.py:188
  // GridSampler.cuh grid_sampler_unnormalize(align_corners=True):
  //   ((coord + 1) / 2) * (size - 1)
  float i = mul_rn(mul_rn(add_rn(c, 1.0f), 0.5f), sm1);
  // clip_coordinates: min(size-1, max(i, 0)); NaN -> 0 through fmaxf
  return fminf(sm1, fmaxf(i, 0.0f));
}

__device__ __forceinline__ Taps make_taps(const WarpGeom& g, int h, int w, float fx,
                                          float fy) {
  float ix = source_index(linspace_pm1(w, g.W, g.step_x), fx, g.norm_x, g.wm1, g.ieee_div);
  float iy = source_index(linspace_pm1(h, g.H, g.step_y), fy, g.norm_y, g.hm1, g.ieee_div);
  float x0f = floorf(ix), y0f = floorf(iy);
  float x1f = x0f + 1.0f, y1f = y0f + 1.0f;
  Taps t;
  t.x0 = (int)x0f;
  t.y0 = (int)y0f;
  t.dx = (t.x0 + 1 < g.W) ? 1 : 0;  // within_bounds_2d of the east taps
  t.dy = (t.y0 + 1 < g.H) ? 1 : 0;
  float ax = sub_rn(x1f, ix), bx = sub_rn(ix, x0f);
  float ay = sub_rn(y1f, iy), by = sub_rn(iy, y0f);
  t.nw = mul_rn(ax, ay);
  t.ne = mul_rn(bx, ay);
  t.sw = mul_rn(ax, by);
  t.se = mul_rn(bx, by);
  return t;
}

// out_acc = 0; out_acc += v*w for nw, ne, sw, se in that order (each `+=` is
// one FFMA in ATen's kernel as compiled by nvcc).
__device__ __forceinline__ float blend(float vnw, float vne, float vsw, float vse,
                                       const Taps& t) {
  float acc = mul_rn(vnw, t.nw);
  acc = fmaf(vne, t.ne, acc);
  acc = fmaf(vsw, t.sw, acc);
  acc = fmaf(vse, t.se, acc);
  return acc;
}

// ---------------------------------------------------------------------------
// flow fetch, optionally through the reference's flow pyramid
//   level 1: bilineardownsacling(mv) / 2          (video_model.py:499)
//   level 2: bilineardownsacling(level 1) / 2     (video_model.py:500)
// evaluated per sample with exactly the arithmetic of flow_pyramid_kernel, so
// warping with flow_level = k is bit-identical to warping with the
// materialised mv2 / mv3.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float down2_at(const float* __restrict__ q, long long sh,
                                          long long sw) {
  const float a = __ldg(q), b = __ldg(q + sw);
  const float c = __ldg(q + sh), d = __ldg(q + sh + sw);
  return mul_rn(mul_rn(add_rn(add_rn(a, b), add_rn(c, d)), 0.25f), 0.5f);
}

__device__ __forceinline__ float fetch_flow(const float* __restrict__ plane, long long sh,
                                            long long sw, int h, int w, int level) {
  if (level == 0) return __ldg(plane + h * sh + w * sw);
  if (level == 1) return down2_at(plane + (2 * h) * sh + (2 * w) * sw, sh, sw);
  const float* q = plane + (4 * h) * sh + (4 * w) * sw;
  const float m00 = down2_at(q, sh, sw), m01 = down2_at(q + 2 * sw, sh, sw);
  const float m10 = down2_at(q + 2 * sh, sh, sw), m11 = down2_at(q + 2 * sh + 2 * sw, sh, sw);
  return mul_rn(mul_rn(add_rn(add_rn(m00, m01), add_rn(m10, m11)), 0.25f), 0.5f);
}

// ---------------------------------------------------------------------------
// float4 (4-channel group) helpers of the channels_last paths
// ---------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

__device__ __forceinline__ float4 blend4(const float4& a, const float4& b, const float4& c,
                                         const float4& d, const float4& w) {
  float4 o;
  o.x = fmaf(d.x, w.w, fmaf(c.x, w.z, fmaf(b.x, w.y, mul_rn(a.x, w.x))));
  o.y = fmaf(d.y, w.w, fmaf(c.y, w.z, fmaf(b.y, w.y, mul_rn(a.y, w.x))));
  o.z = fmaf(d.z, w.w, fmaf(c.z, w.z, fmaf(b.z, w.y, mul_rn(a.z, w.x))));
  o.w = fmaf(d.w, w.w, fmaf(c.w, w.z, fmaf(b.w, w.y, mul_rn(a.w, w.x))));
  return o;
}

constexpr unsigned kEastIn = 1u << 30, kSouthIn = 1u << 31, kOffMask = (1u << 30) - 1u;

// host side: constants of the pipeline for an H x W image
static inline int fill_geom(WarpGeom& g, int64_t H, int64_t W, int flags) {
  g.H = (int)H;
  g.W = (int)W;
  // ATen RangeFactories: step = (end - start) / (steps - 1) in fp32
  g.step_x = 2.0f / (float)(W - 1);
  g.step_y = 2.0f / (float)(H - 1);
  // layers.py:185-186: the divisor is the python double (S - 1.0) / 2.0,
  // converted to fp32 when it meets the fp32 tensor.
  const float half_w = (float)(((double)W - 1.0) / 2.0);
  const float half_h = (float)(((double)H - 1.0) / 2.0);
  g.ieee_div = (flags & DVC_WARP_IEEE_DIV) ? 1 : 0;
  g.norm_x = g.ieee_div ? half_w : 1.0f / half_w;
  g.norm_y = g.ieee_div ? half_h : 1.0f / half_h;
  g.wm1 = (float)(W - 1);
  g.hm1 = (float)(H - 1);
  return DVC_OK;
}


}  // namespace dvc
