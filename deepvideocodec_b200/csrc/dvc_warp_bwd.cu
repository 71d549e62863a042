// dvc_warp_bwd.cu -- backward of the flow warp and of the 2x bilinear downscale.
//
// Replaces the autograd of
//   flow_warp            /root/reference/dmc/models/layers.py:175-198
//     (ATen grid_sampler_2d_backward, bilinear/border/align_corners=True:
//      GridSampler.cuh grid_sampler_compute_source_index_set_grad -- gradient
//      multiplier (S-1)/2, ZERO through the border clip when ix <= 0 or
//      ix >= S-1; grad_input scattered with atomics) chained through
//      grid = base + flow / ((S-1)/2)                      (layers.py:185-188)
//   bilineardownsacling  /root/reference/dmc/models/layers.py:201-206
// first used by the reference at dmc/train.py:301 (loss.backward()).
//
// Layouts as in the forward: channels_last float4 path (vector atomics,
// red.global.add.v4.f32 on sm_90+) and a strided path.
#include "dvc_common.cuh"

namespace dvc {

struct BwdGeom {
  int H, W;
  float step_x, step_y, norm_x, norm_y, wm1, hm1;
  float gmul_x, gmul_y;   // (S-1)/2
  float back_x, back_y;   // d(flow/((S-1)/2))/dflow as eager evaluates it (reciprocal multiply)
  int ieee_div;
};

struct BwdP {
  const float* gout;
  const float* im;
  const float* flow;
  float* gim;
  float* gflow;
  BwdGeom g;
  int N, C;
  long long go_n, go_c, go_h, go_w;
  long long im_n, im_c, im_h, im_w;
  long long fl_n, fl_c, fl_h, fl_w;
  long long gi_n, gi_c, gi_h, gi_w;
  long long gf_n, gf_c, gf_h, gf_w;
  int c4, ppw;
};

__device__ __forceinline__ float lin_pm1(int j, int S, float step) {
  return (j < (S >> 1)) ? fmaf(step, (float)j, -1.0f) : fmaf(-step, (float)(S - 1 - j), 1.0f);
}

struct BTaps {
  int x0, y0, dx, dy;
  float ax, bx, ay, by;   // x1-ix, ix-x0, y1-iy, iy-y0
  float mx, my;           // d(ix)/d(flow_x), d(iy)/d(flow_y) incl. clip gradient
};

__device__ __forceinline__ float src_index_grad(float base, float f, float norm, float sm1,
                                                int ieee_div, float gmul, float back,
                                                float& mult) {
  float fn = ieee_div ? div_rn(f, norm) : mul_rn(f, norm);
  float c = add_rn(base, fn);
  float i = mul_rn(mul_rn(add_rn(c, 1.0f), 0.5f), sm1);
  // clip_coordinates_set_grad: gradient 0 on and outside the border
  float clip = 1.0f;
  if (!(i > 0.0f)) { i = 0.0f; clip = 0.0f; }        // also catches NaN like fmaxf(NaN,0)
  else if (i >= sm1) { i = sm1; clip = 0.0f; }
  mult = gmul * clip * back;
  return i;
}

__device__ __forceinline__ BTaps make_btaps(const BwdGeom& g, int h, int w, float fx, float fy) {
  BTaps t;
  float ix = src_index_grad(lin_pm1(w, g.W, g.step_x), fx, g.norm_x, g.wm1, g.ieee_div,
                            g.gmul_x, g.back_x, t.mx);
  float iy = src_index_grad(lin_pm1(h, g.H, g.step_y), fy, g.norm_y, g.hm1, g.ieee_div,
                            g.gmul_y, g.back_y, t.my);
  float x0f = floorf(ix), y0f = floorf(iy);
  t.x0 = (int)x0f;
  t.y0 = (int)y0f;
  t.dx = (t.x0 + 1 < g.W) ? 1 : 0;
  t.dy = (t.y0 + 1 < g.H) ? 1 : 0;
  t.ax = (x0f + 1.0f) - ix;
  t.bx = ix - x0f;
  t.ay = (y0f + 1.0f) - iy;
  t.by = iy - y0f;
  return t;
}

// ---- strided: one thread per pixel ------------------------------------------
__global__ void __launch_bounds__(256) warp_bwd_strided_kernel(const BwdP p) {
  const int w = blockIdx.x * 32 + (threadIdx.x & 31);
  const int h = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int n = blockIdx.z;
  if (w >= p.g.W || h >= p.g.H) return;
  const float* fl = p.flow + n * p.fl_n + h * p.fl_h + w * p.fl_w;
  const BTaps t = make_btaps(p.g, h, w, __ldg(fl), __ldg(fl + p.fl_c));
  const float nw = t.ax * t.ay, ne = t.bx * t.ay, sw = t.ax * t.by, se = t.bx * t.by;
  const long long o_nw = t.y0 * p.im_h + t.x0 * p.im_w;
  const long long g_nw = t.y0 * p.gi_h + t.x0 * p.gi_w;
  const float* go = p.gout + n * p.go_n + h * p.go_h + w * p.go_w;
  const float* imn = p.im + n * p.im_n;
  float* gin = p.gim ? p.gim + n * p.gi_n : nullptr;
  float gix = 0.f, giy = 0.f;
  // The channel loop is latency bound (ncu r02: long-scoreboard stalls 46 per issue, issue slots
  // 22 %): four channels per trip, all of their loads issued before the first use.
  constexpr int kU = 4;
  const long long o_e = t.dx ? p.im_w : 0, o_s = t.dy ? p.im_h : 0;
  int c = 0;
  for (; c + kU <= p.C; c += kU) {
    float g[kU], vnw[kU], vne[kU], vsw[kU], vse[kU];
#pragma unroll
    for (int k = 0; k < kU; ++k) g[k] = __ldg(go + (c + k) * p.go_c);
    if (p.gflow) {
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        const float* q = imn + (c + k) * p.im_c + o_nw;
        vnw[k] = __ldg(q);
        vne[k] = __ldg(q + o_e);
        vsw[k] = __ldg(q + o_s);
        vse[k] = __ldg(q + o_s + o_e);
      }
    }
    if (gin) {
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        float* q = gin + (c + k) * p.gi_c + g_nw;
        atomicAdd(q, nw * g[k]);
        if (t.dx) atomicAdd(q + p.gi_w, ne * g[k]);
        if (t.dy) atomicAdd(q + p.gi_h, sw * g[k]);
        if (t.dx && t.dy) atomicAdd(q + p.gi_h + p.gi_w, se * g[k]);
      }
    }
    if (p.gflow) {
#pragma unroll
      for (int k = 0; k < kU; ++k) {      // same per-channel expressions and order as the tail loop
        const float a = vnw[k];
        const float b = t.dx ? vne[k] : 0.f;
        const float cc = t.dy ? vsw[k] : 0.f;
        const float d = (t.dx && t.dy) ? vse[k] : 0.f;
        gix += ((b - a) * t.ay + (d - cc) * t.by) * g[k];
        giy += ((cc - a) * t.ax + (d - b) * t.bx) * g[k];
      }
    }
  }
  for (; c < p.C; ++c) {
    const float g = __ldg(go + c * p.go_c);
    if (gin) {
      float* q = gin + c * p.gi_c + g_nw;
      atomicAdd(q, nw * g);
      if (t.dx) atomicAdd(q + p.gi_w, ne * g);
      if (t.dy) atomicAdd(q + p.gi_h, sw * g);
      if (t.dx && t.dy) atomicAdd(q + p.gi_h + p.gi_w, se * g);
    }
    if (p.gflow) {
      const float* q = imn + c * p.im_c + o_nw;
      const float vnw = __ldg(q);
      const float vne = t.dx ? __ldg(q + p.im_w) : 0.f;
      const float vsw = t.dy ? __ldg(q + p.im_h) : 0.f;
      const float vse = (t.dx && t.dy) ? __ldg(q + p.im_h + p.im_w) : 0.f;
      gix += ((vne - vnw) * t.ay + (vse - vsw) * t.by) * g;
      giy += ((vsw - vnw) * t.ax + (vse - vne) * t.bx) * g;
    }
  }
  if (p.gflow) {
    float* gf = p.gflow + n * p.gf_n + h * p.gf_h + w * p.gf_w;
    gf[0] = t.mx * gix;
    gf[p.gf_c] = t.my * giy;
  }
}

// ---- channels_last float4: lane = (pixel column, 4-channel group) ------------
__global__ void __launch_bounds__(256) warp_bwd_vec4_kernel(const BwdP p) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int px = lane / p.c4, grp = lane - px * p.c4;
  const int w = (blockIdx.x * 8 + wid) * p.ppw + px;
  const int h = blockIdx.y;
  const int n = blockIdx.z;
  const bool active = w < p.g.W;
  const int wc = active ? w : p.g.W - 1;
  const float* fl = p.flow + n * p.fl_n + h * p.fl_h + wc * p.fl_w;
  const BTaps t = make_btaps(p.g, h, wc, __ldg(fl), __ldg(fl + p.fl_c));
  const float nw = t.ax * t.ay, ne = t.bx * t.ay, sw = t.ax * t.by, se = t.bx * t.by;
  float gix = 0.f, giy = 0.f;
  if (active) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(
                               p.gout + n * p.go_n + h * p.go_h + w * p.go_w) + grp);
    if (p.gim) {
      float4* q = reinterpret_cast<float4*>(p.gim + n * p.gi_n + t.y0 * p.gi_h + t.x0 * p.gi_w) + grp;
      const long long e = p.gi_w >> 2, s = p.gi_h >> 2;
      atomicAdd(q, make_float4(nw * g.x, nw * g.y, nw * g.z, nw * g.w));
      if (t.dx) atomicAdd(q + e, make_float4(ne * g.x, ne * g.y, ne * g.z, ne * g.w));
      if (t.dy) atomicAdd(q + s, make_float4(sw * g.x, sw * g.y, sw * g.z, sw * g.w));
      if (t.dx && t.dy) atomicAdd(q + s + e, make_float4(se * g.x, se * g.y, se * g.z, se * g.w));
    }
    if (p.gflow) {
      const float4* q = reinterpret_cast<const float4*>(
                            p.im + n * p.im_n + t.y0 * p.im_h + t.x0 * p.im_w) + grp;
      const long long e = p.im_w >> 2, s = p.im_h >> 2;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 a = __ldg(q);
      const float4 b = t.dx ? __ldg(q + e) : z;
      const float4 c = t.dy ? __ldg(q + s) : z;
      const float4 d = (t.dx && t.dy) ? __ldg(q + s + e) : z;
      gix = ((b.x - a.x) * t.ay + (d.x - c.x) * t.by) * g.x + ((b.y - a.y) * t.ay + (d.y - c.y) * t.by) * g.y +
            ((b.z - a.z) * t.ay + (d.z - c.z) * t.by) * g.z + ((b.w - a.w) * t.ay + (d.w - c.w) * t.by) * g.w;
      giy = ((c.x - a.x) * t.ax + (d.x - b.x) * t.bx) * g.x + ((c.y - a.y) * t.ax + (d.y - b.y) * t.bx) * g.y +
            ((c.z - a.z) * t.ax + (d.z - b.z) * t.bx) * g.z + ((c.w - a.w) * t.ax + (d.w - b.w) * t.bx) * g.w;
    }
  }
  if (p.gflow) {
    // sum over the c4 lanes of this pixel (c4 is a power of two dividing 32)
    for (int o = p.c4 >> 1; o > 0; o >>= 1) {
      gix += __shfl_xor_sync(0xffffffffu, gix, o);
      giy += __shfl_xor_sync(0xffffffffu, giy, o);
    }
    if (active && grp == 0) {
      float* gf = p.gflow + n * p.gf_n + h * p.gf_h + w * p.gf_w;
      gf[0] = t.mx * gix;
      gf[p.gf_c] = t.my * giy;
    }
  }
}

// ---- bilinear 2x downscale backward ------------------------------------------
struct DownBP {
  const float* gy;
  float* gx;
  int N, C, H, W, Ho, Wo;
  long long ys_n, ys_c, ys_h, ys_w, xs_n, xs_c, xs_h, xs_w;
  float scale_h, scale_w, post;
  int exact2;
};

__device__ __forceinline__ float area_src_b(float scale, int dst) {
  float s = fmaf(scale, (float)dst + 0.5f, -0.5f);
  return s < 0.f ? 0.f : s;
}

// even sizes: each input pixel belongs to exactly one output pixel, weight 1/4
__global__ void __launch_bounds__(256) down2_bwd_exact_kernel(const DownBP p) {
  const long long total = (long long)p.N * p.C * p.H * p.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % p.W);
    long long r = i / p.W;
    const int h = (int)(r % p.H);
    r /= p.H;
    const int c = (int)(r % p.C);
    const int n = (int)(r / p.C);
    const float g = __ldg(p.gy + n * p.ys_n + c * p.ys_c + (h >> 1) * p.ys_h + (w >> 1) * p.ys_w);
    p.gx[n * p.xs_n + c * p.xs_c + h * p.xs_h + w * p.xs_w] = mul_rn(mul_rn(g, p.post), 0.25f);
  }
}

// general sizes: scatter (grad_x must be zero-filled)
__global__ void __launch_bounds__(256) down2_bwd_general_kernel(const DownBP p) {
  const long long total = (long long)p.N * p.C * p.Ho * p.Wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int wo = (int)(i % p.Wo);
    long long r = i / p.Wo;
    const int ho = (int)(r % p.Ho);
    r /= p.Ho;
    const int c = (int)(r % p.C);
    const int n = (int)(r / p.C);
    const float g = __ldg(p.gy + n * p.ys_n + c * p.ys_c + ho * p.ys_h + wo * p.ys_w) * p.post;
    const float h1r = area_src_b(p.scale_h, ho), w1r = area_src_b(p.scale_w, wo);
    const int h1 = (int)h1r, w1 = (int)w1r;
    const int h1p = (h1 < p.H - 1) ? 1 : 0, w1p = (w1 < p.W - 1) ? 1 : 0;
    const float h1l = h1r - (float)h1, h0l = 1.f - h1l;
    const float w1l = w1r - (float)w1, w0l = 1.f - w1l;
    float* q = p.gx + n * p.xs_n + c * p.xs_c + h1 * p.xs_h + w1 * p.xs_w;
    atomicAdd(q, h0l * w0l * g);
    atomicAdd(q + w1p * p.xs_w, h0l * w1l * g);
    atomicAdd(q + h1p * p.xs_h, h1l * w0l * g);
    atomicAdd(q + h1p * p.xs_h + w1p * p.xs_w, h1l * w1l * g);
  }
}

}  // namespace dvc

using namespace dvc;

extern "C" {

int dvc_flow_warp_bwd(const float* grad_out, const float* im, const float* flow, float* grad_im,
                      float* grad_flow, int64_t N, int64_t C, int64_t H, int64_t W,
                      const int64_t gout_st[4], const int64_t im_st[4], const int64_t flow_st[4],
                      const int64_t gim_st[4], const int64_t gflow_st[4], int flags,
                      dvc_stream_t stream) {
  DVC_REQUIRE(grad_out && im && flow && gout_st && im_st && flow_st, "flow_warp_bwd: null input");
  DVC_REQUIRE(grad_im || grad_flow, "flow_warp_bwd: nothing to compute");
  DVC_REQUIRE(!grad_im || gim_st, "flow_warp_bwd: grad_im without strides");
  DVC_REQUIRE(!grad_flow || gflow_st, "flow_warp_bwd: grad_flow without strides");
  DVC_REQUIRE(N > 0 && C > 0 && H >= 2 && W >= 2, "flow_warp_bwd: bad extents");
  DVC_REQUIRE(N < 65536 && H < 65536 * 8, "flow_warp_bwd: extent too large");
  BwdP p;
  p.gout = grad_out; p.im = im; p.flow = flow; p.gim = grad_im; p.gflow = grad_flow;
  p.N = (int)N; p.C = (int)C;
  p.g.H = (int)H; p.g.W = (int)W;
  p.g.step_x = 2.0f / (float)(W - 1);
  p.g.step_y = 2.0f / (float)(H - 1);
  const float half_w = (float)(((double)W - 1.0) / 2.0);
  const float half_h = (float)(((double)H - 1.0) / 2.0);
  p.g.ieee_div = (flags & DVC_WARP_IEEE_DIV) ? 1 : 0;
  p.g.norm_x = p.g.ieee_div ? half_w : 1.0f / half_w;
  p.g.norm_y = p.g.ieee_div ? half_h : 1.0f / half_h;
  p.g.wm1 = (float)(W - 1);
  p.g.hm1 = (float)(H - 1);
  p.g.gmul_x = (float)(W - 1) / 2.0f;
  p.g.gmul_y = (float)(H - 1) / 2.0f;
  p.g.back_x = 1.0f / half_w;
  p.g.back_y = 1.0f / half_h;
  const Strides4 sg = make_strides(gout_st), si = make_strides(im_st), sf = make_strides(flow_st);
  p.go_n = sg.n; p.go_c = sg.c; p.go_h = sg.h; p.go_w = sg.w;
  p.im_n = si.n; p.im_c = si.c; p.im_h = si.h; p.im_w = si.w;
  p.fl_n = sf.n; p.fl_c = sf.c; p.fl_h = sf.h; p.fl_w = sf.w;
  Strides4 sgi = si, sgf = sf;
  if (grad_im) sgi = make_strides(gim_st);
  if (grad_flow) sgf = make_strides(gflow_st);
  p.gi_n = sgi.n; p.gi_c = sgi.c; p.gi_h = sgi.h; p.gi_w = sgi.w;
  p.gf_n = sgf.n; p.gf_c = sgf.c; p.gf_h = sgf.h; p.gf_w = sgf.w;
  const int64_t c4 = C / 4;
  const bool vec = nhwc_vec4_ok(grad_out, sg, C) && nhwc_vec4_ok(im, si, C) &&
                   (!grad_im || nhwc_vec4_ok(grad_im, sgi, C)) && c4 >= 1 && c4 <= 32 &&
                   (32 % c4) == 0;
  if (vec) {
    p.c4 = (int)c4;
    p.ppw = 32 / p.c4;
    dim3 grid((unsigned)((W + 8 * p.ppw - 1) / (8 * p.ppw)), (unsigned)H, (unsigned)N);
    DVC_REQUIRE(H <= 65535, "flow_warp_bwd: H > 65535 on the float4 path");
    warp_bwd_vec4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("warp_bwd_vec4_kernel");
  }
  p.c4 = 0; p.ppw = 0;
  dim3 grid((unsigned)((W + 31) / 32), (unsigned)((H + 7) / 8), (unsigned)N);
  warp_bwd_strided_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("warp_bwd_strided_kernel");
}

int dvc_bilinear_down2_bwd(const float* grad_y, float* grad_x, int64_t N, int64_t C, int64_t H,
                           int64_t W, const int64_t gy_st[4], const int64_t gx_st[4],
                           float post_scale, dvc_stream_t stream) {
  DVC_REQUIRE(grad_y && grad_x && gy_st && gx_st, "bilinear_down2_bwd: null pointer");
  DVC_REQUIRE(N > 0 && C > 0 && H >= 2 && W >= 2, "bilinear_down2_bwd: needs H,W >= 2");
  DownBP p;
  p.gy = grad_y; p.gx = grad_x;
  p.N = (int)N; p.C = (int)C; p.H = (int)H; p.W = (int)W;
  p.Ho = (int)(H / 2); p.Wo = (int)(W / 2);
  p.ys_n = gy_st[0]; p.ys_c = gy_st[1]; p.ys_h = gy_st[2]; p.ys_w = gy_st[3];
  p.xs_n = gx_st[0]; p.xs_c = gx_st[1]; p.xs_h = gx_st[2]; p.xs_w = gx_st[3];
  p.scale_h = (float)H / (float)p.Ho;
  p.scale_w = (float)W / (float)p.Wo;
  p.post = post_scale;
  p.exact2 = ((H % 2) == 0 && (W % 2) == 0) ? 1 : 0;
  const int sms = sm_count();
  if (p.exact2) {
    const long long total = (long long)N * C * H * W;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    down2_bwd_exact_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("down2_bwd_exact_kernel");
  }
  // odd sizes: the caller zero-fills grad_x (documented in the header)
  const long long total = (long long)N * C * p.Ho * p.Wo;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
  down2_bwd_general_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("down2_bwd_general_kernel");
}

}  // extern "C"
