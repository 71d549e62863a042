// dvc_entropy_bwd.cu -- backward of the dual prior, Gaussian conditional and
// entropy bottleneck kernels (training, /root/reference/dmc/train.py:301).
//
// Gradient rules replayed from the reference graph:
//   quantize_ste (utils.py:149-152)        d round(x)/dx := 1
//   process_with_mask (video_model.py:161-167)
//        y_hat = STE((y - mu*m)*m) + mu*m  -> dy_hat/dy = m, dy_hat/dmu = 0
//   LowerBound (CompressAI)                pass iff (x >= bound) or (grad < 0)
//   GaussianConditional._likelihood        erfc' = -2/sqrt(pi) exp(-x^2); |.|' = sign
//        eval : outputs = round(x - mu) + mu  -> d/dx = 0, d/dmu = 1, and
//               values = outputs - mu         -> d values/dmu = 0
//        train: outputs = x + noise           -> d/dx = 1, d values/dmu = -1
//   EntropyBottleneck._likelihood          sigmoid' , tanh', softplus' (threshold 20)
// Gradients are analytic in fp32; parity with the reference autograd is judged
// at 1e-4 relative.
#include "dvc_common.cuh"

namespace dvc {

struct TSb { long long n, c, h, w; };
static inline TSb tsb(const int64_t s[4]) {
  TSb r;
  if (s) { r.n = s[0]; r.c = s[1]; r.h = s[2]; r.w = s[3]; }
  else { r.n = r.c = r.h = r.w = 0; }
  return r;
}
__device__ __forceinline__ long long offb(const TSb& s, int n, int c, int h, int w) {
  return n * s.n + c * s.c + h * s.h + w * s.w;
}

struct ShapeB { int N, C, H, W, E, c_fast; };
__device__ __forceinline__ void decode_b(const ShapeB& s, int i, int& c, int& h, int& w) {
  if (s.c_fast) { c = i % s.C; int r = i / s.C; w = r % s.W; h = r / s.W; }
  else { w = i % s.W; int r = i / s.W; h = r % s.H; c = r / s.H; }
}
static int make_shape_b(ShapeB& s, int64_t N, int64_t C, int64_t H, int64_t W,
                        const int64_t lead[4], const char* who) {
  if (!(N > 0 && C > 0 && H > 0 && W > 0)) return fail(DVC_ERR_INVALID_ARGUMENT, "%s: empty tensor", who);
  if (N > 65535) return fail(DVC_ERR_INVALID_ARGUMENT, "%s: N > 65535", who);
  const long long E = (long long)C * H * W;
  if (E >= 2147483647LL) return fail(DVC_ERR_INVALID_ARGUMENT, "%s: C*H*W too large", who);
  s.N = (int)N; s.C = (int)C; s.H = (int)H; s.W = (int)W; s.E = (int)E;
  s.c_fast = (lead && lead[1] == 1 && C > 1 && lead[3] != 1) ? 1 : 0;
  return DVC_OK;
}
static unsigned blocks_for(long long E, int per_thread) {
  long long b = (E + 256LL * per_thread - 1) / (256LL * per_thread);
  if (b < 1) b = 1;
  if (b > 65535) b = 65535;
  return (unsigned)b;
}

// d p / d(values, scale) of the Gaussian conditional, with both LowerBound rules
__device__ __forceinline__ void gc_grads(float outv, float mean, bool has_mean, float scale,
                                         float scale_bound, float lik_bound, float g_lik,
                                         bool has_glik, const double* g_logsum, int n,
                                         float& g_d, float& g_scale) {
  const float kNegRsqrt2 = -0.70710678118654752440f;
  const float kInvSqrt2Pi = 0.39894228040143267794f;
  const float d = has_mean ? outv - mean : outv;
  const float v = fabsf(d);
  const float s = (scale < scale_bound) ? scale_bound : scale;
  const float u = (0.5f - v) / s, l = (-0.5f - v) / s;
  const float praw = 0.5f * erfcf(kNegRsqrt2 * u) - 0.5f * erfcf(kNegRsqrt2 * l);
  const float p = (praw < lik_bound) ? lik_bound : praw;
  float gp = has_glik ? g_lik : 0.f;
  if (g_logsum) gp += (float)(g_logsum[n]) / p;      // d sum(ln p) / dp
  if (!(praw >= lik_bound || gp < 0.f)) gp = 0.f;    // LowerBound(likelihood)
  const float phi_u = kInvSqrt2Pi * expf(-0.5f * u * u);
  const float phi_l = kInvSqrt2Pi * expf(-0.5f * l * l);
  const float dp_dv = (phi_l - phi_u) / s;
  const float dp_ds = -(phi_u * u - phi_l * l) / s;
  const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
  g_d = gp * dp_dv * sgn;
  float gs = gp * dp_ds;
  if (!(scale >= scale_bound || gs < 0.f)) gs = 0.f;  // LowerBound(scale)
  g_scale = gs;
}

// ---------------------------------------------------------------------------
// module-level Gaussian conditional backward
// ---------------------------------------------------------------------------
struct GcBwdP {
  const float* g_lik; const double* g_logsum; const float* g_out;
  const float* x; const float* scales; const float* means; const float* noise;
  float* gx; float* gs; float* gm;
  ShapeB s;
  TSb xs, ss, ms, ns, gls, gos, grs;
  float scale_bound, lik_bound;
};

__global__ void __launch_bounds__(256) gc_bwd_kernel(const GcBwdP p) {
  const int n = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.s.E; i += gridDim.x * blockDim.x) {
    int c, h, w;
    decode_b(p.s, i, c, h, w);
    const float x = __ldg(p.x + offb(p.xs, n, c, h, w));
    const float sg = __ldg(p.scales + offb(p.ss, n, c, h, w));
    const bool hm = p.means != nullptr;
    const float mu = hm ? __ldg(p.means + offb(p.ms, n, c, h, w)) : 0.f;
    const bool train = p.noise != nullptr;
    const float outv = train ? x + __ldg(p.noise + offb(p.ns, n, c, h, w))
                             : (hm ? rintf(x - mu) + mu : rintf(x));
    float gd, gsc;
    gc_grads(outv, mu, hm, sg, p.scale_bound, p.lik_bound,
             p.g_lik ? __ldg(p.g_lik + offb(p.gls, n, c, h, w)) : 0.f, p.g_lik != nullptr,
             p.g_logsum, n, gd, gsc);
    const long long o = offb(p.grs, n, c, h, w);
    float gxv = 0.f, gmv = 0.f;
    if (train) {
      gxv = gd;           // values = (x + noise) - mu
      gmv = -gd;
      if (p.g_out) gxv += __ldg(p.g_out + offb(p.gos, n, c, h, w));   // outputs = x + noise
    }
    if (p.gx) p.gx[o] = gxv;
    if (p.gs) p.gs[o] = gsc;
    if (p.gm) p.gm[o] = gmv;
  }
}

// ---------------------------------------------------------------------------
// dual prior stage A backward: params = cat(y_hat_00, y_hat_11, means, scales)
// ---------------------------------------------------------------------------
struct StageABwdP {
  const float* g_params;
  float* gy; float* gm; float* gs;
  ShapeB s;
  TSb ps, os;
};
__global__ void __launch_bounds__(256) stage_a_bwd_kernel(const StageABwdP p) {
  const int n = blockIdx.y;
  const int half = p.s.C >> 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.s.E; i += gridDim.x * blockDim.x) {
    int c, h, w;
    decode_b(p.s, i, c, h, w);
    const bool sel = (((h + w) & 1) != 0) == (c >= half);
    const long long o = offb(p.os, n, c, h, w);
    if (p.gy) p.gy[o] = sel ? __ldg(p.g_params + offb(p.ps, n, c, h, w)) : 0.f;
    if (p.gm) p.gm[o] = __ldg(p.g_params + offb(p.ps, n, p.s.C + c, h, w));
    if (p.gs) p.gs[o] = __ldg(p.g_params + offb(p.ps, n, 2 * p.s.C + c, h, w));
  }
}

// ---------------------------------------------------------------------------
// dual prior stage B + Gaussian conditional backward
// ---------------------------------------------------------------------------
struct StageBBwdP {
  const float* g_yhat; const float* g_mh; const float* g_sh; const float* g_lik;
  const double* g_logsum;
  const float* y; const float* means; const float* scales; const float* prior; const float* noise;
  float* gy; float* gm; float* gs; float* gp;
  ShapeB s;
  TSb ys, ms, ss, prs, ns, gis, gos, gps;   // gis: incoming grads, gos: gy/gm/gs, gps: gp
  float scale_bound, lik_bound;
};
__global__ void __launch_bounds__(256) stage_b_bwd_kernel(const StageBBwdP p) {
  const int n = blockIdx.y;
  const int half = p.s.C >> 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.s.E; i += gridDim.x * blockDim.x) {
    int c, h, w;
    decode_b(p.s, i, c, h, w);
    const bool second = c >= half;
    const bool from_prior = (((h + w) & 1) != 0) == second;
    const int cm = second ? (p.s.C + (c - half)) : c;
    const float y = __ldg(p.y + offb(p.ys, n, c, h, w));
    float mu, sg;
    if (from_prior) {
      mu = __ldg(p.means + offb(p.ms, n, c, h, w));
      sg = __ldg(p.scales + offb(p.ss, n, c, h, w));
    } else {
      mu = __ldg(p.prior + offb(p.prs, n, cm, h, w));
      sg = __ldg(p.prior + offb(p.prs, n, cm + half, h, w));
    }
    const bool train = p.noise != nullptr;
    const float outv = train ? y + __ldg(p.noise + offb(p.ns, n, c, h, w)) : rintf(y - mu) + mu;
    const long long gi = offb(p.gis, n, c, h, w);
    float gd, gsc;
    gc_grads(outv, mu, true, sg, p.scale_bound, p.lik_bound, p.g_lik ? __ldg(p.g_lik + gi) : 0.f,
             p.g_lik != nullptr, p.g_logsum, n, gd, gsc);
    float g_y = p.g_yhat ? __ldg(p.g_yhat + gi) : 0.f;   // STE: dy_hat/dy = 1, dy_hat/dmu = 0
    float g_mu = p.g_mh ? __ldg(p.g_mh + gi) : 0.f;
    float g_sg = gsc + (p.g_sh ? __ldg(p.g_sh + gi) : 0.f);
    if (train) { g_y += gd; g_mu -= gd; }
    const long long o = offb(p.gos, n, c, h, w);
    if (p.gy) p.gy[o] = g_y;
    if (p.gm) p.gm[o] = from_prior ? g_mu : 0.f;
    if (p.gs) p.gs[o] = from_prior ? g_sg : 0.f;
    if (p.gp) {
      p.gp[offb(p.gps, n, cm, h, w)] = from_prior ? 0.f : g_mu;
      p.gp[offb(p.gps, n, cm + half, h, w)] = from_prior ? 0.f : g_sg;
    }
  }
}

// ---------------------------------------------------------------------------
// entropy bottleneck backward, filters = (3,3,3,3).  One CTA per channel: the
// 58 per-channel parameter gradients are reduced in shared memory and written
// without global atomics.
// ---------------------------------------------------------------------------
struct EbBwdP {
  const float* g_out; const float* g_zhat; const float* g_lik; const double* g_logsum;
  const float* z; const float* noise;
  const float* matrices; const float* biases; const float* factors; const float* medians;
  float* gz; float* g_mat; float* g_bias; float* g_fact; float* g_med;
  int N, C, H, W, HW;
  TSb zs, ns, gis, gos;
  float lik_bound;
};

// forward of one chain keeping the pre-activations; returns logits
__device__ __forceinline__ float eb_chain_fwd(float t, const float* sp, const float* b,
                                              const float* tf, float a[4][3], float l[4][3]) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    a[0][j] = sp[j] * t + b[j];
    l[0][j] = a[0][j] + tf[j] * tanhf(a[0][j]);
  }
#pragma unroll
  for (int k = 1; k <= 3; ++k) {
    const float* m = sp + 3 + 9 * (k - 1);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      a[k][j] = m[3 * j] * l[k - 1][0] + m[3 * j + 1] * l[k - 1][1] + m[3 * j + 2] * l[k - 1][2] +
                b[3 * k + j];
      l[k][j] = a[k][j] + tf[3 * k + j] * tanhf(a[k][j]);
    }
  }
  const float* m = sp + 30;
  return m[0] * l[3][0] + m[1] * l[3][1] + m[2] * l[3][2] + b[12];
}

// backward of one chain: accumulates d/d(sp, b, tf) into acc[58], returns d/dt
__device__ __forceinline__ float eb_chain_bwd(float g, float t, const float* sp, const float* tf,
                                              const float a[4][3], const float l[4][3],
                                              float* acc) {
  float gl[3];
  acc[33 + 12] += g;                    // bias of layer 4
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    acc[30 + i] += g * l[3][i];
    gl[i] = g * sp[30 + i];
  }
#pragma unroll
  for (int k = 3; k >= 1; --k) {
    const float* m = sp + 3 + 9 * (k - 1);
    float ga[3], gprev[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float th = tanhf(a[k][j]);
      acc[46 + 3 * k + j] += gl[j] * th;                       // d/d tanh(factor)
      ga[j] = gl[j] * (1.f + tf[3 * k + j] * (1.f - th * th));
      acc[33 + 3 * k + j] += ga[j];                            // bias
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        acc[3 + 9 * (k - 1) + 3 * j + i] += ga[j] * l[k - 1][i];
        gprev[i] += ga[j] * m[3 * j + i];
      }
    }
    gl[0] = gprev[0]; gl[1] = gprev[1]; gl[2] = gprev[2];
  }
  float gt = 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float th = tanhf(a[0][j]);
    acc[46 + j] += gl[j] * th;
    const float ga = gl[j] * (1.f + tf[j] * (1.f - th * th));
    acc[33 + j] += ga;
    acc[j] += ga * t;
    gt += ga * sp[j];
  }
  return gt;
}

__global__ void __launch_bounds__(128) eb_bwd_kernel(const EbBwdP p) {
  __shared__ float sp[33], bb[13], tf[12], red[59];
  const int c = blockIdx.x;
  if (threadIdx.x < 33) {
    const float a = __ldg(p.matrices + c * 33 + threadIdx.x);
    sp[threadIdx.x] = (a > 20.f) ? a : log1pf(expf(a));
  } else if (threadIdx.x < 46) {
    bb[threadIdx.x - 33] = __ldg(p.biases + c * 13 + threadIdx.x - 33);
  } else if (threadIdx.x < 58) {
    tf[threadIdx.x - 46] = tanhf(__ldg(p.factors + c * 12 + threadIdx.x - 46));
  }
  if (threadIdx.x < 59) red[threadIdx.x] = 0.f;
  __syncthreads();
  const float med = __ldg(p.medians + c);
  const bool train = p.noise != nullptr;
  float acc[59];
#pragma unroll
  for (int i = 0; i < 59; ++i) acc[i] = 0.f;
  const int total = p.N * p.HW;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int n = i / p.HW, r = i - n * p.HW;
    const int h = r / p.W, w = r - h * p.W;
    const float z = __ldg(p.z + offb(p.zs, n, c, h, w));
    const float outv = train ? z + __ldg(p.noise + offb(p.ns, n, c, h, w)) : rintf(z - med) + med;
    float al[4][3], ll[4][3], au[4][3], lu[4][3];
    const float tl = outv - 0.5f, tu = outv + 0.5f;
    const float lower = eb_chain_fwd(tl, sp, bb, tf, al, ll);
    const float upper = eb_chain_fwd(tu, sp, bb, tf, au, lu);
    const float s = lower + upper;
    const float sgn = (float)((s < 0.f) - (0.f < s));
    const float su = 1.f / (1.f + expf(-sgn * upper)), sl = 1.f / (1.f + expf(-sgn * lower));
    const float diff = su - sl;
    const float praw = fabsf(diff);
    const float pr = (praw < p.lik_bound) ? p.lik_bound : praw;
    const long long gi = offb(p.gis, n, c, h, w);
    float gp = p.g_lik ? __ldg(p.g_lik + gi) : 0.f;
    if (p.g_logsum) gp += (float)(p.g_logsum[n]) / pr;
    if (!(praw >= p.lik_bound || gp < 0.f)) gp = 0.f;
    const float sd = (diff > 0.f) ? 1.f : ((diff < 0.f) ? -1.f : 0.f);
    const float g_up = gp * sd * su * (1.f - su) * sgn;
    const float g_lo = -gp * sd * sl * (1.f - sl) * sgn;
    float g_outv = eb_chain_bwd(g_up, tu, sp, tf, au, lu, acc) +
                   eb_chain_bwd(g_lo, tl, sp, tf, al, ll, acc);
    if (p.g_out) g_outv += __ldg(p.g_out + gi);
    float gz = p.g_zhat ? __ldg(p.g_zhat + gi) : 0.f;   // z_hat = STE(z - med) + med
    if (train) gz += g_outv;          // outputs = z + noise
    else acc[58] += g_outv;           // outputs = round(z - med) + med
    if (p.gz) p.gz[offb(p.gos, n, c, h, w)] = gz;
  }
#pragma unroll
  for (int i = 0; i < 59; ++i) {
    const float v = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[i], v);
  }
  __syncthreads();
  if (threadIdx.x < 33) {
    // softplus'(a) = sigmoid(a) (1 above the threshold)
    const float a = __ldg(p.matrices + c * 33 + threadIdx.x);
    const float d = (a > 20.f) ? 1.f : 1.f / (1.f + expf(-a));
    if (p.g_mat) p.g_mat[c * 33 + threadIdx.x] = red[threadIdx.x] * d;
  } else if (threadIdx.x < 46) {
    if (p.g_bias) p.g_bias[c * 13 + threadIdx.x - 33] = red[threadIdx.x];
  } else if (threadIdx.x < 58) {
    const float t = tf[threadIdx.x - 46];
    if (p.g_fact) p.g_fact[c * 12 + threadIdx.x - 46] = red[threadIdx.x] * (1.f - t * t);
  } else if (threadIdx.x == 58) {
    if (p.g_med) p.g_med[c] = red[58];
  }
}

}  // namespace dvc

using namespace dvc;

extern "C" {

int dvc_gc_likelihood_bwd(const float* grad_lik, const double* grad_logsum, const float* grad_out,
                          const float* inputs, const float* scales, const float* means,
                          const float* noise, float* grad_inputs, float* grad_scales,
                          float* grad_means, int64_t N, int64_t C, int64_t H, int64_t W,
                          const int64_t in_st[4], const int64_t scales_st[4],
                          const int64_t means_st[4], const int64_t noise_st[4],
                          const int64_t glik_st[4], const int64_t gout_st[4],
                          const int64_t grad_st[4], float scale_bound, float likelihood_bound,
                          dvc_stream_t stream) {
  DVC_REQUIRE(inputs && scales && in_st && scales_st, "gc_likelihood_bwd: null input");
  DVC_REQUIRE(grad_lik || grad_logsum || grad_out, "gc_likelihood_bwd: no incoming gradient");
  DVC_REQUIRE(!grad_lik || glik_st, "gc_likelihood_bwd: grad_lik without strides");
  DVC_REQUIRE(!grad_out || gout_st, "gc_likelihood_bwd: grad_out without strides");
  DVC_REQUIRE(!means || means_st, "gc_likelihood_bwd: means without strides");
  DVC_REQUIRE(!noise || noise_st, "gc_likelihood_bwd: noise without strides");
  DVC_REQUIRE((grad_inputs || grad_scales || grad_means) && grad_st,
              "gc_likelihood_bwd: no output gradient requested");
  GcBwdP p;
  int rc = make_shape_b(p.s, N, C, H, W, in_st, "gc_likelihood_bwd");
  if (rc) return rc;
  p.g_lik = grad_lik; p.g_logsum = grad_logsum; p.g_out = grad_out;
  p.x = inputs; p.scales = scales; p.means = means; p.noise = noise;
  p.gx = grad_inputs; p.gs = grad_scales; p.gm = grad_means;
  p.xs = tsb(in_st); p.ss = tsb(scales_st); p.ms = tsb(means_st); p.ns = tsb(noise_st);
  p.gls = tsb(glik_st); p.gos = tsb(gout_st); p.grs = tsb(grad_st);
  p.scale_bound = scale_bound; p.lik_bound = likelihood_bound;
  dim3 grid(blocks_for(p.s.E, 2), (unsigned)N);
  gc_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("gc_bwd_kernel");
}

int dvc_dual_prior_stage_a_bwd(const float* grad_params, float* grad_y, float* grad_means,
                               float* grad_scales, int64_t N, int64_t C, int64_t H, int64_t W,
                               const int64_t gparams_st[4], const int64_t grad_st[4],
                               dvc_stream_t stream) {
  DVC_REQUIRE(grad_params && gparams_st && grad_st, "dual_prior_stage_a_bwd: null pointer");
  DVC_REQUIRE(grad_y || grad_means || grad_scales, "dual_prior_stage_a_bwd: nothing to compute");
  DVC_REQUIRE((C % 2) == 0, "dual_prior_stage_a_bwd: C must be even");
  StageABwdP p;
  int rc = make_shape_b(p.s, N, C, H, W, grad_st, "dual_prior_stage_a_bwd");
  if (rc) return rc;
  p.g_params = grad_params; p.gy = grad_y; p.gm = grad_means; p.gs = grad_scales;
  p.ps = tsb(gparams_st); p.os = tsb(grad_st);
  dim3 grid(blocks_for(p.s.E, 2), (unsigned)N);
  stage_a_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("stage_a_bwd_kernel");
}

int dvc_dual_prior_stage_b_gc_bwd(
    const float* grad_y_hat, const float* grad_means_hat, const float* grad_scales_hat,
    const float* grad_lik, const double* grad_logsum, const float* y, const float* means,
    const float* scales, const float* prior, const float* noise, float* grad_y, float* grad_means,
    float* grad_scales, float* grad_prior, int64_t N, int64_t C, int64_t H, int64_t W,
    const int64_t y_st[4], const int64_t means_st[4], const int64_t scales_st[4],
    const int64_t prior_st[4], const int64_t noise_st[4], const int64_t gin_st[4],
    const int64_t grad_st[4], const int64_t gprior_st[4], float scale_bound,
    float likelihood_bound, dvc_stream_t stream) {
  DVC_REQUIRE(y && means && scales && prior && y_st && means_st && scales_st && prior_st,
              "dual_prior_stage_b_gc_bwd: null input");
  DVC_REQUIRE(grad_y_hat || grad_means_hat || grad_scales_hat || grad_lik || grad_logsum,
              "dual_prior_stage_b_gc_bwd: no incoming gradient");
  DVC_REQUIRE(!(grad_y_hat || grad_means_hat || grad_scales_hat || grad_lik) || gin_st,
              "dual_prior_stage_b_gc_bwd: incoming gradients without strides");
  DVC_REQUIRE(!noise || noise_st, "dual_prior_stage_b_gc_bwd: noise without strides");
  DVC_REQUIRE(!(grad_y || grad_means || grad_scales) || grad_st,
              "dual_prior_stage_b_gc_bwd: outputs without strides");
  DVC_REQUIRE(!grad_prior || gprior_st, "dual_prior_stage_b_gc_bwd: grad_prior without strides");
  DVC_REQUIRE((C % 2) == 0, "dual_prior_stage_b_gc_bwd: C must be even");
  StageBBwdP p;
  int rc = make_shape_b(p.s, N, C, H, W, y_st, "dual_prior_stage_b_gc_bwd");
  if (rc) return rc;
  p.g_yhat = grad_y_hat; p.g_mh = grad_means_hat; p.g_sh = grad_scales_hat; p.g_lik = grad_lik;
  p.g_logsum = grad_logsum;
  p.y = y; p.means = means; p.scales = scales; p.prior = prior; p.noise = noise;
  p.gy = grad_y; p.gm = grad_means; p.gs = grad_scales; p.gp = grad_prior;
  p.ys = tsb(y_st); p.ms = tsb(means_st); p.ss = tsb(scales_st); p.prs = tsb(prior_st);
  p.ns = tsb(noise_st); p.gis = tsb(gin_st); p.gos = tsb(grad_st); p.gps = tsb(gprior_st);
  p.scale_bound = scale_bound; p.lik_bound = likelihood_bound;
  dim3 grid(blocks_for(p.s.E, 2), (unsigned)N);
  stage_b_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("stage_b_bwd_kernel");
}

int dvc_eb_likelihood_bwd(const float* grad_outputs, const float* grad_z_hat,
                          const float* grad_lik, const double* grad_logsum, const float* z,
                          const float* noise, const float* matrices, const float* biases,
                          const float* factors, const float* medians, float* grad_z,
                          float* grad_matrices, float* grad_biases, float* grad_factors,
                          float* grad_medians, int64_t N, int64_t C, int64_t H, int64_t W,
                          const int64_t z_st[4], const int64_t noise_st[4],
                          const int64_t gin_st[4], const int64_t gz_st[4],
                          float likelihood_bound, dvc_stream_t stream) {
  DVC_REQUIRE(z && matrices && biases && factors && medians && z_st, "eb_likelihood_bwd: null input");
  DVC_REQUIRE(grad_outputs || grad_z_hat || grad_lik || grad_logsum,
              "eb_likelihood_bwd: no incoming gradient");
  DVC_REQUIRE(!(grad_outputs || grad_z_hat || grad_lik) || gin_st,
              "eb_likelihood_bwd: incoming gradients without strides");
  DVC_REQUIRE(!noise || noise_st, "eb_likelihood_bwd: noise without strides");
  DVC_REQUIRE(!grad_z || gz_st, "eb_likelihood_bwd: grad_z without strides");
  DVC_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, "eb_likelihood_bwd: empty tensor");
  DVC_REQUIRE((long long)N * H * W < 2147483647LL && C <= 65535, "eb_likelihood_bwd: too large");
  EbBwdP p;
  p.g_out = grad_outputs; p.g_zhat = grad_z_hat; p.g_lik = grad_lik; p.g_logsum = grad_logsum;
  p.z = z; p.noise = noise; p.matrices = matrices; p.biases = biases; p.factors = factors;
  p.medians = medians;
  p.gz = grad_z; p.g_mat = grad_matrices; p.g_bias = grad_biases; p.g_fact = grad_factors;
  p.g_med = grad_medians;
  p.N = (int)N; p.C = (int)C; p.H = (int)H; p.W = (int)W; p.HW = (int)(H * W);
  p.zs = tsb(z_st); p.ns = tsb(noise_st); p.gis = tsb(gin_st); p.gos = tsb(gz_st);
  p.lik_bound = likelihood_bound;
  eb_bwd_kernel<<<(unsigned)C, 128, 0, (cudaStream_t)stream>>>(p);
  return check_launch("eb_bwd_kernel");
}

}  // extern "C"
