// dvc_entropy.cu -- quantisation, checkerboard dual prior, Gaussian-conditional
// and factorised entropy-bottleneck likelihoods, rate reduction (sm_100a).
//
// Replaces (arithmetic replayed op for op in fp32):
//   quantize_ste                      /root/reference/dmc/models/utils.py:149-152
//   get_mask / process_with_mask /
//   forward_dual_prior                /root/reference/dmc/models/video_model.py:152-216 (== :324-388)
//   z quantisation                    /root/reference/dmc/models/video_model.py:222-224, :394-396
//   GaussianConditional.forward       CompressAI (call sites video_model.py:232, :405)
//   EntropyBottleneck.forward         CompressAI (call sites video_model.py:220, :392)
//   collect_likelihoods_list          /root/reference/dmc/train.py:74-93
//
// PyTorch eager runs ~150 tiny kernels per context model for this (one per
// tensor op, each a full HBM round trip, plus an H2D copy of the checkerboard
// mask per forward).  Here each stage is ONE launch, every operand is read
// once, the checkerboard is computed from (h+w)&1, and the per-sample
// sum(ln p) is reduced with warp shuffles in the same kernel that produces p.
//
// Rounding discipline: eager rounds after every op, so products/sums that are
// separate torch ops use the *_rn helpers (never contracted); libdevice erfcf /
// tanhf / expf / log1pf / logf and IEEE division are the same routines ATen's
// CUDA kernels call.
#include <stdlib.h>

#include "dvc_common.cuh"

namespace dvc {

constexpr int kEThreads = 128;  // 128 threads x <= 32 registers = 4096 registers: fits beside 6 resident warp CTAs

// Per-sample iteration space, enumerated in the memory order of the lead tensor
// so that consecutive threads touch consecutive addresses:
//   mode 0 (NCHW)          a = c (grid.y), b = h*W + w
//   mode 1 (channels_last) a = h (grid.y), b = w*C + c
//   mode 2 / 3             a = 0, b = flat NCHW / NHWC index (any shape; real divisions)
// The inner split b -> (h,w) or (w,c) is one multiply-high by a precomputed
// reciprocal (exact while b*d < 2^32) instead of the ~40-instruction integer
// division the first version spent per element (ncu: the element-wise kernels
// were issue bound at 57% on index arithmetic).
struct It {
  int N, C, H, W, E;
  int mode, A, B;
  unsigned magic;   // ceil(2^32 / d), d = W (mode 0) or C (mode 1); 0 => divide
  int chunks;       // CTAs along b (grid.x)
  int per_block;    // ceil(B / chunks)
};
constexpr int kBatch = 2;  // elements per thread per trip: all loads first, then the math

__device__ __forceinline__ void decode(const It& s, int a, int b, int& c, int& h, int& w) {
  if (s.mode == 0) {
    c = a;
    h = s.magic ? (int)__umulhi((unsigned)b, s.magic) : b / s.W;
    w = b - h * s.W;
  } else if (s.mode == 1) {
    h = a;
    w = s.magic ? (int)__umulhi((unsigned)b, s.magic) : b / s.C;
    c = b - w * s.C;
  } else if (s.mode == 2) {
    w = b % s.W;
    int r = b / s.W;
    h = r % s.H;
    c = r / s.H;
  } else {
    c = b % s.C;
    int r = b / s.C;
    w = r % s.W;
    h = r / s.W;
  }
}

struct TS {  // element strides; (c,h,w) part of an offset always fits int32 (validated)
  long long n;
  int c, h, w;
};
static inline TS ts(const int64_t s[4]) {
  TS r;
  if (s) { r.n = s[0]; r.c = (int)s[1]; r.h = (int)s[2]; r.w = (int)s[3]; }
  else { r.n = 0; r.c = r.h = r.w = 0; }
  return r;
}
static inline bool ts_fits(const int64_t s[4], int64_t C, int64_t H, int64_t W) {
  if (!s) return true;
  long double m = 0;
  const int64_t e[3] = {C - 1, H - 1, W - 1};
  for (int i = 0; i < 3; ++i) m += (long double)(s[i + 1] < 0 ? -s[i + 1] : s[i + 1]) * e[i];
  return m < 2147483647.0L;
}
__device__ __forceinline__ long long off(const TS& s, int n, int c, int h, int w) {
  return n * s.n + (long long)(c * s.c + h * s.h + w * s.w);
}

// ---------------------------------------------------------------------------
// deterministic per-sample reduction of sum(ln p)
//   per-thread fp32 partial -> fp64 warp shuffle -> fp64 block partial in the
//   workspace -> the last block of the sample (ticket) adds the partials in a
//   fixed order.  Same grid => bit-identical result run to run.
// ---------------------------------------------------------------------------
struct RateWS {
  double* partial;       // [N][DVC_RATE_MAX_BLOCKS]
  unsigned int* ticket;  // [N]
  double* logsum;        // [N] (output; may be null => no reduction)
};

__device__ __forceinline__ double block_sum_fixed(double v, double* smem) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < (blockDim.x >> 5)) ? smem[threadIdx.x] : 0.0;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;  // valid in warp 0
}

// blocks_in_sample = number of blocks that contribute to sample n,
// block_in_sample  = this block's index among them.
__device__ __forceinline__ void rate_commit(const RateWS& ws, int n, int block_in_sample,
                                            int blocks_in_sample, float thread_sum) {
  if (ws.logsum == nullptr) return;
  __shared__ double red[32];
  __shared__ int is_last;
  const double bsum = block_sum_fixed((double)thread_sum, red);
  if (threadIdx.x == 0) {
    ws.partial[(long long)n * DVC_RATE_MAX_BLOCKS + block_in_sample] = bsum;
    __threadfence();
    const unsigned int t = atomicAdd(&ws.ticket[n], 1u);
    is_last = (t == (unsigned)blocks_in_sample - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    const volatile double* part = ws.partial + (long long)n * DVC_RATE_MAX_BLOCKS;
    double acc = 0.0;
    for (int i = threadIdx.x; i < blocks_in_sample; i += blockDim.x) acc += part[i];
    const double total = block_sum_fixed(acc, red);
    if (threadIdx.x == 0) {
      ws.logsum[n] = total;
      ws.ticket[n] = 0u;  // ready for the next launch
    }
  }
}

static int rate_ws(RateWS& ws, void* workspace, double* logsum, int64_t N) {
  ws.logsum = logsum;
  ws.partial = nullptr;
  ws.ticket = nullptr;
  if (logsum) {
    if (!workspace)
      return fail(DVC_ERR_WORKSPACE, "logsum requested but workspace is NULL "
                                     "(need dvc_rate_workspace_bytes(N) bytes, ticket words zeroed)");
    ws.partial = reinterpret_cast<double*>(workspace);
    ws.ticket = reinterpret_cast<unsigned int*>(ws.partial + (long long)N * DVC_RATE_MAX_BLOCKS);
  }
  return DVC_OK;
}

static unsigned magic_for(long long extent_b, int d) {
  // floor(b / d) == umulhi(b, ceil(2^32 / d)) for all b < extent_b iff extent_b * d < 2^32
  if (d <= 1 || extent_b * (long long)d >= (1LL << 32)) return 0u;
  return (unsigned)(((1ULL << 32) + (unsigned long long)d - 1ULL) / (unsigned long long)d);
}

static int make_it(It& s, int64_t N, int64_t C, int64_t H, int64_t W, const int64_t lead_st[4],
                   int per_thread, const char* who) {
  if (!(N > 0 && C > 0 && H > 0 && W > 0))
    return fail(DVC_ERR_INVALID_ARGUMENT, "%s: empty tensor", who);
  if (N > 65535) return fail(DVC_ERR_INVALID_ARGUMENT, "%s: N > 65535", who);
  const long long E = (long long)C * H * W;
  if (E >= 2147483647LL) return fail(DVC_ERR_INVALID_ARGUMENT, "%s: C*H*W too large", who);
  s.N = (int)N; s.C = (int)C; s.H = (int)H; s.W = (int)W;
  s.E = (int)E;
  const bool c_fast = lead_st && lead_st[1] == 1 && C > 1 && lead_st[3] != 1;
  long long A, B;
  int d;
  if (!c_fast) { A = C; B = (long long)H * W; d = (int)W; s.mode = 0; }
  else         { A = H; B = (long long)W * C; d = (int)C; s.mode = 1; }
  if (A > DVC_RATE_MAX_BLOCKS || A > 65535) {   // odd shapes: flat enumeration
    s.mode = c_fast ? 3 : 2;
    A = 1; B = E; d = 1;
  }
  s.A = (int)A; s.B = (int)B;
  s.magic = (s.mode <= 1) ? magic_for(B, d) : 0u;
  long long chunks = (B + (long long)kEThreads * per_thread - 1) / ((long long)kEThreads * per_thread);
  const long long cap = DVC_RATE_MAX_BLOCKS / A;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  s.chunks = (int)chunks;
  s.per_block = (int)((B + chunks - 1) / chunks);
  return DVC_OK;
}

static dim3 grid_of(const It& s) { return dim3((unsigned)s.chunks, (unsigned)s.A, (unsigned)s.N); }

#define DVC_REQUIRE_FITS(st, C_, who)                                                   \
  DVC_REQUIRE(ts_fits(st, C_, H, W), "%s: tensor extent exceeds 2^31 elements per sample", who)

// ---------------------------------------------------------------------------
// Gaussian conditional core (CompressAI GaussianConditional._likelihood +
// likelihood_lower_bound), SURVEY.md A.4
// ---------------------------------------------------------------------------
__device__ __forceinline__ float lower_bound(float x, float bound) {
  return (x < bound) ? bound : x;  // torch.max(x, bound): NaN propagates
}

__device__ __forceinline__ float gc_prob(float out, float mean, bool has_mean, float scale,
                                         float scale_bound, float lik_bound) {
  const float kNegRsqrt2 = -0.70710678118654752440f;  // float(-(2 ** -0.5))
  float v = has_mean ? sub_rn(out, mean) : out;
  v = fabsf(v);
  const float s = lower_bound(scale, scale_bound);
  const float u = div_rn(sub_rn(0.5f, v), s);
  const float l = div_rn(sub_rn(-0.5f, v), s);
  const float upper = mul_rn(0.5f, erfcf(mul_rn(kNegRsqrt2, u)));
  const float lower = mul_rn(0.5f, erfcf(mul_rn(kNegRsqrt2, l)));
  return lower_bound(sub_rn(upper, lower), lik_bound);
}

// EntropyModel.quantize(mode="dequantize"): round(x - m) + m, half to even
__device__ __forceinline__ float dequantize(float x, float m) {
  return add_rn(rintf(sub_rn(x, m)), m);
}

// ---------------------------------------------------------------------------
// quantize_ste forward (optionally around a per-channel offset)
// ---------------------------------------------------------------------------
struct QuantP {
  const float* x;
  const float* offset;
  float* q;
  It s;
  TS xs, qs;
  long long offset_st;
};

#define DVC_FOR_BATCH(s)                                                                   \
  const int n = blockIdx.z, a = blockIdx.y;                                                \
  const int b_begin = blockIdx.x * (s).per_block;                                          \
  const int b_end = min((s).B, b_begin + (s).per_block);                                   \
  for (int b0 = b_begin + threadIdx.x; b0 < b_end; b0 += kEThreads * kBatch)

__global__ void __launch_bounds__(kEThreads) quantize_kernel(const QuantP p) {
  DVC_FOR_BATCH(p.s) {
    float x[kBatch], m[kBatch];
    long long o[kBatch];
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      const int b = b0 + k * kEThreads;
      if (b < b_end) {
        int c, h, w;
        decode(p.s, a, b, c, h, w);
        x[k] = __ldg(p.x + off(p.xs, n, c, h, w));
        m[k] = p.offset ? __ldg(p.offset + c * p.offset_st) : 0.f;
        o[k] = off(p.qs, n, c, h, w);
      }
    }
#pragma unroll
    for (int k = 0; k < kBatch; ++k)
      if (b0 + k * kEThreads < b_end)
        p.q[o[k]] = p.offset ? add_rn(rintf(sub_rn(x[k], m[k])), m[k]) : rintf(x[k]);
  }
}

// ---------------------------------------------------------------------------
// dual prior, stage A: params = cat(y_hat_00, y_hat_11, means, scales)
// ---------------------------------------------------------------------------
struct StageAP {
  const float* y;
  const float* means;
  const float* scales;
  float* params;
  It s;
  TS ys, ms, ss, ps;
};

__global__ void __launch_bounds__(kEThreads, 16) stage_a_kernel(const StageAP p) {
  const int half = p.s.C >> 1;
  DVC_FOR_BATCH(p.s) {
    float y[kBatch], mu[kBatch], sg[kBatch];
    int po[kBatch];
    bool sel[kBatch];
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      const int b = b0 + k * kEThreads;
      if (b < b_end) {
        int c, h, w;
        decode(p.s, a, b, c, h, w);
        y[k] = __ldg(p.y + off(p.ys, n, c, h, w));
        mu[k] = __ldg(p.means + off(p.ms, n, c, h, w));
        sg[k] = __ldg(p.scales + off(p.ss, n, c, h, w));
        // mask_0 = [(h+w) even] for the first channel half, mask_1 for the second
        sel[k] = (((h + w) & 1) != 0) == (c >= half);
        po[k] = c * p.ps.c + h * p.ps.h + w * p.ps.w;
      }
    }
    float* __restrict__ pp = p.params + n * p.ps.n;
    const int cs = p.s.C * p.ps.c;
#pragma unroll
    for (int k = 0; k < kBatch; ++k)
      if (b0 + k * kEThreads < b_end) {
        pp[po[k]] = sel[k] ? dequantize(y[k], mu[k]) : 0.0f;
        pp[po[k] + cs] = mu[k];
        pp[po[k] + 2 * cs] = sg[k];
      }
  }
}

// ---------------------------------------------------------------------------
// dual prior, stage B + merge + Gaussian conditional + rate partial
// ---------------------------------------------------------------------------
struct StageBP {
  const float* y;
  const float* means;
  const float* scales;
  const float* prior;   // [N,2C,H,W] or null => module-level GC (means/scales final)
  const float* noise;
  float* y_hat;
  float* means_hat;
  float* scales_hat;
  float* lik;
  float* q_w0;
  float* q_w1;
  float* s_w0;
  float* s_w1;
  It s;
  TS ys, ms, ss, prs, ns, os, hs;
  float scale_bound, lik_bound;
  int has_mean;
  RateWS ws;
};

// kDense: every operand is a dense NCHW tensor (element (c,h,w) of a sample at c*H*W + h*W + w)
// and the iteration space is mode 0 (one channel plane per blockIdx.y).  Then a thread's index
// b IS the in-plane offset of every operand: the plane base pointers are per-block constants
// and the per-element integer work is one multiply-high (h, for the checkerboard parity)
// instead of seven 3-term stride products.  ncu on the strided version (r01): 230 warp
// instructions per element, a third of them IMAD/ISETP of the offset arithmetic, at a 32-register
// cap; this path is the one every contiguous latent takes (profiles/r02_gc_diet.md).
template <bool kDense>
__global__ void __launch_bounds__(kEThreads, 16) stage_b_gc_kernel(const StageBP p) {
  const int half = p.s.C >> 1;
  float lsum = 0.f;
  if (kDense) {
    const int n = blockIdx.z, c = blockIdx.y;
    const int HW = p.s.B;
    const bool second = c >= half;
    const long long plane = (long long)c * HW;
    const float* __restrict__ yb = p.y + n * p.ys.n + plane;
    const float* __restrict__ mb = p.has_mean ? p.means + n * p.ms.n + plane : nullptr;
    const float* __restrict__ sb = p.scales + n * p.ss.n + plane;
    const float* __restrict__ nb = p.noise ? p.noise + n * p.ns.n + plane : nullptr;
    // y_spatial_prior(params).chunk(4, 1) = (means_0, scales_0, means_1, scales_1)
    const float* __restrict__ pmb = nullptr;
    const float* __restrict__ psb = nullptr;
    if (p.prior) {
      const int cm = second ? (p.s.C + (c - half)) : c;
      pmb = p.prior + n * p.prs.n + (long long)cm * HW;
      psb = pmb + (long long)half * HW;
    }
    const long long ob = n * p.os.n + plane;
    const long long hb = n * p.hs.n + (long long)(second ? c - half : c) * HW;
    const int b_begin = blockIdx.x * p.s.per_block;
    const int b_end = min(HW, b_begin + p.s.per_block);
    for (int b0 = b_begin + threadIdx.x; b0 < b_end; b0 += kEThreads * kBatch) {
      float y[kBatch], mu[kBatch], sg[kBatch], nz[kBatch];
      bool fp[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const int b = b0 + k * kEThreads;
        fp[k] = true;
        if (b < b_end) {
          y[k] = __ldg(yb + b);
          const float* pm = mb + b;
          const float* ps = sb + b;
          if (p.prior) {
            const int h = p.s.magic ? (int)__umulhi((unsigned)b, p.s.magic) : b / p.s.W;
            const int w = b - h * p.s.W;
            fp[k] = (((h + w) & 1) != 0) == second;   // stage-A positions keep (means, scales)
            if (!fp[k]) { pm = pmb + b; ps = psb + b; }
          }
          mu[k] = p.has_mean ? __ldg(pm) : 0.f;
          sg[k] = __ldg(ps);
          nz[k] = nb ? __ldg(nb + b) : 0.f;
        }
      }
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const int b = b0 + k * kEThreads;
        if (b < b_end) {
          const float q = rintf(p.has_mean ? sub_rn(y[k], mu[k]) : y[k]);
          const float yh = p.has_mean ? add_rn(q, mu[k]) : q;
          const float outv = nb ? add_rn(y[k], nz[k]) : yh;
          const float pr = gc_prob(outv, mu[k], p.has_mean != 0, sg[k], p.scale_bound, p.lik_bound);
          if (p.y_hat) p.y_hat[ob + b] = (p.prior || !nb) ? yh : outv;
          if (p.means_hat) p.means_hat[ob + b] = mu[k];
          if (p.scales_hat) p.scales_hat[ob + b] = sg[k];
          if (p.lik) p.lik[ob + b] = pr;
          if (p.q_w0) {  // mode='compress' planes (video_model.py:209-214)
            if (fp[k]) { p.q_w0[hb + b] = q; p.s_w0[hb + b] = sg[k]; }
            else       { p.q_w1[hb + b] = q; p.s_w1[hb + b] = sg[k]; }
          }
          lsum += logf(pr);
        }
      }
    }
    rate_commit(p.ws, blockIdx.z, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y, lsum);
    return;
  }
  DVC_FOR_BATCH(p.s) {
    float y[kBatch], mu[kBatch], sg[kBatch], nz[kBatch];
    int oo[kBatch], oh[kBatch];
    bool fp[kBatch];
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      const int b = b0 + k * kEThreads;
      if (b < b_end) {
        int c, h, w;
        decode(p.s, a, b, c, h, w);
        y[k] = __ldg(p.y + off(p.ys, n, c, h, w));
        const float* pm = p.means + off(p.ms, n, c, h, w);
        const float* psg = p.scales + off(p.ss, n, c, h, w);
        fp[k] = true;
        if (p.prior) {
          const bool second = c >= half;
          fp[k] = (((h + w) & 1) != 0) == second;   // stage-A positions keep (means, scales)
          if (!fp[k]) {
            // y_spatial_prior(params).chunk(4, 1) = (means_0, scales_0, means_1, scales_1)
            const int cm = second ? (p.s.C + (c - half)) : c;
            pm = p.prior + off(p.prs, n, cm, h, w);
            psg = pm + (long long)half * p.prs.c;
          }
        }
        mu[k] = p.has_mean ? __ldg(pm) : 0.f;
        sg[k] = __ldg(psg);
        nz[k] = p.noise ? __ldg(p.noise + off(p.ns, n, c, h, w)) : 0.f;
        oo[k] = c * p.os.c + h * p.os.h + w * p.os.w;
        oh[k] = ((c >= half) ? c - half : c) * p.hs.c + h * p.hs.h + w * p.hs.w;
      }
    }
#pragma unroll
    for (int k = 0; k < kBatch; ++k)
      if (b0 + k * kEThreads < b_end) {
        // STE-rounded latent (always rounded, training or not: video_model.py:165-166)
        const float q = rintf(p.has_mean ? sub_rn(y[k], mu[k]) : y[k]);
        const float yh = p.has_mean ? add_rn(q, mu[k]) : q;
        // GaussianConditional.quantize: noise in training, dequantize in eval
        const float outv = p.noise ? add_rn(y[k], nz[k]) : yh;
        const float pr = gc_prob(outv, mu[k], p.has_mean != 0, sg[k], p.scale_bound, p.lik_bound);
        const long long o = n * p.os.n + oo[k];
        if (p.y_hat) p.y_hat[o] = (p.prior || !p.noise) ? yh : outv;
        if (p.means_hat) p.means_hat[o] = mu[k];
        if (p.scales_hat) p.scales_hat[o] = sg[k];
        if (p.lik) p.lik[o] = pr;
        if (p.q_w0) {  // mode='compress' planes (video_model.py:209-214)
          const long long o2 = n * p.hs.n + oh[k];
          if (fp[k]) { p.q_w0[o2] = q; p.s_w0[o2] = sg[k]; }
          else       { p.q_w1[o2] = q; p.s_w1[o2] = sg[k]; }
        }
        lsum += logf(pr);
      }
  }
  rate_commit(p.ws, blockIdx.z, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y, lsum);
}

// ---------------------------------------------------------------------------
// factorised entropy bottleneck, filters = (3,3,3,3)   
// ---------------------------------------------------------------------------
struct EBP {
  const float* z;
  const float* noise;
  const float* matrices;  // [C][33]
  const float* biases;    // [C][13]
  const float* factors;   // [C][12]
  const float* medians;   // [C]
  float* outputs;
  float* z_hat;
  float* lik;
  int N, C, H, W, HW;
  int chunks;             // blocks per (n, c) plane
  TS zs, ns, os;
  float lik_bound;
  RateWS ws;
};

__device__ __forceinline__ float softplus_t(float a) {
  // ATen softplus (beta=1, threshold=20): a > 20 ? a : log1p(exp(a))
  return (a > 20.0f) ? a : log1pf(expf(a));
}
__device__ __forceinline__ float sigmoid_t(float a) {
  return div_rn(1.0f, add_rn(1.0f, expf(-a)));  // ATen: 1 / (1 + exp(-a))
}

// logits_cumulative for one scalar; sp = softplus(matrices), tf = tanh(factors)
__device__ __forceinline__ float eb_logits(float t, const float* sp, const float* b,
                                           const float* tf) {
  float l[3];
  // layer 0: [3x1] @ [1] + bias, gated tanh
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float a = add_rn(mul_rn(sp[j], t), b[j]);
    l[j] = add_rn(a, mul_rn(tf[j], tanhf(a)));
  }
  // layers 1..3: [3x3]
#pragma unroll
  for (int k = 1; k <= 3; ++k) {
    const float* m = sp + 3 + 9 * (k - 1);
    float o[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float acc = mul_rn(m[3 * j + 0], l[0]);     // K<=3 dot product, FMA chain
      acc = fmaf(m[3 * j + 1], l[1], acc);
      acc = fmaf(m[3 * j + 2], l[2], acc);
      float a = add_rn(acc, b[3 * k + j]);
      o[j] = add_rn(a, mul_rn(tf[3 * k + j], tanhf(a)));
    }
    l[0] = o[0]; l[1] = o[1]; l[2] = o[2];
  }
  // layer 4: [1x3], no gate
  const float* m = sp + 30;
  float acc = mul_rn(m[0], l[0]);
  acc = fmaf(m[1], l[1], acc);
  acc = fmaf(m[2], l[2], acc);
  return add_rn(acc, b[12]);
}

__global__ void __launch_bounds__(128, 16) eb_kernel(const EBP p) {
  __shared__ float sp[33], bb[13], tf[12];
  const int c = blockIdx.y, n = blockIdx.z;
  if (threadIdx.x < 33) sp[threadIdx.x] = softplus_t(__ldg(p.matrices + c * 33 + threadIdx.x));
  else if (threadIdx.x < 46) bb[threadIdx.x - 33] = __ldg(p.biases + c * 13 + threadIdx.x - 33);
  else if (threadIdx.x < 58) tf[threadIdx.x - 46] = tanhf(__ldg(p.factors + c * 12 + threadIdx.x - 46));
  __syncthreads();
  const float med = __ldg(p.medians + c);
  float lsum = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.HW; i += gridDim.x * blockDim.x) {
    const int h = i / p.W, w = i - h * p.W;
    const float z = __ldg(p.z + off(p.zs, n, c, h, w));
    const float zq = dequantize(z, med);
    const float outv = p.noise ? add_rn(z, __ldg(p.noise + off(p.ns, n, c, h, w))) : zq;
    const float lower = eb_logits(sub_rn(outv, 0.5f), sp, bb, tf);
    const float upper = eb_logits(add_rn(outv, 0.5f), sp, bb, tf);
    const float s = add_rn(lower, upper);
    const float sgn = (float)((s < 0.f) - (0.f < s));  // -sign(lower + upper)
    float pr = fabsf(sub_rn(sigmoid_t(mul_rn(sgn, upper)), sigmoid_t(mul_rn(sgn, lower))));
    pr = lower_bound(pr, p.lik_bound);
    const long long o = off(p.os, n, c, h, w);
    if (p.outputs) p.outputs[o] = outv;
    if (p.z_hat) p.z_hat[o] = zq;
    if (p.lik) p.lik[o] = pr;
    lsum += logf(pr);
  }
  rate_commit(p.ws, n, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y, lsum);
}

// ---------------------------------------------------------------------------
// stand-alone log-sum and finalise
// ---------------------------------------------------------------------------
struct LogSumP {
  const float* lik;
  It s;
  TS ls;
  RateWS ws;
};
__global__ void __launch_bounds__(kEThreads) log_sum_kernel(const LogSumP p) {
  float lsum = 0.f;
  DVC_FOR_BATCH(p.s) {
    float v[kBatch];
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      const int b = b0 + k * kEThreads;
      v[k] = 1.0f;
      if (b < b_end) {
        int c, h, w;
        decode(p.s, a, b, c, h, w);
        v[k] = __ldg(p.lik + off(p.ls, n, c, h, w));
      }
    }
#pragma unroll
    for (int k = 0; k < kBatch; ++k)
      if (b0 + k * kEThreads < b_end) lsum += logf(v[k]);
  }
  rate_commit(p.ws, blockIdx.z, blockIdx.y * gridDim.x + blockIdx.x, gridDim.x * gridDim.y, lsum);
}

__global__ void rate_finalize_kernel(const double* __restrict__ logsums, int K, int N,
                                     double denom, float* __restrict__ bpp,
                                     float* __restrict__ bpp_total, double* __restrict__ bits) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
    float tot = 0.f;
    double lt = 0.0;
    for (int k = 0; k < K; ++k) {
      const double ls = logsums[(long long)k * N + n];
      const float b = (float)(ls / denom);   // train.py:83
      if (bpp) bpp[(long long)k * N + n] = b;
      tot = add_rn(tot, b);                  // train.py:85 bpp_loss += bpp
      lt += ls;
    }
    if (bpp_total) bpp_total[n] = tot;
    if (bits) bits[n] = -lt / 0.69314718055994530942;
  }
}

}  // namespace dvc

using namespace dvc;

extern "C" {

int64_t dvc_rate_workspace_bytes(int64_t N) {
  if (N < 1) N = 1;
  return N * (int64_t)DVC_RATE_MAX_BLOCKS * (int64_t)sizeof(double) +
         ((N * (int64_t)sizeof(unsigned int) + 15) / 16) * 16;
}

int dvc_quantize_fwd(const float* x, const float* offset, float* q, int64_t N, int64_t C,
                     int64_t H, int64_t W, const int64_t x_st[4], int64_t offset_st,
                     const int64_t q_st[4], dvc_stream_t stream) {
  DVC_REQUIRE(x && q && x_st && q_st, "quantize: null pointer");
  DVC_REQUIRE_FITS(x_st, C, "quantize");
  DVC_REQUIRE_FITS(q_st, C, "quantize");
  QuantP p;
  int rc = make_it(p.s, N, C, H, W, x_st, kBatch, "quantize");
  if (rc) return rc;
  p.x = x; p.offset = offset; p.q = q;
  p.xs = ts(x_st); p.qs = ts(q_st);
  p.offset_st = offset_st;
  quantize_kernel<<<grid_of(p.s), kEThreads, 0, (cudaStream_t)stream>>>(p);
  return check_launch("quantize_kernel");
}

int dvc_dual_prior_stage_a_fwd(const float* y, const float* means, const float* scales,
                               float* params, int64_t N, int64_t C, int64_t H, int64_t W,
                               const int64_t y_st[4], const int64_t means_st[4],
                               const int64_t scales_st[4], const int64_t params_st[4],
                               dvc_stream_t stream) {
  DVC_REQUIRE(y && means && scales && params, "dual_prior_stage_a: null pointer");
  DVC_REQUIRE(y_st && means_st && scales_st && params_st, "dual_prior_stage_a: null strides");
  DVC_REQUIRE((C % 2) == 0, "dual_prior_stage_a: C must be even (got %lld)", (long long)C);
  DVC_REQUIRE((H % 2) == 0 && (W % 2) == 0,
              "dual_prior_stage_a: checkerboard needs even H and W (got %lld x %lld)",
              (long long)H, (long long)W);
  DVC_REQUIRE_FITS(y_st, C, "dual_prior_stage_a");
  DVC_REQUIRE_FITS(means_st, C, "dual_prior_stage_a");
  DVC_REQUIRE_FITS(scales_st, C, "dual_prior_stage_a");
  DVC_REQUIRE_FITS(params_st, 3 * C, "dual_prior_stage_a");
  StageAP p;
  int rc = make_it(p.s, N, C, H, W, y_st, kBatch, "dual_prior_stage_a");
  if (rc) return rc;
  p.y = y; p.means = means; p.scales = scales; p.params = params;
  p.ys = ts(y_st); p.ms = ts(means_st); p.ss = ts(scales_st); p.ps = ts(params_st);
  stage_a_kernel<<<grid_of(p.s), kEThreads, 0, (cudaStream_t)stream>>>(p);
  return check_launch("stage_a_kernel");
}

static int launch_stage_b(StageBP& p, int64_t N, double* logsum, void* workspace,
                          cudaStream_t stream, const char* who) {
  int rc = rate_ws(p.ws, workspace, logsum, N);
  if (rc) return rc;
  // dense NCHW operands in plane-per-block enumeration -> the offset-free path
  auto dense = [&](const TS& t, const void* ptr, int C) {
    return !ptr || (t.w == 1 && t.h == p.s.W && t.c == p.s.H * p.s.W && C > 0);
  };
  const bool all_dense =
      p.s.mode == 0 && dense(p.ys, p.y, p.s.C) && dense(p.ms, p.has_mean ? p.means : nullptr, p.s.C) &&
      dense(p.ss, p.scales, p.s.C) && dense(p.prs, p.prior, 2 * p.s.C) && dense(p.ns, p.noise, p.s.C) &&
      dense(p.os, (p.y_hat || p.means_hat || p.scales_hat || p.lik) ? (const void*)p.y : nullptr, p.s.C) &&
      dense(p.hs, p.q_w0, p.s.C / 2);
  static int use_dense = -1;   // tuning knob (not API): DVC_GC_DENSE=0 forces the strided path (A/B)
  if (use_dense < 0) {
    const char* e = getenv("DVC_GC_DENSE");
    use_dense = e ? atoi(e) : 1;
  }
  if (all_dense && use_dense)
    stage_b_gc_kernel<true><<<grid_of(p.s), kEThreads, 0, stream>>>(p);
  else
    stage_b_gc_kernel<false><<<grid_of(p.s), kEThreads, 0, stream>>>(p);
  return check_launch(who);
}

int dvc_dual_prior_stage_b_gc_fwd(
    const float* y, const float* means, const float* scales, const float* prior,
    const float* noise, float* y_hat, float* means_hat, float* scales_hat, float* lik,
    float* q_w0, float* q_w1, float* s_w0, float* s_w1, double* logsum, void* workspace,
    int64_t N, int64_t C, int64_t H, int64_t W, const int64_t y_st[4],
    const int64_t means_st[4], const int64_t scales_st[4], const int64_t prior_st[4],
    const int64_t noise_st[4], const int64_t out_st[4], const int64_t half_st[4],
    float scale_bound, float likelihood_bound, dvc_stream_t stream) {
  DVC_REQUIRE(y && means && scales && prior, "dual_prior_stage_b_gc: null input");
  DVC_REQUIRE(y_st && means_st && scales_st && prior_st, "dual_prior_stage_b_gc: null strides");
  DVC_REQUIRE((C % 2) == 0, "dual_prior_stage_b_gc: C must be even (got %lld)", (long long)C);
  DVC_REQUIRE((H % 2) == 0 && (W % 2) == 0,
              "dual_prior_stage_b_gc: checkerboard needs even H and W (got %lld x %lld)",
              (long long)H, (long long)W);
  DVC_REQUIRE(!noise || noise_st, "dual_prior_stage_b_gc: noise without strides");
  DVC_REQUIRE(!(y_hat || means_hat || scales_hat || lik) || out_st,
              "dual_prior_stage_b_gc: outputs without strides");
  const bool any_half = q_w0 || q_w1 || s_w0 || s_w1;
  DVC_REQUIRE(!any_half || (q_w0 && q_w1 && s_w0 && s_w1 && half_st),
              "dual_prior_stage_b_gc: compress planes must be given together");
  DVC_REQUIRE_FITS(y_st, C, "dual_prior_stage_b_gc");
  DVC_REQUIRE_FITS(means_st, C, "dual_prior_stage_b_gc");
  DVC_REQUIRE_FITS(scales_st, C, "dual_prior_stage_b_gc");
  DVC_REQUIRE_FITS(prior_st, 2 * C, "dual_prior_stage_b_gc");
  DVC_REQUIRE_FITS(noise_st, C, "dual_prior_stage_b_gc");
  DVC_REQUIRE_FITS(out_st, C, "dual_prior_stage_b_gc");
  DVC_REQUIRE_FITS(half_st, C / 2, "dual_prior_stage_b_gc");
  StageBP p;
  int rc = make_it(p.s, N, C, H, W, y_st, kBatch, "dual_prior_stage_b_gc");
  if (rc) return rc;
  p.y = y; p.means = means; p.scales = scales; p.prior = prior; p.noise = noise;
  p.y_hat = y_hat; p.means_hat = means_hat; p.scales_hat = scales_hat; p.lik = lik;
  p.q_w0 = q_w0; p.q_w1 = q_w1; p.s_w0 = s_w0; p.s_w1 = s_w1;
  p.ys = ts(y_st); p.ms = ts(means_st); p.ss = ts(scales_st); p.prs = ts(prior_st);
  p.ns = ts(noise_st); p.os = ts(out_st); p.hs = ts(half_st);
  p.scale_bound = scale_bound; p.lik_bound = likelihood_bound;
  p.has_mean = 1;
  return launch_stage_b(p, N, logsum, workspace, (cudaStream_t)stream, "stage_b_gc_kernel");
}

int dvc_gc_likelihood_fwd(const float* inputs, const float* scales, const float* means,
                          const float* noise, float* outputs, float* lik, double* logsum,
                          void* workspace, int64_t N, int64_t C, int64_t H, int64_t W,
                          const int64_t in_st[4], const int64_t scales_st[4],
                          const int64_t means_st[4], const int64_t noise_st[4],
                          const int64_t out_st[4], float scale_bound, float likelihood_bound,
                          dvc_stream_t stream) {
  DVC_REQUIRE(inputs && scales && in_st && scales_st, "gc_likelihood: null input");
  DVC_REQUIRE(!means || means_st, "gc_likelihood: means without strides");
  DVC_REQUIRE(!noise || noise_st, "gc_likelihood: noise without strides");
  DVC_REQUIRE(!(outputs || lik) || out_st, "gc_likelihood: outputs without strides");
  DVC_REQUIRE_FITS(in_st, C, "gc_likelihood");
  DVC_REQUIRE_FITS(scales_st, C, "gc_likelihood");
  DVC_REQUIRE_FITS(means_st, C, "gc_likelihood");
  DVC_REQUIRE_FITS(noise_st, C, "gc_likelihood");
  DVC_REQUIRE_FITS(out_st, C, "gc_likelihood");
  StageBP p;
  int rc = make_it(p.s, N, C, H, W, in_st, kBatch, "gc_likelihood");
  if (rc) return rc;
  p.y = inputs; p.means = means; p.scales = scales; p.prior = nullptr; p.noise = noise;
  p.y_hat = outputs; p.means_hat = nullptr; p.scales_hat = nullptr; p.lik = lik;
  p.q_w0 = p.q_w1 = p.s_w0 = p.s_w1 = nullptr;
  p.ys = ts(in_st); p.ms = ts(means_st); p.ss = ts(scales_st); p.prs = ts(nullptr);
  p.ns = ts(noise_st); p.os = ts(out_st); p.hs = ts(nullptr);
  p.scale_bound = scale_bound; p.lik_bound = likelihood_bound;
  p.has_mean = means ? 1 : 0;
  return launch_stage_b(p, N, logsum, workspace, (cudaStream_t)stream, "gc_likelihood_kernel");
}

int dvc_eb_likelihood_fwd(const float* z, const float* noise, const float* matrices,
                          const float* biases, const float* factors, const float* medians,
                          float* outputs, float* z_hat, float* lik, double* logsum,
                          void* workspace, int64_t N, int64_t C, int64_t H, int64_t W,
                          const int64_t z_st[4], const int64_t noise_st[4],
                          const int64_t out_st[4], float likelihood_bound,
                          dvc_stream_t stream) {
  DVC_REQUIRE(z && matrices && biases && factors && medians && z_st, "eb_likelihood: null input");
  DVC_REQUIRE(!noise || noise_st, "eb_likelihood: noise without strides");
  DVC_REQUIRE(!(outputs || z_hat || lik) || out_st, "eb_likelihood: outputs without strides");
  DVC_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, "eb_likelihood: empty tensor");
  DVC_REQUIRE(N <= 65535 && C <= DVC_RATE_MAX_BLOCKS, "eb_likelihood: N <= 65535 and C <= %d",
              DVC_RATE_MAX_BLOCKS);
  DVC_REQUIRE((long long)H * W < 2147483647LL, "eb_likelihood: H*W too large");
  EBP p;
  p.z = z; p.noise = noise; p.matrices = matrices; p.biases = biases; p.factors = factors;
  p.medians = medians; p.outputs = outputs; p.z_hat = z_hat; p.lik = lik;
  p.N = (int)N; p.C = (int)C; p.H = (int)H; p.W = (int)W; p.HW = (int)(H * W);
  int chunks = (p.HW + 127) / 128;
  const int cap = DVC_RATE_MAX_BLOCKS / (int)C;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  p.chunks = chunks;
  p.zs = ts(z_st); p.ns = ts(noise_st); p.os = ts(out_st);
  p.lik_bound = likelihood_bound;
  int rc = rate_ws(p.ws, workspace, logsum, N);
  if (rc) return rc;
  dim3 grid((unsigned)chunks, (unsigned)C, (unsigned)N);
  eb_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(p);
  return check_launch("eb_kernel");
}

int dvc_log_sum_fwd(const float* lik, double* logsum, void* workspace, int64_t N, int64_t C,
                    int64_t H, int64_t W, const int64_t lik_st[4], dvc_stream_t stream) {
  DVC_REQUIRE(lik && logsum && lik_st, "log_sum: null pointer");
  DVC_REQUIRE_FITS(lik_st, C, "log_sum");
  LogSumP p;
  int rc = make_it(p.s, N, C, H, W, lik_st, 2 * kBatch, "log_sum");
  if (rc) return rc;
  p.lik = lik;
  p.ls = ts(lik_st);
  rc = rate_ws(p.ws, workspace, logsum, N);
  if (rc) return rc;
  log_sum_kernel<<<grid_of(p.s), kEThreads, 0, (cudaStream_t)stream>>>(p);
  return check_launch("log_sum_kernel");
}

int dvc_rate_finalize(const double* logsums, int K, int64_t N, double num_pixels, float* bpp,
                      float* bpp_total, double* bits, dvc_stream_t stream) {
  DVC_REQUIRE(logsums && K >= 1 && N >= 1, "rate_finalize: bad arguments");
  DVC_REQUIRE(num_pixels > 0, "rate_finalize: num_pixels must be positive");
  const double denom = -0.69314718055994530942 * num_pixels;   // -math.log(2) * num_pixels
  const int threads = 128;
  const unsigned blocks = (unsigned)((N + threads - 1) / threads);
  rate_finalize_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(logsums, K, (int)N, denom,
                                                                     bpp, bpp_total, bits);
  return check_launch("rate_finalize_kernel");
}

}  // extern "C"
