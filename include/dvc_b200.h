/*
 * dvc_b200.h -- C ABI of libdvc_b200.so: the B200 (sm_100a) hot path of the
 * DMC contextual P-frame codec (lumingzzz/DeepVideoCodec, dmc/models).
 *
 * The reference has no FFI: its seam is Python name binding.
 * Each entry point below therefore names the reference Python function whose
 * arithmetic it replaces; deepvideocodec_b200/*.py re-creates those Python
 * names on top of this ABI (ctypes), INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to fp32 unless stated otherwise;
 *  - tensors are logical [N,C,H,W]; `*_st` arguments are ELEMENT strides in
 *    (N,C,H,W) order, so NCHW-contiguous and channels_last both work.  Fast
 *    paths need: channel stride 1, C % 4 == 0, 16-byte aligned base and
 *    strides (channels_last); everything else takes the strided path;
 *  - the library never allocates, frees or synchronises; all launches go to
 *    `stream`; every call is CUDA-graph capturable and re-entrant;
 *  - return value: 0 on success, negative dvc_status on failure; the message
 *    of the last failure on the calling thread is dvc_last_error_string();
 *  - nullable arguments are marked [opt].
 */
#ifndef DVC_B200_H_
#define DVC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* dvc_stream_t; /* == cudaStream_t */

typedef enum dvc_status {
  DVC_OK = 0,
  DVC_ERR_INVALID_ARGUMENT = -1,
  DVC_ERR_UNSUPPORTED = -2,
  DVC_ERR_CUDA = -3,
  DVC_ERR_WORKSPACE = -4
} dvc_status;

/* flags for the warp entry points */
enum {
  /* Divide the flow by (S-1)/2 with an IEEE division, as PyTorch-CPU does.
   * Default (flag clear) multiplies by the fp32 reciprocal, as PyTorch-CUDA
   * eager does for `tensor / python_float` -- the two differ in the last bit
   * of the source coordinate. */
  DVC_WARP_IEEE_DIV = 1
};

int dvc_version(void);                     /* MAJOR*10000 + MINOR*100 + PATCH */
const char* dvc_last_error_string(void);   /* thread local, never NULL */
/* "src=<sha256[:16] of csrc/ *.cu *.cuh, this header and the nvcc flags> arch=sm_100a": lets a
 * reader check that a shipped (git-ignored) binary was built from the committed sources --
 * `python -c "import deepvideocodec_b200 as d; print(d.source_hash(), d.lib().dvc_build_info())"`. */
const char* dvc_build_info(void);
/* Number of SMs / compute capability of the current device (diagnostics). */
int dvc_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------
 * Piece 1: warp.  Replaces flow_warp/torch_warp, dmc/models/layers.py:175-198
 * (linspace base grid + flow/((S-1)/2) + F.grid_sample bilinear/border/
 * align_corners=True), replayed op for op in fp32.
 *   im   [N,C,H,W]   flow [N,2,H,W] (pixels; ch0 = x, ch1 = y)   out [N,C,H,W]
 * ------------------------------------------------------------------------- */
int dvc_flow_warp_fwd(const float* im, const float* flow, float* out,
                      int64_t N, int64_t C, int64_t H, int64_t W,
                      const int64_t im_st[4], const int64_t flow_st[4],
                      const int64_t out_st[4], int flags, dvc_stream_t stream);

/* Backward of the above (ATen grid_sampler_2d_backward semantics: zero
 * gradient through the border clip, scatter-add into grad_im).
 *   grad_im   [opt] [N,C,H,W], MUST be zero-filled by the caller (atomics)
 *   grad_flow [opt] [N,2,H,W] */
int dvc_flow_warp_bwd(const float* grad_out, const float* im, const float* flow,
                      float* grad_im, float* grad_flow,
                      int64_t N, int64_t C, int64_t H, int64_t W,
                      const int64_t gout_st[4], const int64_t im_st[4],
                      const int64_t flow_st[4], const int64_t gim_st[4],
                      const int64_t gflow_st[4], int flags, dvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * Flow pyramid.  Replaces bilineardownsacling(x) (layers.py:201-206) followed
 * by `* post_scale` (video_model.py:499-500 uses / 2 -> post_scale = 0.5).
 *   x [N,C,H,W] -> y [N,C,H/2,W/2];  any H,W >= 2 (odd sizes use the general
 *   align_corners=False formula; for odd sizes the backward scatters with
 *   atomics and grad_x MUST be zero-filled by the caller).
 * ------------------------------------------------------------------------- */
int dvc_bilinear_down2_fwd(const float* x, float* y, int64_t N, int64_t C,
                           int64_t H, int64_t W, const int64_t x_st[4],
                           const int64_t y_st[4], float post_scale,
                           dvc_stream_t stream);
int dvc_bilinear_down2_bwd(const float* grad_y, float* grad_x, int64_t N,
                           int64_t C, int64_t H, int64_t W,
                           const int64_t gy_st[4], const int64_t gx_st[4],
                           float post_scale, dvc_stream_t stream);
/* mv -> (mv/2 downscaled, mv/4 downscaled) in ONE launch; needs H%4==0,W%4==0
 * (always true after the reference pads to x64, dmc/test.py:75-88). */
int dvc_flow_pyramid_fwd(const float* mv, float* mv2, float* mv3, int64_t N,
                         int64_t H, int64_t W, const int64_t mv_st[4],
                         const int64_t mv2_st[4], const int64_t mv3_st[4],
                         dvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * Motion-compensation warps in one launch.  Replaces the non-conv part of
 * DMC.motion_compensation, dmc/models/video_model.py:497-504: warp x_ref by mv,
 * feat1 by mv, feat2 by mv2, feat3 by mv3.  mv2/mv3 either come from
 * dvc_flow_pyramid_fwd or are derived inside the kernel (flow_downscale).  All
 * four problems are tiled into one grid (largest first, so the small ones fill
 * the tail).
 * ------------------------------------------------------------------------- */
typedef struct dvc_warp_task {
  const float* im;
  const float* flow;
  float* out;
  int64_t N, C, H, W;            /* extents of im / out */
  int64_t im_st[4], flow_st[4], out_st[4];
  /* 0: flow is [N,2,H,W].  k = 1, 2: flow is the full-resolution motion field
   * [N,2,H<<k,W<<k]; the kernel reduces it on the fly with the reference's
   * pyramid arithmetic (bilineardownsacling(.)/2 applied k times), so mv2/mv3
   * never have to be written to HBM.  Bit-identical to warping with the
   * materialised pyramid level. */
  int64_t flow_downscale;
} dvc_warp_task;
int dvc_warp_multi_fwd(const dvc_warp_task* tasks /* HOST array */, int n_tasks,
                       int flags, dvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * warp fused into the 3x3 convolution that consumes it, as a
 * tcgen05 (TF32, fp32 accumulate) implicit GEMM.  Replaces, for one scale,
 *   context = flow_warp(ref_feature, mv)                    video_model.py:502-504
 *   conv    = convK_out(torch.cat((up, context), dim=1))    video_model.py:55-61
 * (nn.Conv2d(Ce + Cf, 64, 3, padding=1); `up` is `extra`, may be absent).
 *   feat  [N,Cf,H,W]  channels_last DENSE (strides H*W*Cf, 1, W*Cf, Cf)
 *   extra [N,Ce,H,W]  channels_last dense, [opt] (Ce = 0)
 *   flow  [N,2,H<<k,W<<k] any strides, k = flow_downscale (see dvc_warp_task)
 *   out_warp [N,Cf,H,W] channels_last dense, [opt]: bit-identical to
 *            dvc_flow_warp_fwd;  out_conv [N,64,H,W] channels_last dense.
 * Cf, Ce multiples of 16; Co = 64.  The weight [64, Ce+Cf, 3, 3] is re-laid
 * once by dvc_conv3x3_pack_weights into dvc_conv3x3_packed_weight_floats()
 * floats (16-channel slices in the kernel's consumption order -- `extra` and
 * `feat` slices alternate --, each as [K-step][tap][4-channel group][co][4], the
 * UMMA K-major operand layout; the packing therefore depends on the Ce/Cf split).
 * PyTorch's cuDNN convolutions run in TF32 by default; this kernel truncates
 * the operands to TF32 in the tensor core and accumulates in fp32.
 * ------------------------------------------------------------------------- */
int64_t dvc_conv3x3_packed_weight_floats(int64_t Co, int64_t Ci);
int dvc_conv3x3_pack_weights(const float* weight, const int64_t w_st[4],
                             int64_t Co, int64_t Ce, int64_t Cf, float* packed,
                             dvc_stream_t stream);
int dvc_warp_conv3x3_fwd(const float* feat, const float* flow,
                         const float* extra /*[opt]*/,
                         const float* packed_weight, const float* bias /*[opt]*/,
                         float* out_warp /*[opt]*/, float* out_conv, int64_t N,
                         int64_t Cf, int64_t Ce, int64_t Co, int64_t H,
                         int64_t W, const int64_t flow_st[4],
                         int64_t flow_downscale, int flags,
                         dvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * Piece 2: quantisation.  quantize_ste forward, dmc/models/utils.py:149-152
 * (round half to even).  [opt] offset[C] implements the hyper-latent form
 * z_hat = round(z - med_c) + med_c (video_model.py:222-224); pass NULL for the
 * plain round.  offset_st is the element stride between channels of `offset`.
 * ------------------------------------------------------------------------- */
int dvc_quantize_fwd(const float* x, const float* offset, float* q, int64_t N,
                     int64_t C, int64_t H, int64_t W, const int64_t x_st[4],
                     int64_t offset_st, const int64_t q_st[4],
                     dvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * Checkerboard dual prior, stage A.  Replaces video_model.py:176-189 (==
 * :348-361): writes the spatial-prior conv input
 *   params[N,3C,H,W] = cat(y_hat_00, y_hat_11, means, scales).
 * ------------------------------------------------------------------------- */
int dvc_dual_prior_stage_a_fwd(const float* y, const float* means,
                               const float* scales, float* params, int64_t N,
                               int64_t C, int64_t H, int64_t W,
                               const int64_t y_st[4], const int64_t means_st[4],
                               const int64_t scales_st[4],
                               const int64_t params_st[4], dvc_stream_t stream);

/* Rate accumulation workspace: one per likelihood kernel launch in flight.
 * Layout: double partial[DVC_RATE_MAX_BLOCKS * N] followed by uint32 ticket[N]
 * (the ticket words must be zero before first use; kernels reset them). */
#define DVC_RATE_MAX_BLOCKS 1024
int64_t dvc_rate_workspace_bytes(int64_t N);

/* ---------------------------------------------------------------------------
 * Stage B + Gaussian conditional + rate partial.  Replaces
 * video_model.py:192-207 (stage B + merge), the GaussianConditional call at
 * :232 / :405 (CompressAI GaussianConditional.forward: quantise, |v|, scale
 * lower bound, erfc CDF difference, likelihood lower bound) and the
 * log(p).sum(dim=(1,2,3)) of dmc/train.py:83.
 *   y, means, scales [N,C,H,W]; prior [N,2C,H,W] = y_spatial_prior(params)
 *   noise  [opt] [N,C,H,W] U(-1/2,1/2): training-mode likelihood on y+noise
 *   y_hat, means_hat, scales_hat, lik [opt] [N,C,H,W]
 *   q_w0,q_w1,s_w0,s_w1 [opt] [N,C/2,H,W]  (mode='compress', :209-214)
 *   logsum [opt] double[N]  = sum over C,H,W of ln(lik); needs `workspace`
 * ------------------------------------------------------------------------- */
int dvc_dual_prior_stage_b_gc_fwd(
    const float* y, const float* means, const float* scales, const float* prior,
    const float* noise, float* y_hat, float* means_hat, float* scales_hat,
    float* lik, float* q_w0, float* q_w1, float* s_w0, float* s_w1,
    double* logsum, void* workspace, int64_t N, int64_t C, int64_t H, int64_t W,
    const int64_t y_st[4], const int64_t means_st[4], const int64_t scales_st[4],
    const int64_t prior_st[4], const int64_t noise_st[4],
    const int64_t out_st[4] /* y_hat, means_hat, scales_hat, lik */,
    const int64_t half_st[4] /* q_w*, s_w* */, float scale_bound,
    float likelihood_bound, dvc_stream_t stream);

/* Module-level Gaussian conditional (no dual prior).  Replaces CompressAI
 * GaussianConditional.forward(inputs, scales, means, training).
 *   means [opt]; noise [opt] (training);  outputs [opt]; lik [opt]; logsum [opt] */
int dvc_gc_likelihood_fwd(const float* inputs, const float* scales,
                          const float* means, const float* noise, float* outputs,
                          float* lik, double* logsum, void* workspace, int64_t N,
                          int64_t C, int64_t H, int64_t W, const int64_t in_st[4],
                          const int64_t scales_st[4], const int64_t means_st[4],
                          const int64_t noise_st[4], const int64_t out_st[4],
                          float scale_bound, float likelihood_bound,
                          dvc_stream_t stream);
/* Backward of dvc_gc_likelihood_fwd (autograd of CompressAI
 * GaussianConditional.forward as the reference's loss.backward() runs it,
 * dmc/train.py:301).  Incoming: grad_lik [opt] [N,C,H,W], grad_logsum [opt]
 * double[N] (gradient of the fused sum ln p), grad_out [opt] (gradient of
 * `outputs`; only meaningful with noise).  Outgoing (all [opt], strides
 * grad_st): grad_inputs, grad_scales, grad_means.
 * LowerBound rule: pass iff (x >= bound) or (grad < 0).  In eval mode the
 * rounding has zero gradient, so only grad_scales is non-zero; with noise the
 * gradient reaches inputs and means as well. */
int dvc_gc_likelihood_bwd(const float* grad_lik, const double* grad_logsum,
                          const float* grad_out, const float* inputs,
                          const float* scales, const float* means,
                          const float* noise, float* grad_inputs,
                          float* grad_scales, float* grad_means, int64_t N,
                          int64_t C, int64_t H, int64_t W, const int64_t in_st[4],
                          const int64_t scales_st[4], const int64_t means_st[4],
                          const int64_t noise_st[4], const int64_t glik_st[4],
                          const int64_t gout_st[4], const int64_t grad_st[4],
                          float scale_bound, float likelihood_bound,
                          dvc_stream_t stream);

/* Backward of dvc_dual_prior_stage_a_fwd: grad_params [N,3C,H,W] ->
 * grad_y (= checkerboard-selected STE gradient), grad_means, grad_scales. */
int dvc_dual_prior_stage_a_bwd(const float* grad_params, float* grad_y,
                               float* grad_means, float* grad_scales, int64_t N,
                               int64_t C, int64_t H, int64_t W,
                               const int64_t gparams_st[4],
                               const int64_t grad_st[4], dvc_stream_t stream);

/* Backward of dvc_dual_prior_stage_b_gc_fwd.  Incoming (all [opt], strides
 * gin_st): grad_y_hat, grad_means_hat, grad_scales_hat, grad_lik; grad_logsum
 * [opt] double[N].  Outgoing (all [opt]): grad_y, grad_means, grad_scales
 * (strides grad_st; zero where the spatial prior was selected) and grad_prior
 * [N,2C,H,W] (zero where the first prior was selected). */
int dvc_dual_prior_stage_b_gc_bwd(
    const float* grad_y_hat, const float* grad_means_hat,
    const float* grad_scales_hat, const float* grad_lik,
    const double* grad_logsum, const float* y, const float* means,
    const float* scales, const float* prior, const float* noise, float* grad_y,
    float* grad_means, float* grad_scales, float* grad_prior, int64_t N,
    int64_t C, int64_t H, int64_t W, const int64_t y_st[4],
    const int64_t means_st[4], const int64_t scales_st[4],
    const int64_t prior_st[4], const int64_t noise_st[4],
    const int64_t gin_st[4], const int64_t grad_st[4],
    const int64_t gprior_st[4], float scale_bound, float likelihood_bound,
    dvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * Factorised entropy bottleneck.  Replaces CompressAI EntropyBottleneck.forward
 * (call sites video_model.py:220, :392) fused with the z quantisation of
 * :222-224 and the rate partial.  Parameters are the module's own tensors:
 *   matrices  float[C*33]  (_matrix0..4 concatenated per channel: 3,9,9,9,3)
 *   biases    float[C*13]  (_bias0..4: 3,3,3,3,1)
 *   factors   float[C*12]  (_factor0..3: 3 each)
 *   medians   float[C]     (quantiles[:,0,1])
 * i.e. filters=(3,3,3,3) (the only configuration the reference constructs).
 *   z [N,C,H,W]; noise [opt]; outputs [opt] (= z+noise or round(z-med)+med);
 *   z_hat [opt] (= round(z-med)+med always); lik [opt]; logsum [opt].
 * ------------------------------------------------------------------------- */
int dvc_eb_likelihood_fwd(const float* z, const float* noise,
                          const float* matrices, const float* biases,
                          const float* factors, const float* medians,
                          float* outputs, float* z_hat, float* lik,
                          double* logsum, void* workspace, int64_t N, int64_t C,
                          int64_t H, int64_t W, const int64_t z_st[4],
                          const int64_t noise_st[4], const int64_t out_st[4],
                          float likelihood_bound, dvc_stream_t stream);

/* Backward of dvc_eb_likelihood_fwd.  Incoming (all [opt], strides gin_st):
 * grad_outputs, grad_z_hat, grad_lik; grad_logsum [opt] double[N].  Outgoing
 * (all [opt]): grad_z (strides gz_st), grad_matrices[C*33], grad_biases[C*13],
 * grad_factors[C*12] (w.r.t. the raw parameters, i.e. through softplus/tanh),
 * grad_medians[C] (non-zero in eval mode only). */
int dvc_eb_likelihood_bwd(const float* grad_outputs, const float* grad_z_hat,
                          const float* grad_lik, const double* grad_logsum,
                          const float* z, const float* noise,
                          const float* matrices, const float* biases,
                          const float* factors, const float* medians,
                          float* grad_z, float* grad_matrices,
                          float* grad_biases, float* grad_factors,
                          float* grad_medians, int64_t N, int64_t C, int64_t H,
                          int64_t W, const int64_t z_st[4],
                          const int64_t noise_st[4], const int64_t gin_st[4],
                          const int64_t gz_st[4], float likelihood_bound,
                          dvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * Piece 4: rate.  Replaces collect_likelihoods_list, dmc/train.py:74-93.
 *   logsums  double[K*N]  K per-tensor ln-likelihood sums (from the kernels
 *            above, or from dvc_log_sum_fwd for a likelihood tensor)
 *   bpp      float[K*N]   = logsum / (-ln2 * num_pixels)
 *   bpp_total float[N]    = sum over K (left to right, as train.py:85)
 *   bits     [opt] double[N] = -sum_k logsum / ln2   (bits per sample)
 * ------------------------------------------------------------------------- */
int dvc_rate_finalize(const double* logsums, int K, int64_t N,
                      double num_pixels, float* bpp, float* bpp_total,
                      double* bits, dvc_stream_t stream);
/* log(p).sum(dim=(1,2,3)) of an arbitrary likelihood tensor (module boundary,
 * when the likelihood was not produced by one of the fused kernels). */
int dvc_log_sum_fwd(const float* lik, double* logsum, void* workspace,
                    int64_t N, int64_t C, int64_t H, int64_t W,
                    const int64_t lik_st[4], dvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * SURVEY.md 8f row f1 (first "next" row): preparation of the real entropy
 * coder's inputs.  Replaces, in one launch,
 *   GaussianConditional.build_indexes(scales)  (CompressAI; call sites
 *       video_model.py:248-249, 272, 282, 422-423, 447, 457):
 *       s = max(scales, scale_bound); index = (T-1) - #{k < T-1 : s <= table[k]}
 *       -- the reference does it in 63 compare-and-subtract passes;
 *   EntropyModel.quantize(x, "symbols")        (CompressAI, used by compress):
 *       symbol = int32(round(x [- means])).
 *   scales [opt] [N,C,H,W] -> indexes [opt] int32 (contiguous NCHW order)
 *   x      [opt] [N,C,H,W], means [opt]  -> symbols [opt] int32 (contiguous)
 *   table: device float[T], ascending (exp(linspace(ln 0.11, ln 256, 64)),
 *   base_model.py:43-49), T <= 256.
 * ------------------------------------------------------------------------- */
int dvc_symbols_indexes_fwd(const float* x, const float* means,
                            const float* scales, const float* table, int64_t T,
                            int32_t* symbols, int32_t* indexes, int64_t N,
                            int64_t C, int64_t H, int64_t W,
                            const int64_t x_st[4], const int64_t means_st[4],
                            const int64_t scales_st[4], float scale_bound,
                            dvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * the bit streams.  Replaces CompressAI's CPU coder as the
 * reference reaches it through
 *   GaussianConditional.compress(inputs, indexes) / .decompress(strings, indexes)
 *       (video_model.py:250-251, 273, 283, 424-425, 448, 458)
 *   EntropyBottleneck.compress(x) / .decompress(strings, size)
 *       (video_model.py:238-239, 257, 411-412, 431)
 * i.e. EntropyModel.compress: symbols = int(round(x - means)); per sample
 * RansEncoder.encode_with_indexes(symbols, indexes, cdf, cdf_length, offset)
 * (rans64, 16-bit probabilities, 4-bit bypass nibbles), and the inverse.
 *
 * Layout of one sample's output (`stream_symbols` = S > 0): the L = C*H*W
 * symbols in NCHW order are cut into ceil(L/S) sub-streams; each sub-stream is
 * a complete stock rans64 stream over its slice; one warp encodes/decodes one
 * sub-stream.  Container, little-endian u32 words:
 *   'DVC1', L, S, n_streams, words[n_streams], sub-streams back to back.
 * S = 0: ONE raw stock stream, no header -- byte-identical to CompressAI's
 * RansEncoder.encode_with_indexes (interop mode; serial, slow).
 *
 * `lanes` = 32 (S > 0 and a multiple of 1024) selects the lane-interleaved layout, magic 'DVC3':
 * the 32 lanes of the warp that owns a sub-stream each carry a stock rans64
 * state (item k of the sub-stream -> lane k % 32, round k / 32; same per-symbol
 * arithmetic and bypass coding as the stock coder) and share one word stream,
 * renormalising lanes taking consecutive words in lane order.  Sub-stream:
 *   mask (bit l: lane l's final state has a high word), the 32 states (1 or 2
 *   words each, lane 0 first), then the words of rounds 0, 1, ... in decode order.
 * 32 chains cost 132..260 bytes instead of 32 x 12.  `skip_rows` [opt, lanes =
 * 32 only; magic 'DVS3']: one device byte per table row, != 0 where value 0
 * (table position -offset) holds >= 65528/65536 of the row's mass; both sides
 * derive it from the tables.  Positions are taken in chunks of 1024 = 32 groups
 * of 32.  Pass 1 of a chunk codes, group by group, the symbols of unmarked rows
 * and, for a group with k >= 1 marked positions, one flag "a marked symbol of
 * the group is not 0" with P(1) = 8k/65536; pass 2 codes the marked symbols of
 * the flagged groups with their own rows.  Unflagged marked symbols are 0 and
 * take no chain step.  Lossless; both layouts round-trip any int32 symbols.
 * skip_max_flagged >= 0 makes the marks adaptive: the encoder first counts the
 * groups whose flag would be set (each costs ~8-13 bits) and, above that
 * number, ignores skip_rows and writes a plain 'DVC3' (all on the device; the
 * decoder reads the magic).  -1: always use the marks.
 * lanes = 1: the 'DVC1' layout above.
 *
 * Symbols come from `symbols` (int32, contiguous [N][L]) or from `x` [N,C,H,W]
 * (strided) minus `means` [opt] (strided; 0-strides broadcast, e.g. per-channel
 * medians).  Table indexes come from `indexes` (int32, contiguous [N][L]), or
 * are derived from `scales` (strided) exactly like build_indexes, or -- both
 * NULL -- are the channel number (EntropyBottleneck._build_indexes).
 * CDF tables are the module buffers: cdf int32[n_cdf][cdf_stride]
 * (_quantized_cdf), cdf_size int32[n_cdf] (_cdf_length), offset int32[n_cdf]
 * (_offset), all on the device.
 *   out        device bytes, sample n at out + n*out_stride_bytes
 *   out_bytes  device int64[N]: container size of each sample; NEGATIVE
 *              (-needed) when out_stride_bytes was too small (nothing written)
 *   scratch    dvc_rans_scratch_bytes(N, L, S, lanes) bytes
 *   status     [opt] device int, set to 1 if an index fell outside [0, n_cdf)
 * dvc_rans_max_bytes(L, S, lanes) is the worst-case container size of one sample.
 * ------------------------------------------------------------------------- */
/* HOST pointers (setup time, once per model): CompressAI's
 * pmf_to_quantized_cdf as EntropyModel._pmf_to_cdf calls it from
 * GaussianConditional.update_scale_table / EntropyBottleneck.update
 * (video_model.py:669-677).  pmf float[n] -> cdf int32[n+1], cdf[0] = 0,
 * cdf[n] = 2^precision, strictly increasing. */
int dvc_pmf_to_quantized_cdf(const float* pmf, int64_t n, int precision,
                             int32_t* cdf);
int64_t dvc_rans_scratch_bytes(int64_t N, int64_t L, int64_t stream_symbols,
                               int lanes);
int64_t dvc_rans_max_bytes(int64_t L, int64_t stream_symbols, int lanes);
/* scratch the decoder needs (0 for lanes = 1); 16-byte aligned device memory */
int64_t dvc_rans_decode_scratch_bytes(int64_t N, int64_t L, int64_t stream_symbols,
                                      int lanes);
int dvc_rans_encode(const float* x, const float* means, const int32_t* symbols,
                    const int32_t* indexes, const float* scales,
                    const float* scale_table, int64_t T, float scale_bound,
                    const int32_t* cdf, const int32_t* cdf_size,
                    const int32_t* offset, int64_t n_cdf, int64_t cdf_stride,
                    uint8_t* out, int64_t out_stride_bytes, int64_t* out_bytes,
                    void* scratch, int* status, int64_t N, int64_t C, int64_t H,
                    int64_t W, const int64_t x_st[4], const int64_t means_st[4],
                    const int64_t scales_st[4], int64_t stream_symbols,
                    int lanes, const uint8_t* skip_rows, int64_t skip_max_flagged,
                    dvc_stream_t stream);
/* Inverse.  in: device bytes (sample n at in + n*in_stride_bytes, in_bytes
 * device int64[N] = size of each container).  Writes out [opt] = float(symbol)
 * + means [opt] (EntropyModel.dequantize; strided [N,C,H,W]) and/or
 * out_symbols [opt] int32 contiguous.  status [opt]: 1 bad index, 2 malformed
 * container (header/lengths inconsistent with L, S or in_bytes; lanes = 32
 * also: a sub-stream that does not end on the encoder's initial states with
 * every word consumed); a malformed container never reads out of bounds.
 * `lanes` / `skip_rows` must match the container's magic (the Python face
 * reads it).  cdf_pack [opt, lanes = 32, n_cdf <= 256]: the tables re-packed
 * for the decoder's shared memory, 16-byte aligned device memory, derived from
 * the tables by the caller (cdf_pack_entries = number of u16 CDF entries):
 *   u16 lut[n_cdf][418]  per key of the 16-bit cum: the first table position lo a
 *                        symbol with such a cum can start at (cdf[r][lo] <= cum).
 *                        416 keys, monotone in cum: cums closer than 2048 to 0
 *                        (keys 0..87) or to 65535 (415..328) by distance d: 8 keys
 *                        per octave of d -- 8 e + the three bits below the leading
 *                        one of d | 1, e = floor(log2(d | 1)); 80 + (cum >> 8)
 *                        (88..327) in between.  A key no cum maps to repeats the
 *                        next one, so entry k + 1 (+ 1) bounds the bracket of entry
 *                        k; entry 416 = cdf_size - 2, entry 417 = 0
 *                        (csrc/dvc_coder.cu::lut_key)
 *   u32 row_start[n_cdf] offset of every row in the array below
 *   u16 cdf[entries]     the rows back to back, (value - 1) mod 2^16 (so that
 *                        "cum >= value" is "cum > stored" and 65536 fits), each
 *                        row followed by 4 entries 0xffff; entries = sum of
 *                        (cdf_size + 4)
 * It never changes a result; without it (or when it exceeds 124 KB) the decoder
 * searches the tables in global memory.  scratch: lanes = 32 only,
 * dvc_rans_decode_scratch_bytes(N, L, S, lanes) bytes, 16-byte aligned. */
int dvc_rans_decode(const uint8_t* in, int64_t in_stride_bytes,
                    const int64_t* in_bytes, const int32_t* indexes,
                    const float* scales, const float* scale_table, int64_t T,
                    float scale_bound, const int32_t* cdf,
                    const int32_t* cdf_size, const int32_t* offset,
                    int64_t n_cdf, int64_t cdf_stride, const float* means,
                    float* out, int32_t* out_symbols, int* status, int64_t N,
                    int64_t C, int64_t H, int64_t W, const int64_t scales_st[4],
                    const int64_t means_st[4], const int64_t out_st[4],
                    int64_t stream_symbols, int cb_parity, int64_t cb_alt,
                    int lanes, const uint8_t* skip_rows, const uint16_t* cdf_pack,
                    int64_t cdf_pack_entries, void* scratch, dvc_stream_t stream);

/* ---------------------------------------------------------------------------
 * Decoder side of the checkerboard dual prior: the element-wise glue of
 * MotionContextModel.decompress / FrameContextModel.decompress
 * (video_model.py:259-289 == :433-464) around the spatial-prior conv.
 *
 * dvc_rans_decode(cb_parity, cb_alt): the reference builds the scale plane of a
 * decoding pass as `scales_a * mask_p + scales_b * mask_q` (:268, :281) before
 * build_indexes.  With cb_parity = 0/1 the decoder reads scales[...] where
 * (h + w) & 1 == cb_parity and scales[... + cb_alt] (element offset; the other
 * channel group) elsewhere -- no mask tensors, no plane.  cb_parity = -1: off.
 *
 * stage a (:274-277): params [N,3C,H,W] = cat((q0 + means_0) * mask_0,
 *     (q0 + means_1) * mask_1, means, scales)  -- the spatial-prior conv input.
 * stage b (:284-289): y_hat [N,C,H,W] from q0, q1, means and the conv output
 *     prior [N,2C,H,W] = (means_0', scales_0', means_1', scales_1').
 * q0, q1: decoded symbols, int32, contiguous [N, C/2, H, W] (dvc_rans_decode's
 * out_symbols).  C, H, W even.
 * ------------------------------------------------------------------------- */
int dvc_dual_prior_decode_stage_a(const int32_t* q0, const float* means,
                                  const float* scales, float* params, int64_t N,
                                  int64_t C, int64_t H, int64_t W,
                                  const int64_t means_st[4],
                                  const int64_t scales_st[4],
                                  const int64_t params_st[4],
                                  dvc_stream_t stream);
int dvc_dual_prior_decode_stage_b(const int32_t* q0, const int32_t* q1,
                                  const float* means, const float* prior,
                                  float* y_hat, int64_t N, int64_t C, int64_t H,
                                  int64_t W, const int64_t means_st[4],
                                  const int64_t prior_st[4],
                                  const int64_t y_hat_st[4],
                                  dvc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DVC_B200_H_ */
