// dvc_warp_conv.cu -- SURVEY.md row f3: the backward warp of a context feature
// fused into the 3x3 convolution that consumes it, as a tcgen05 implicit GEMM.
//
// Replaces, for one scale of DMC.motion_compensation + MultiScaleContextFusion,
//   context = flow_warp(ref_feature, mv)                          video_model.py:502-504
//   conv    = convK_out(torch.cat((context_up, context), dim=1))  video_model.py:55-61
// (dmc/models/video_model.py; flow_warp is layers.py:175-198).  `context` is
// still written (the fusion net adds it back as a residual, :64-66) but it is
// never re-read from HBM: the conv consumes the warped tile straight out of
// shared memory.
//
// GEMM view: D[pixel, co] = sum_{tap, ci} A_tap[pixel, ci] * W[tap, ci, co]
//   M = 128 consecutive pixels of one image row, N = 64 output channels,
//   K = 9 taps x Ci input channels, TF32 operands, fp32 accumulation in TMEM.
//   PyTorch's cuDNN convolution computes in TF32 by default
//   (torch.backends.cudnn.allow_tf32 = True), so this is the reference's own
//   arithmetic class; it is not bit-identical to cuDNN (different summation
//   order) -- tests compare against an fp64 convolution of TF32-truncated
//   operands and against the fp32 convolution with a TF32 tolerance.
//
// Tile: 4 output rows x 128 columns.  Shared memory holds, per pipeline stage,
// a 16-channel slice of the 6 x 130 input halo in the UMMA no-swizzle K-major
// layout [k-chunk of 4 channels][halo pixel][4 floats]: a pixel is 16 bytes, 8
// consecutive pixels are one 128-byte core matrix, so the operand view of tap
// (dy, dx) for output row r is the SAME buffer with the descriptor start
// address advanced by ((r + dy) * 130 + dx) * 16 bytes -- no im2col copies.
// The weights of a K = 8 step (9 taps x 8 ci x 64 co, 18 KB) arrive by one bulk copy.
//
// Warp roles (736 threads, persistent, one CTA per SM):
//   warps 0-16  producers: tap records of the tile's halo (op-for-op replay of
//               flow_warp), then per 16-channel slice either copy `extra`
//               (cp.async straight into the stage, completion reported to the
//               stage's mbarrier, no register round trip and no wait) or
//               gather + blend `feat` into the stage; interior pixels of the
//               warped slice are also written to `out_warp`;
//   warps 17-20 epilogue: TMEM -> registers -> + bias -> out_conv, overlapped
//               with the next tile through the second TMEM accumulator set;
//   warp 21     MMA issuer (one lane): 72 tcgen05.mma (M128 N64 K8) per slice;
//   warp 22     weight loader (one lane): one cp.async.bulk per K-step.
// A producer thread owns one (halo column, 4-channel k-chunk) pair and walks
// the 6 halo rows in two batches of 3 vertically adjacent pixels: the south
// taps of a row are the north taps of the next one, so back-to-back gathers of
// one warp can share their sectors in L1 (ncu: the kernel moves ~7 GB through L2
// for 2.1 GB of HBM traffic -- L2/L1 bandwidth, not HBM, is what the gathers cost).
#include "dvc_common.cuh"
#include "dvc_warp_math.cuh"

namespace dvc {
namespace wc {

constexpr int kTileW = 128;   // UMMA M
constexpr int kRows = 4;      // output rows per tile = accumulators per TMEM set
constexpr int kCo = 64;       // UMMA N
constexpr int kChunk = 16;    // input channels per pipeline stage (2 UMMA K-steps of 8)
constexpr int kHaloW = kTileW + 2;
constexpr int kHaloH = kRows + 2;
constexpr int kHaloPix = kHaloW * kHaloH;  // 780
// pixels per k-chunk plane: >= kHaloPix and == 2 (mod 8), so the 4 k-chunk
// lanes of two neighbouring pixels (one 8-lane store phase) hit 32 distinct banks
constexpr int kPlanePix = 786;
constexpr int kAStageBytes = 4 * kPlanePix * 16;  // 50304
constexpr int kBStageBytes = 9 * 2 * kCo * 16;    // 18432: the weights of ONE K = 8 step
// Pipeline shape, measured on B200 at 1080p (us; profiles/r01_warp_conv.md):
//   A x B stages, interleave:  2x2,0: 835   2x3,0: 857   2x3,1: 893   3x3,0: 980   3x3,1: 1005
// Deeper is SLOWER: the gathers live off the L1 that shared memory leaves over
// (156 KB of stages -> 92 KB of L1; 220 KB -> 28 KB), and two A stages cannot
// average a cheap `extra` slice with an expensive warped one anyway.
#ifndef WC_STAGES_A
#define WC_STAGES_A 2
#endif
#ifndef WC_STAGES_B
#define WC_STAGES_B 2
#endif
#ifndef WC_EPI_ST   // epilogue store flavour (write-once output)
// L1::no_allocate: the write-once conv output must not take L1 lines away from
// the tap gathers (measured 836 -> 797 us; evict-first ".cs" 832; the context
// stores already stream, and cache hints on the gathers themselves changed nothing)
#define WC_EPI_ST "st.global.L1::no_allocate.v8.f32"
#endif
constexpr int kStagesA = WC_STAGES_A, kStagesB = WC_STAGES_B;
constexpr int kProducerWarps = 17, kEpilogueWarps = 4;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kThreads = (kProducerWarps + kEpilogueWarps + 2) * 32;  // 736
constexpr int kTmemCols = 512;  // 2 sets x 4 accumulators x 64 columns

// shared memory map (bytes)
constexpr int kOffA = 0;
constexpr int kOffB = kOffA + kStagesA * kAStageBytes;
constexpr int kOffWgt = kOffB + kStagesB * kBStageBytes;  // float4[780]
constexpr int kOffPos = kOffWgt + kHaloPix * 16;          // int[780]
constexpr int kOffGidx = kOffPos + kHaloPix * 4;          // int[780]
constexpr int kOffBias = kOffGidx + kHaloPix * 4;         // float[64]
constexpr int kNumBars = 2 * kStagesA + 2 * kStagesB + 4;
constexpr int kOffBar = kOffBias + kCo * 4;               // uint64[kNumBars]
constexpr int kOffTmem = kOffBar + kNumBars * 8;          // uint32
constexpr int kSmemBytes = kOffTmem + 16;
static_assert(kOffBar % 8 == 0 && kOffWgt % 16 == 0 && kOffB % 16 == 0, "smem alignment");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

constexpr int kInterior = 1 << 30;  // gidx flag: the CTA owns this pixel of out_warp

// instruction descriptor: D fp32, A/B TF32, both K-major, N = 64, M = 128
constexpr uint32_t kIdescBase = (1u << 4) | (2u << 7) | (2u << 10) | ((kTileW >> 4) << 24);

struct Params {
  const float* feat;
  const float* flow;
  const float* extra;
  const float* wpack;
  const float* bias;
  float* out_warp;
  float* out_conv;
  WarpGeom g;
  int N, H, W, Cf, Ce;
  long long fl_n, fl_c, fl_h, fl_w;
  int flow_level;
  int tiles_x, tiles_y, n_tiles;
  int n_chunks_extra, n_chunks;
  int debug;
};

// Order in which the 16-channel slices are consumed: all `extra` slices, then all
// warped ones (alternating them was measured slower, see above).  Returns the
// slice index within its tensor.
__host__ __device__ inline int slice_of(int c, int n_extra, int n_feat, bool* is_extra) {
  (void)n_feat;
  *is_extra = c < n_extra;
  return c < n_extra ? c : c - n_extra;
}

#ifdef WC_PROFILE
// cycle accounting of CTA 0 (debug builds only): [0] MMA wait-full in extra slices, [1] in
// warped slices, [2] MMA total, [3] producer wait-empty extra, [4] warped, [5] producer fill
// time of warped slices (acquire -> arrive), [6] producer total, [7] taps phase
__device__ unsigned long long g_wc_prof[12];
#define WC_T0() const long long _t0 = clock64()
#define WC_ACC(i) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) atomicAdd(&g_wc_prof[i], (unsigned long long)(clock64() - _t0)); } while (0)
#else
#define WC_T0()
#define WC_ACC(i)
#endif

// ---- PTX wrappers ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(20000u)   // may sleep up to 20 us; woken by the phase flip
      : "memory");
  return ok != 0;
}
// A wait that cannot hang the device: a lost arrival traps instead (the host
// sees cudaErrorLaunchFailure).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 20)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
// One lane of a converged warp; the compiler knows the guarded region is
// single-threaded, so tcgen05 operands stay in uniform registers without a
// per-lane waterfall loop.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void producer_bar() {
  asm volatile("bar.sync 1, %0;" ::"n"(kProducerThreads) : "memory");
}
// 32 lanes x 32 consecutive columns of one TMEM lane quarter -> 32 registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, no swizzle, K-major: rows (pixels / output
// channels) 16 bytes apart, 8-row core matrices `sbo` bytes apart, the two
// 16-byte halves of a K = 8 step `lbo` bytes apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) |
         ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

struct Tile {
  int n, y0, x0;
};
__device__ __forceinline__ Tile tile_of(const Params& p, int tile) {
  Tile t;
  const int tx = tile % p.tiles_x;
  const int rest = tile / p.tiles_x;
  t.x0 = tx * kTileW;
  t.y0 = (rest % p.tiles_y) * kRows;
  t.n = rest / p.tiles_y;
  return t;
}

// ---------------------------------------------------------------------------
// producers
// ---------------------------------------------------------------------------
__device__ __forceinline__ void compute_taps(const Params& p, const Tile& t, int ptid,
                                             float4* s_wgt, int* s_pos, int* s_gidx) {
  const int Cf4 = p.Cf >> 2;
  const float* __restrict__ fl = p.flow + t.n * p.fl_n;
  for (int q = ptid; q < kHaloPix; q += kProducerThreads) {
    const int hy = q / kHaloW, hx = q - hy * kHaloW;
    const int y = t.y0 - 1 + hy, x = t.x0 - 1 + hx;
    int gi = -1, pos = 0;
    float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
      gi = y * p.W + x;
      if (hy >= 1 && hy <= kRows && hx >= 1 && hx <= kTileW) gi |= kInterior;
      const float fx = fetch_flow(fl, p.fl_h, p.fl_w, y, x, p.flow_level);
      const float fy = fetch_flow(fl + p.fl_c, p.fl_h, p.fl_w, y, x, p.flow_level);
      const Taps T = make_taps(p.g, y, x, fx, fy);
      w = make_float4(T.nw, T.ne, T.sw, T.se);
      pos = (int)((unsigned)((T.y0 * p.W + T.x0) * Cf4) | (T.dx ? kEastIn : 0u) |
                  (T.dy ? kSouthIn : 0u));
    }
    s_wgt[q] = w;
    s_pos[q] = pos;
    s_gidx[q] = gi;
  }
}

// A producer thread owns the pair (halo column, k-chunk) = (ptid >> 2, ptid & 3)
// and the 6 halo pixels of that column; 130 x 4 = 520 of the 544 threads work.
constexpr int kFillThreads = kHaloW * 4;
static_assert(kFillThreads <= kProducerThreads, "producer mapping");

// The tap records of a thread's 6 pixels are read from shared memory once per
// tile and kept in registers: the load phase of a slice is address arithmetic
// + LDG only.
struct Items {
  int pos[kHaloH];   // float4 offset of the nw tap | kEastIn | kSouthIn
  int gi[kHaloH];    // pixel index | kInterior, or -1 outside the image / idle thread
};

__device__ __forceinline__ void load_items(Items& it, int ptid, const int* s_pos,
                                           const int* s_gidx) {
  const int col = ptid < kFillThreads ? (ptid >> 2) : 0;
#pragma unroll
  for (int r = 0; r < kHaloH; ++r) {
    it.pos[r] = s_pos[r * kHaloW + col];
    it.gi[r] = ptid < kFillThreads ? s_gidx[r * kHaloW + col] : -1;
  }
}

__device__ __forceinline__ uint32_t item_dst(uint32_t a_stage, int ptid, int r) {
  return a_stage + (uint32_t)((ptid & 3) * kPlanePix + r * kHaloW + (ptid >> 2)) * 16u;
}

__device__ __forceinline__ void fill_extra(const Params& p, const Tile& t, int chunk, int ptid,
                                           uint32_t a_stage, const Items& it) {
  if (ptid >= kFillThreads) return;
  const int Ce4 = p.Ce >> 2;
  const float4* __restrict__ ex = reinterpret_cast<const float4*>(p.extra) +
                                  (long long)t.n * p.H * p.W * Ce4 + chunk * 4 + (ptid & 3);
#pragma unroll
  for (int r = 0; r < kHaloH; ++r) {
    const int gi = it.gi[r];
    // outside the image: src-size 0 -> 16 bytes of zeros (the conv's zero padding)
    const float4* src = ex + (gi >= 0 ? (long long)(gi & (kInterior - 1)) * Ce4 : 0);
    const uint32_t nbytes = gi >= 0 ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(
                     item_dst(a_stage, ptid, r)),
                 "l"(src), "r"(nbytes)
                 : "memory");
  }
}

// Rows of gathers in flight ahead of the row being blended.  2 would hide more
// latency but needs 48 registers of tap data: with 736 threads the cap is 80
// registers and the spills cost more than the latency (measured 777 us at
// distance 1, 885 us at distance 2; batches without pipelining 801 us).
// The un-warped half of K for one tile.  cp.async (LDGSTS) straight into the
// stage looked ideal and was the bottleneck of this phase: the LSU throttles it
// to ~21 B/clk per SM with this 16-bytes-per-lane pattern, and a slice could only
// be requested after its stage had been released, so the tensor core waited for a
// third of the phase (in-kernel cycle accounting, tools/_wc_prof.py).  Plain
// LDG.128 reach 2.3x that rate, and a row costs only 4 registers here, so the
// copy is a register pipeline a whole slice deep: the 6 rows of slice c + 1 are
// in flight while slice c is stored, the loads do not wait for the stage (only
// the STS do), and when the stage frees up the data is already in registers.
__device__ __forceinline__ float4 ldg4_stream(const float4* p) {
  float4 v;   // read-once data: do not take L1 lines away from the tap gathers
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ void fill_extra_tile(const Params& p, const Tile& t, int ptid,
                                                uint32_t s_base, uint32_t bar_full0,
                                                uint32_t bar_empty0, int& sa, uint32_t& pha,
                                                const Items& it) {
  const int n_e = p.n_chunks_extra;
  if (n_e <= 0) return;
  const int Ce4 = p.Ce >> 2;
  const float4* __restrict__ ex = reinterpret_cast<const float4*>(p.extra) +
                                  (long long)t.n * p.H * p.W * Ce4 + (ptid & 3);
  const bool active = ptid < kFillThreads;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* src[kHaloH];
#pragma unroll
  for (int r = 0; r < kHaloH; ++r)
    src[r] = ex + (it.gi[r] >= 0 ? (long long)(it.gi[r] & (kInterior - 1)) * Ce4 : 0);
  const uint32_t t_off = (uint32_t)((ptid & 3) * kPlanePix + (ptid >> 2)) * 16u;
  float4 v[kHaloH];
#pragma unroll
  for (int r = 0; r < kHaloH; ++r) v[r] = it.gi[r] >= 0 ? ldg4_stream(src[r]) : z;
#pragma unroll 1
  for (int c = 0; c < n_e; ++c) {
    { WC_T0(); mbar_wait(bar_empty0 + 8u * sa, pha ^ 1); if (ptid < 32) WC_ACC(3); }
    const uint32_t dst = s_base + kOffA + sa * kAStageBytes + t_off;
    {
    WC_T0();
#pragma unroll
    for (int r = 0; r < kHaloH; ++r) {
      if (active)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(
                         dst + (uint32_t)(r * kHaloW * 16)),
                     "f"(v[r].x), "f"(v[r].y), "f"(v[r].z), "f"(v[r].w)
                     : "memory");
      // outside the image: zeros (the conv's zero padding)
      if (c + 1 < n_e) v[r] = it.gi[r] >= 0 ? ldg4_stream(src[r] + (c + 1) * 4) : z;
    }
    if (ptid < 32) WC_ACC(8);
    }
    // generic-proxy stores -> visible to the tensor core's async proxy; ONE arrival per
    // warp (544 per-thread arrivals on one mbarrier serialise for over a microsecond)
    {
    WC_T0();
    fence_proxy_async();
    if (ptid < 32) WC_ACC(9);
    }
    {
    WC_T0();
    __syncwarp();
    if ((ptid & 31) == 0) mbar_arrive(bar_full0 + 8u * sa);
    if (ptid < 32) WC_ACC(10);
    }
    if (++sa == kStagesA) { sa = 0; pha ^= 1; }
  }
}

#ifndef WC_DIST
#define WC_DIST 1
#endif
constexpr int kDist = WC_DIST;   // rows of gathers in flight ahead of the row being blended
static_assert(kHaloH % (kDist + 1) == 0, "row buffers must line up across slices");

// One halo pixel's four taps of a 4-channel group, in flight or landed.
struct RowBuf {
  float4 a[4];
};

// Issue the gathers of one (pixel, k-chunk) item; no wait.
__device__ __forceinline__ void issue_row(const float4* __restrict__ im, int Cf4, int south,
                                          int pos_i, int gi, RowBuf& b) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  b.a[0] = b.a[1] = b.a[2] = b.a[3] = z;
  if (gi >= 0) {
    const unsigned pos = (unsigned)pos_i;
    const float4* __restrict__ north = im + (pos & kOffMask);
    const bool e = (pos & kEastIn) != 0, sth = (pos & kSouthIn) != 0;
    // ATen skips out-of-bounds taps (their weight is 0 anyway)
    b.a[0] = ldg4(north);
    if (e) b.a[1] = ldg4(north + Cf4);
    if (sth) b.a[2] = ldg4(north + south);
    if (e && sth) b.a[3] = ldg4(north + south + Cf4);
  }
}

// The warped half of K for one tile.  A thread walks its column's 6 halo rows
// slice after slice as ONE software pipeline: the gathers of row r + 2 are issued
// before row r is blended and stored, across slice boundaries too (loads only
// need registers; the stage has to be free for the STS alone), so two rows of
// gathers are always in flight and the L2 / DRAM latency hides behind the
// blending and behind the wait for the stage instead of being paid per batch.
// (Sharing the east taps of neighbouring columns by warp shuffle was also tried:
// it removes 44 % of the loads of a rigid flow and was SLOWER, 807-921 vs 762-787
// us -- the gathers are latency-, not bandwidth-limited; an i.i.d. flow with no
// reuse at all costs only 9 % more than a rigid one.)
__device__ __forceinline__ void fill_warped_tile(const Params& p, const Tile& t, int ptid,
                                                 uint32_t s_base, uint32_t bar_full0,
                                                 uint32_t bar_empty0, int& sa, uint32_t& pha,
                                                 const float4* s_wgt, const Items& it) {
  const int n_w = p.n_chunks - p.n_chunks_extra;
  if (n_w <= 0) return;
  const int Cf4 = p.Cf >> 2;
  const long long sample = (long long)t.n * p.H * p.W * Cf4;
  const float4* __restrict__ im = reinterpret_cast<const float4*>(p.feat) + sample + (ptid & 3);
  float4* __restrict__ ow =
      p.out_warp ? reinterpret_cast<float4*>(p.out_warp) + sample + (ptid & 3) : nullptr;
  const int south = p.W * Cf4;
  const int col = ptid < kFillThreads ? (ptid >> 2) : 0;
  const bool active = ptid < kFillThreads;
  RowBuf buf[kDist + 1];
#pragma unroll
  for (int r = 0; r < kDist; ++r) issue_row(im, Cf4, south, it.pos[r], it.gi[r], buf[r]);
#pragma unroll 1
  for (int wc = 0; wc < n_w; ++wc) {
    { WC_T0(); mbar_wait(bar_empty0 + 8u * sa, pha ^ 1); if (ptid < 32) WC_ACC(4); }
    WC_T0();
    const uint32_t a_stage = s_base + kOffA + sa * kAStageBytes;
    const float4* __restrict__ im_c = im + wc * 4;
#pragma unroll
    for (int r = 0; r < kHaloH; ++r) {
      if (r + kDist < kHaloH)
        issue_row(im_c, Cf4, south, it.pos[r + kDist], it.gi[r + kDist],
                  buf[(r + kDist) % (kDist + 1)]);
      else if (wc + 1 < n_w)
        issue_row(im_c + 4, Cf4, south, it.pos[r + kDist - kHaloH], it.gi[r + kDist - kHaloH],
                  buf[(r + kDist) % (kDist + 1)]);
      if (active) {
        const RowBuf& b = buf[r % (kDist + 1)];
        const float4 v = blend4(b.a[0], b.a[1], b.a[2], b.a[3], s_wgt[r * kHaloW + col]);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(
                         item_dst(a_stage, ptid, r)),
                     "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
        const int gi = it.gi[r];
        if (ow != nullptr && gi >= 0 && (gi & kInterior))
          st_streaming(ow + (long long)(gi & (kInterior - 1)) * Cf4 + wc * 4, v);
      }
    }
    // generic-proxy stores -> visible to the tensor core's async proxy; ONE arrival per
    // warp (544 per-thread arrivals on one mbarrier serialise for over a microsecond)
    fence_proxy_async();
    __syncwarp();
    if ((ptid & 31) == 0) mbar_arrive(bar_full0 + 8u * sa);
    if (ptid < 32) WC_ACC(5);
    if (++sa == kStagesA) { sa = 0; pha ^= 1; }
  }
}

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
warp_conv3x3_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s_base = smem_u32(smem);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kOffTmem);
  float* s_bias = reinterpret_cast<float*>(smem + kOffBias);
  // barriers: fullA[3], emptyA[3], fullB[3], emptyB[3], tmem_full[2], tmem_empty[2]
  const uint32_t bar = s_base + kOffBar;
  auto bar_full_a = [&](int s) { return bar + 8u * s; };
  auto bar_empty_a = [&](int s) { return bar + 8u * (kStagesA + s); };
  auto bar_full_b = [&](int s) { return bar + 8u * (2 * kStagesA + s); };
  auto bar_empty_b = [&](int s) { return bar + 8u * (2 * kStagesA + kStagesB + s); };
  auto bar_tfull = [&](int b) { return bar + 8u * (2 * kStagesA + 2 * kStagesB + b); };
  auto bar_tempty = [&](int b) { return bar + 8u * (2 * kStagesA + 2 * kStagesB + 2 + b); };

  if (threadIdx.x < kCo) s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  if (warp == kProducerWarps + kEpilogueWarps) {
    if (lane == 0) {
      for (int s = 0; s < kStagesA; ++s) {
        mbar_init(bar_full_a(s), kProducerWarps);
        mbar_init(bar_empty_a(s), 1);                // tcgen05.commit
      }
      for (int s = 0; s < kStagesB; ++s) {
        mbar_init(bar_full_b(s), 1);                 // the weight loader (+ bulk-copy bytes)
        mbar_init(bar_empty_b(s), 1);                // tcgen05.commit
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(bar_tfull(b), 1);                  // tcgen05.commit
        mbar_init(bar_tempty(b), kEpilogueWarps);
      }
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(s_tmem)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp < kProducerWarps) {
    // ===================== producers =====================
    const int ptid = threadIdx.x;
    float4* s_wgt = reinterpret_cast<float4*>(smem + kOffWgt);
    int* s_pos = reinterpret_cast<int*>(smem + kOffPos);
    int* s_gidx = reinterpret_cast<int*>(smem + kOffGidx);
    int sa = 0;
    uint32_t pha = 0;
#ifdef WC_PROFILE
    const long long _role_t0 = clock64();
#endif
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      const Tile t = tile_of(p, tile);
      {
        WC_T0();
        producer_bar();  // everyone is done with the previous tile's tap records
        compute_taps(p, t, ptid, s_wgt, s_pos, s_gidx);
        producer_bar();
        if (warp == 0) WC_ACC(7);
      }
      Items items;
      load_items(items, ptid, s_pos, s_gidx);
      // un-warped half of K
      fill_extra_tile(p, t, ptid, s_base, bar_full_a(0), bar_empty_a(0), sa, pha, items);
      // warped half of K
      fill_warped_tile(p, t, ptid, s_base, bar_full_a(0), bar_empty_a(0), sa, pha, s_wgt, items);
    }
#ifdef WC_PROFILE
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_wc_prof[6], (unsigned long long)(clock64() - _role_t0));
#endif
  } else if (warp < kProducerWarps + kEpilogueWarps) {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tl) {
      const Tile t = tile_of(p, tile);
      const int buf = tl & 1;
      mbar_wait(bar_tfull(buf), (tl >> 1) & 1);
      tc_fence_after();
      const int x = t.x0 + q * 32 + lane;
#pragma unroll 1
      for (int r = 0; r < kRows; ++r) {
        const int y = t.y0 + r;
        const bool ok = (y < p.H) && (x < p.W);
        float4* __restrict__ o = reinterpret_cast<float4*>(p.out_conv) +
                                 (((long long)t.n * p.H + y) * p.W + x) * (kCo / 4);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) +
                                 (uint32_t)(buf * (kRows * kCo) + r * kCo + half * 32);
          tmem_ld32(taddr, v);
          tmem_ld_wait();
          if (ok) {
            // one 256-bit store per lane = one full 32-byte sector per request
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e)
                f[e] = __uint_as_float(v[8 * j + e]) + s_bias[half * 32 + 8 * j + e];
              asm volatile(
                  WC_EPI_ST " [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(
                      o + half * 8 + 2 * j),
                  "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]), "f"(f[4]), "f"(f[5]), "f"(f[6]),
                  "f"(f[7])
                  : "memory");
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty(buf));
    }
  } else if (warp == kProducerWarps + kEpilogueWarps) {
    // ===================== MMA issuer =====================
    uint32_t tl = 0, pha = 0, phb = 0;
    int sa = 0, sb = 0;
#ifdef WC_PROFILE
    const long long _role_t0 = clock64();
#endif
    const uint32_t lbo_a = (uint32_t)kPlanePix * 16u, sbo_a = 128u;
    const uint32_t lbo_b = 3u * (uint32_t)kCo * 16u, sbo_b = 128u;   // k-chunk planes of 192 rows
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tl) {
      const int buf = tl & 1;
      mbar_wait(bar_tempty(buf), ((tl >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(buf * (kRows * kCo));
      for (int c = 0; c < p.n_chunks; ++c) {
        { WC_T0(); mbar_wait(bar_full_a(sa), pha); WC_ACC(c < p.n_chunks_extra ? 0 : 1); }
        const uint64_t a0 = make_desc(s_base + kOffA + sa * kAStageBytes, lbo_a, sbo_a);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          mbar_wait(bar_full_b(sb), phb);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint64_t b0 = make_desc(s_base + kOffB + sb * kBStageBytes, lbo_b, sbo_b);
            // Row-tap fusion: halo row h (shifted by dx) is the A operand of every
            // (output row r, tap row dy) with r + dy = h.  The accumulators of
            // consecutive output rows are adjacent TMEM columns and the weights are
            // packed [dx][k-chunk][dy = 2, 1, 0][co], so those up to three products are
            // ONE tcgen05.mma with N = 64, 128 or 192: 18 instead of 36 per K-step, and
            // the 4 KB A operand is read from shared memory half as often.
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
              for (int i = 0; i < kHaloH; ++i) {
                // the very first pass of a tile must not mix fresh and touched accumulators
                // in one instruction: rows 0-2 via h = 2, row 3 via h = 5, then the rest
                constexpr int first_order[kHaloH] = {2, 5, 0, 1, 3, 4};
                const int h = first_order[i];
                const int r_min = h - 2 > 0 ? h - 2 : 0, r_max = h < kRows - 1 ? h : kRows - 1;
                const int n_rows = r_max - r_min + 1;
                const int dy_max = h - r_min;                  // tap row of output row r_min
                const uint64_t bd = b0 + (uint64_t)(dx * 2 * 3 * kCo + (2 - dy_max) * kCo);
                const uint64_t ad = a0 + (uint64_t)(ks * 2 * kPlanePix + h * kHaloW + dx);
                const uint32_t acc = (ks | dx) != 0 || i >= 2 ? 1u : (uint32_t)(c != 0);
                tc_mma_tf32(d0 + (uint32_t)(r_min * kCo), ad, bd,
                            kIdescBase | ((uint32_t)(n_rows * kCo >> 3) << 17), acc);
              }
            }
            tc_commit(bar_empty_b(sb));      // weight stage free when these MMAs retire
            if (ks == 1) {
              tc_commit(bar_empty_a(sa));    // activation stage free
              if (c == p.n_chunks - 1) tc_commit(bar_tfull(buf));  // accumulators complete
            }
          }
          __syncwarp();
          if (++sb == kStagesB) { sb = 0; phb ^= 1; }
        }
        if (++sa == kStagesA) { sa = 0; pha ^= 1; }
      }
    }
#ifdef WC_PROFILE
    if (blockIdx.x == 0 && lane == 0) atomicAdd(&g_wc_prof[2], (unsigned long long)(clock64() - _role_t0));
#endif
  } else {
    // ===================== weight loader =====================
    if (lane == 0) {
      int sb = 0;
      uint32_t phb = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int j = 0; j < 2 * p.n_chunks; ++j) {
          mbar_wait(bar_empty_b(sb), phb ^ 1);
          mbar_arrive_expect_tx(bar_full_b(sb), kBStageBytes);
          bulk_g2s(s_base + kOffB + sb * kBStageBytes,
                   reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)j * kBStageBytes,
                   kBStageBytes, bar_full_b(sb));
          if (++sb == kStagesB) { sb = 0; phb ^= 1; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kProducerWarps + kEpilogueWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(kTmemCols)
                 : "memory");
  }
}

// weight [Co=64, Ce+Cf, 3, 3] (element strides) ->
//   [slice in consumption order][K-step 2][dx 3][k-chunk 2][dy 2,1,0][co 64][4 channels]
__global__ void pack_weights_kernel(const float* __restrict__ w, long long s_co, long long s_ci,
                                    long long s_ky, long long s_kx, float* __restrict__ out,
                                    int n_extra, int n_feat, int total) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // [slice][K-step][dx 3][k-chunk 2][dy = 2, 1, 0][co 64][4 channels]
  const int j = idx & 3, co = (idx >> 2) & 63;
  int rest = idx >> 8;
  const int dyi = rest % 3;
  rest /= 3;
  const int kc = rest & 1;
  rest >>= 1;
  const int kx = rest % 3;
  rest /= 3;
  const int ks = rest & 1, c = rest >> 1;
  bool is_extra;
  const int slice = slice_of(c, n_extra, n_feat, &is_extra);
  const int ci = (is_extra ? 0 : n_extra * kChunk) + slice * kChunk + ks * 8 + kc * 4 + j;
  const int ky = 2 - dyi;
  out[idx] = w[co * s_co + ci * s_ci + ky * s_ky + kx * s_kx];
}

}  // namespace wc
}  // namespace dvc

using namespace dvc;

#ifdef WC_PROFILE
extern "C" int dvc_debug_warp_conv_profile(unsigned long long* out8, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out8, wc::g_wc_prof, sizeof(unsigned long long) * 12);
  if (reset) {
    unsigned long long z[12] = {0};
    cudaMemcpyToSymbol(wc::g_wc_prof, z, sizeof(z));
  }
  return 0;
}
#endif

extern "C" int64_t dvc_conv3x3_packed_weight_floats(int64_t Co, int64_t Ci) {
  if (Co != wc::kCo || Ci <= 0 || Ci % wc::kChunk) return 0;
  return Ci * 9 * Co;
}

extern "C" int dvc_conv3x3_pack_weights(const float* weight, const int64_t w_st[4], int64_t Co,
                                        int64_t Ce, int64_t Cf, float* packed,
                                        dvc_stream_t stream) {
  DVC_REQUIRE(weight && packed && w_st, "conv3x3_pack_weights: null pointer");
  DVC_REQUIRE(Co == wc::kCo, "conv3x3_pack_weights: Co must be 64 (got %lld)", (long long)Co);
  DVC_REQUIRE(Cf > 0 && Cf % wc::kChunk == 0 && Ce >= 0 && Ce % wc::kChunk == 0,
              "conv3x3_pack_weights: Ce and Cf must be multiples of 16");
  const int total = (int)((Ce + Cf) * 9 * Co);
  wc::pack_weights_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      weight, w_st[0], w_st[1], w_st[2], w_st[3], packed, (int)(Ce / wc::kChunk),
      (int)(Cf / wc::kChunk), total);
  return check_launch("pack_weights_kernel");
}

extern "C" int dvc_warp_conv3x3_fwd(const float* feat, const float* flow, const float* extra,
                                    const float* packed_weight, const float* bias,
                                    float* out_warp, float* out_conv, int64_t N, int64_t Cf,
                                    int64_t Ce, int64_t Co, int64_t H, int64_t W,
                                    const int64_t flow_st[4], int64_t flow_downscale, int flags,
                                    dvc_stream_t stream) {
  DVC_REQUIRE(feat && flow && packed_weight && out_conv && flow_st,
              "warp_conv3x3: null pointer");
  DVC_REQUIRE(N > 0 && H > 1 && W > 1, "warp_conv3x3: bad extents");
  DVC_REQUIRE(Co == wc::kCo, "warp_conv3x3: Co must be 64 (got %lld)", (long long)Co);
  DVC_REQUIRE(Cf > 0 && Cf % wc::kChunk == 0, "warp_conv3x3: Cf must be a multiple of 16");
  DVC_REQUIRE(Ce >= 0 && Ce % wc::kChunk == 0, "warp_conv3x3: Ce must be a multiple of 16");
  DVC_REQUIRE((Ce == 0) == (extra == nullptr), "warp_conv3x3: extra pointer / Ce mismatch");
  DVC_REQUIRE(flow_downscale >= 0 && flow_downscale <= 2, "warp_conv3x3: flow_downscale in 0..2");
  DVC_REQUIRE(aligned16(feat) && aligned16(extra) && aligned16(packed_weight) &&
                  aligned16(out_warp) && aligned16(out_conv),
              "warp_conv3x3: pointers must be 16-byte aligned");
  DVC_REQUIRE((long double)H * W * (Cf > Ce ? Cf : Ce) / 4 < 1073741824.0L,
              "warp_conv3x3: sample too large for 30-bit tap offsets");

  wc::Params p;
  p.feat = feat;
  p.flow = flow;
  p.extra = extra;
  p.wpack = packed_weight;
  p.bias = bias;
  p.out_warp = out_warp;
  p.out_conv = out_conv;
  fill_geom(p.g, H, W, flags);
  p.N = (int)N;
  p.H = (int)H;
  p.W = (int)W;
  p.Cf = (int)Cf;
  p.Ce = (int)Ce;
  p.fl_n = flow_st[0];
  p.fl_c = flow_st[1];
  p.fl_h = flow_st[2];
  p.fl_w = flow_st[3];
  p.flow_level = (int)flow_downscale;
  p.tiles_x = (int)((W + wc::kTileW - 1) / wc::kTileW);
  p.tiles_y = (int)((H + wc::kRows - 1) / wc::kRows);
  p.n_tiles = (int)(N * p.tiles_x * p.tiles_y);
  p.n_chunks_extra = (int)(Ce / wc::kChunk);
  p.n_chunks = (int)((Ce + Cf) / wc::kChunk);
  p.debug = (flags >> 8) & 0xff;

  cudaError_t e = cudaFuncSetAttribute(wc::warp_conv3x3_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       wc::kSmemBytes);
  if (e != cudaSuccess)
    return fail(DVC_ERR_CUDA, "warp_conv3x3: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
  wc::warp_conv3x3_kernel<<<grid, wc::kThreads, wc::kSmemBytes, (cudaStream_t)stream>>>(p);
  return check_launch("warp_conv3x3_kernel");
}
