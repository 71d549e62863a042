// dvc_warp.cu -- optical-flow backward warp + flow pyramid for sm_100a.
//
// Replaces (arithmetic replayed op for op):
//   flow_warp / torch_warp      /root/reference/dmc/models/layers.py:175-198
//   bilineardownsacling (+ /2)  /root/reference/dmc/models/layers.py:201-206,
//                               /root/reference/dmc/models/video_model.py:499-500
//   DMC.motion_compensation     /root/reference/dmc/models/video_model.py:497-504 (the 4 warps)
//
// The warp is a pure HBM-bound gather: 4*(2C+2) algorithmic bytes per output
// pixel.  Layout decides everything:
//   * channels_last (NHWC) fast path: one thread owns 4 channels of a pixel, a
//     pixel's tap is one contiguous C*4-byte run -> every tap is a coalesced
//     128-bit load, every output a coalesced 128-bit streaming store.  A CTA
//     walks a (256/(C/4)) x ROWS pixel tile row by row so the two source rows
//     a row needs are re-used from L1 by the next row; vertically adjacent
//     tiles meet in L2 (126 MB), so DRAM sees the input about once.
//   * strided path (NCHW, C=3 frames, odd layouts): one thread per pixel,
//     lanes along W (coalesced per channel plane), channel loop unrolled x4.
// Source coordinates are computed once per pixel and reused for all channels.
#include <cuda.h>     // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint
#include <stdlib.h>

#include "dvc_common.cuh"
#include "dvc_warp_math.cuh"

namespace dvc {

// ---------------------------------------------------------------------------
// task descriptors (passed by value as a __grid_constant__)
// ---------------------------------------------------------------------------
enum { kModeVec4 = 0, kModeStrided = 1, kModeVec4C64 = 2 };
constexpr int kThreads = 256;
constexpr int kRows = 8;         // tile height of both paths
constexpr int kStridedTileW = 32;
constexpr int kMaxTasks = 4;

struct WarpTask {
  const float* im;
  const float* flow;
  float* out;
  WarpGeom g;
  int N, C;
  long long im_n, im_c, im_h, im_w;
  long long fl_n, fl_c, fl_h, fl_w;
  long long out_n, out_c, out_h, out_w;
  int mode;
  int c4;            // vec4: float4 groups per pixel
  int ppb;           // pixel columns per CTA tile
  int rows;          // rows per CTA tile
  int flow_level;    // 0: flow has im's resolution; k: flow is 2^k finer, reduced on the fly
  int prefetch;      // vec4: 0 none, 1 south tap row if the flow is coherent (default), 2 both rows, 3 south always
  int tiles_x, tiles_y;
  int first_block, n_blocks;
  int pl_tma;        // planar path: 1 = boxes staged by TMA (fixed 64-float pitch, unaligned origin)
};
struct WarpBatch {
  WarpTask t[kMaxTasks];
  int n_tasks;
  // block schedule: `interleaved` CTAs in cycles of `period`, then per-task tails
  int interleaved, period;
  int quota[kMaxTasks], qprefix[kMaxTasks];
  int tail_first[kMaxTasks], tail_tile0[kMaxTasks];
};

// ---------------------------------------------------------------------------
// NHWC float4 path -- warp-private tiles, no block barrier
//
// A pixel's C channels are C/4 = c4 lanes of one warp (c4 divides 32), so a
// warp covers ppw = 32/c4 adjacent pixel columns and walks c4 rows: 32 pixels.
//   phase 1: each lane computes the source coordinate of ONE of those 32
//            pixels (column lane/c4, row lane%c4) and parks (nw-tap offset +
//            border flags, 4 weights) in the warp's slice of shared memory;
//   phase 2: lane = (column, 4-channel group); per row: 2 broadcast LDS,
//            4 x LDG.128 (the east taps are an immediate offset when the
//            pixel stride is dense), 16 FFMA, 1 streaming STG.128; two rows
//            (8 gathers) are in flight per lane.
// Only __syncwarp separates the phases, so the 8 warps of a CTA (8*ppw columns
// side by side, sharing source rows through L1) never wait for each other.
//
// History (ncu, 1080p motion-compensation launch, 1.49 GB algorithmic):
//   v1 coordinates recomputed in all c4 lanes        387 us, 236 M warp-inst, issue 55%
//   v2 taps in smem, CTA barrier, 8-row tiles        307 us, 144 M warp-inst, barrier+long-sb stalls
//   v3 this version: see profiles/
// ---------------------------------------------------------------------------
struct TapSmem {
  float4 wgt[kThreads];  // nw, ne, sw, se
  int pos[kThreads];     // float4 offset of the nw tap | east-in << 30 | south-in << 31
};

// gather the 4 taps of one pixel row; `east` is the float4 distance to the
// east neighbour (compile-time constant in the dense specialisation)
__device__ __forceinline__ void gather4(const float4* __restrict__ north,
                                        const float4* __restrict__ south, unsigned pos,
                                        int east, bool all_in, float4& v0, float4& v1,
                                        float4& v2, float4& v3) {
  const unsigned o = pos & kOffMask;
  if (all_in) {
    v0 = ldg4(north + o);
    v1 = ldg4(north + o + east);
    v2 = ldg4(south + o);
    v3 = ldg4(south + o + east);
  } else {  // border: ATen skips out-of-bounds taps (their weight is 0 anyway)
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool e = (pos & kEastIn) != 0, sth = (pos & kSouthIn) != 0;
    v0 = ldg4(north + o);
    v1 = e ? ldg4(north + o + east) : z;
    v2 = sth ? ldg4(south + o) : z;
    v3 = (e && sth) ? ldg4(south + o + east) : z;
  }
}

template <int C4T>  // C4T = 16: C = 64 with dense pixels; 0: runtime c4
__device__ __forceinline__ void warp_tile_vec4(const WarpTask& t, int tile, TapSmem& sm) {
  const int c4 = C4T ? C4T : t.c4;
  const int ppw = 32 / c4;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int H = t.g.H, W = t.g.W;
  const int tx = tile % t.tiles_x;
  const int rest = tile / t.tiles_x;
  const int ty = rest % t.tiles_y;
  const int n = rest / t.tiles_y;
  const int w0 = (tx * (kThreads / 32) + wid) * ppw, h0 = ty * c4;
  if (w0 >= W) return;  // warp-uniform: no CTA barrier below
  // strides in float4 units (dispatcher guarantees divisibility and int32 range)
  const int im_h4 = (int)(t.im_h >> 2);
  const int im_w4 = C4T ? C4T : (int)(t.im_w >> 2);

  // ---- phase 1: one tap record per lane --------------------------------------
  {
    const int pc = lane / c4, r = lane - pc * c4;
    const int h = min(h0 + r, H - 1), w = min(w0 + pc, W - 1);
    const float* __restrict__ fl = t.flow + n * t.fl_n;
    const float fx = fetch_flow(fl, t.fl_h, t.fl_w, h, w, t.flow_level);
    const float fy = fetch_flow(fl + t.fl_c, t.fl_h, t.fl_w, h, w, t.flow_level);
    const Taps T = make_taps(t.g, h, w, fx, fy);
    sm.wgt[threadIdx.x] = make_float4(T.nw, T.ne, T.sw, T.se);
    const int o = T.y0 * im_h4 + T.x0 * im_w4;
    sm.pos[threadIdx.x] = (int)((unsigned)o | (T.dx ? kEastIn : 0u) | (T.dy ? kSouthIn : 0u));
    // Fire-and-forget L2 prefetch of this pixel's south tap row (west+east tap =
    // one contiguous run of 2 pixels).  It costs no register and no scoreboard
    // slot, so it deepens the DRAM queue beyond what the register file can hold
    // as in-flight gathers; the demand loads of the tile's later rows (and, for a
    // coherent flow, the north taps of the row below) then hit in L2.  For an
    // incoherent flow (no tap shared between neighbouring pixels) the prefetch
    // only doubles the L2 requests, so the warp votes: prefetch iff most pixels
    // land within 2 px of where their upper neighbour's mapping predicts.
    if (t.prefetch) {
      const int up_o = __shfl_up_sync(0xffffffffu, o, 1);
      const int dxy = o - up_o - im_h4;                 // 0 for a locally rigid flow
      const bool coherent = (lane - pc * c4 == 0) || (abs(dxy) <= 2 * im_w4) ||
                            (abs(dxy - im_h4) <= 2 * im_w4) || (abs(dxy + im_h4) <= 2 * im_w4);
      const unsigned votes = __ballot_sync(0xffffffffu, coherent);
      if (t.prefetch > 2 || __popc(votes) >= 24) {
        const unsigned bytes = (unsigned)(c4 * 16) * (T.dx ? 2u : 1u);
        const float4* base = reinterpret_cast<const float4*>(t.im + n * t.im_n) + o;
        if (T.dy)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base + im_h4), "r"(bytes)
                       : "memory");
        if (t.prefetch == 2)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base), "r"(bytes)
                       : "memory");
      }
    }
  }
  __syncwarp();

  // ---- phase 2: gather + blend -----------------------------------------------
  const int px = lane / c4, grp = lane - px * c4;
  const int w = w0 + px;
  if (w >= W) return;
  const float4* __restrict__ north = reinterpret_cast<const float4*>(t.im + n * t.im_n) + grp;
  const float4* __restrict__ south = north + im_h4;
  const long long out_h4 = t.out_h >> 2;
  float4* __restrict__ out4 =
      reinterpret_cast<float4*>(t.out + n * t.out_n + (long long)w * t.out_w) + grp +
      h0 * out_h4;
  const int rows = min(c4, H - h0);
  const float4* __restrict__ wgt = sm.wgt + (wid << 5) + px * c4;
  const int* __restrict__ pos = sm.pos + (wid << 5) + px * c4;

#ifndef DVC_WARP_RPI
#define DVC_WARP_RPI 1   // rows (x4 gathers) in flight per lane; 1 row x 48 warps/SM measured best
#endif
  int r = 0;
#pragma unroll 1
  for (; r + DVC_WARP_RPI <= rows; r += DVC_WARP_RPI) {
    unsigned ps[DVC_WARP_RPI];
    float4 ws[DVC_WARP_RPI];
    unsigned all = 0xffffffffu;
#pragma unroll
    for (int k = 0; k < DVC_WARP_RPI; ++k) {
      ps[k] = (unsigned)pos[r + k];
      ws[k] = wgt[r + k];
      all &= ps[k];
    }
    const bool all_in = (all >> 30) == 3u;
    float4 v[DVC_WARP_RPI][4];
#pragma unroll
    for (int k = 0; k < DVC_WARP_RPI; ++k)
      gather4(north, south, ps[k], im_w4, all_in, v[k][0], v[k][1], v[k][2], v[k][3]);
#pragma unroll
    for (int k = 0; k < DVC_WARP_RPI; ++k)
      st_streaming(out4 + k * out_h4, blend4(v[k][0], v[k][1], v[k][2], v[k][3], ws[k]));
    out4 += DVC_WARP_RPI * out_h4;
  }
#pragma unroll 1
  for (; r < rows; ++r) {
    const unsigned pa = (unsigned)pos[r];
    float4 a0, a1, a2, a3;
    gather4(north, south, pa, im_w4, false, a0, a1, a2, a3);
    st_streaming(out4, blend4(a0, a1, a2, a3, wgt[r]));
    out4 += out_h4;
  }
}

// ---------------------------------------------------------------------------
// strided path: one thread per pixel, arbitrary element strides
// ---------------------------------------------------------------------------
__device__ __forceinline__ void warp_tile_strided(const WarpTask& t, int tile) {
  const int tx = tile % t.tiles_x;
  const int rest = tile / t.tiles_x;
  const int ty = rest % t.tiles_y;
  const int n = rest / t.tiles_y;
  const int w = tx * kStridedTileW + (threadIdx.x & 31);
  const int h = ty * kRows + (threadIdx.x >> 5);
  if (w >= t.g.W || h >= t.g.H) return;

  const float* __restrict__ fl = t.flow + n * t.fl_n;
  const float fx = fetch_flow(fl, t.fl_h, t.fl_w, h, w, t.flow_level);
  const float fy = fetch_flow(fl + t.fl_c, t.fl_h, t.fl_w, h, w, t.flow_level);
  const Taps T = make_taps(t.g, h, w, fx, fy);

  // out-of-bounds neighbours are skipped by ATen; they alias an in-bounds tap
  // here and are masked to exactly 0 so inf/NaN pixels cannot leak through
  // their zero weight.
  const bool in_e = T.dx != 0, in_s = T.dy != 0;
  // The kernel is issue bound for few channels (ncu: 400 instructions per pixel at C = 3,
  // issue slots 70 % busy), so the channel loop carries two running pointers and
  // the tap offsets as plain element offsets: one address computation per access.
  const float* __restrict__ p = t.im + n * t.im_n + T.y0 * t.im_h + T.x0 * t.im_w;
  float* __restrict__ po = t.out + n * t.out_n + h * t.out_h + w * t.out_w;
  const long long o_e = in_e ? t.im_w : 0, o_s = in_s ? t.im_h : 0, o_se = o_e + o_s;
  const long long ic = t.im_c, oc = t.out_c;
  auto one = [&]() {
    const float vnw = __ldg(p);
    float vne = __ldg(p + o_e), vsw = __ldg(p + o_s), vse = __ldg(p + o_se);
    vne = in_e ? vne : 0.f;
    vsw = in_s ? vsw : 0.f;
    vse = (in_e && in_s) ? vse : 0.f;
    *po = blend(vnw, vne, vsw, vse, T);
    p += ic;
    po += oc;
  };
  if (t.C == 3) {   // frames: x_ref and SpyNet's pyramid
    one();
    one();
    one();
    return;
  }
  int c = 0;
#pragma unroll 1
  for (; c + 4 <= t.C; c += 4) {
    one();
    one();
    one();
    one();
  }
#pragma unroll 1
  for (; c < t.C; ++c) one();
}

#ifndef DVC_WARP_MINB
#define DVC_WARP_MINB 6   // <= 40 registers -> 6 CTAs (48 warps) per SM
#endif
__global__ void __launch_bounds__(kThreads, DVC_WARP_MINB)
warp_multi_kernel(const __grid_constant__ WarpBatch batch) {
  // block -> (task, tile).  The tasks are INTERLEAVED in proportion to their
  // sizes (quota[k] CTAs of task k per cycle of `period` CTAs) instead of laid
  // end to end: the 3-channel frame warp is L1-wavefront bound (scalar gathers)
  // while the 64-channel warps are DRAM bound, so running them side by side
  // overlaps the two limits; end to end the frame warp ran alone at the tail.
  const int b = blockIdx.x;
  int k = 0, tile;
  if (b < batch.interleaved) {
    const int cyc = b / batch.period, r = b - cyc * batch.period;
#pragma unroll
    for (int i = 1; i < kMaxTasks; ++i)
      if (i < batch.n_tasks && r >= batch.qprefix[i]) k = i;
    tile = cyc * batch.quota[k] + (r - batch.qprefix[k]);
  } else {
#pragma unroll
    for (int i = 1; i < kMaxTasks; ++i)
      if (i < batch.n_tasks && b >= batch.tail_first[i]) k = i;
    tile = batch.tail_tile0[k] + (b - batch.tail_first[k]);
  }
  const WarpTask& t = batch.t[k];
  __shared__ TapSmem sm;
  if (t.mode == kModeVec4C64)
    warp_tile_vec4<16>(t, tile, sm);
  else if (t.mode == kModeVec4)
    warp_tile_vec4<0>(t, tile, sm);
  else
    warp_tile_strided(t, tile);
}


// ---------------------------------------------------------------------------
// planar (NCHW) path -- the reference's own memory format -- with the source
// tile staged in shared memory.
//
// The strided path above issues 4 scalar gathers per output element and is
// bound by L1 wavefronts (~0.36 of the HBM copy peak for C = 64).  A pixel's tap
// geometry is the same for every channel, so a CTA (32 x 8 pixels) computes its
// 256 tap records once, takes the bounding box of the taps, and then for every
// channel plane copies that box into shared memory with 16-byte cp.async
// (L2 -> smem, no registers, 4 planes in flight) and gathers from there.  If the
// box does not fit (incoherent flow) the CTA falls back to the global gathers.
// Same arithmetic as every other path (make_taps / blend): bit-identical.
// ---------------------------------------------------------------------------
#ifndef DVC_PLANAR_VERT
#define DVC_PLANAR_VERT 1   // a thread's two pixels are rows h, h + 8 of a 32 x 16 tile (0: columns w, w + 32 of a 64 x 8
                            // tile: wider boxes, measured 6-27 % slower on rough flows, equal on gentle ones)
#endif
#ifndef DVC_PLANAR_PX   // 3 / 4 pixels per thread (taller tiles): 3-5 % faster on gentle flows, 5-100 %
#define DVC_PLANAR_PX 2 // slower on rough ones (fewer tiles fit their box) -- measured, 2 kept
#endif
#ifndef DVC_PLANAR_PITCH
// Row pitch of the staged box in floats.  64 (= 0 mod 32 banks): a tap's bank is its COLUMN
// mod 32 whatever row it sits in, so the 32 lanes of a warp -- adjacent pixels, hence (for a
// coherent flow) adjacent tap columns spread over a few rows -- gather without bank conflicts.
// 0 = round 1's variable odd-float4 pitch (56 % of the gather wavefronts were conflicts).
#define DVC_PLANAR_PITCH 64
#endif
constexpr int kPlPx = DVC_PLANAR_PX;                  // pixels per thread
constexpr int kPlTileW = DVC_PLANAR_VERT ? 32 : 32 * kPlPx, kPlTileH = DVC_PLANAR_VERT ? 8 * kPlPx : 8;
constexpr int kPlBoxW4 = DVC_PLANAR_PITCH ? DVC_PLANAR_PITCH / 4 : (DVC_PLANAR_VERT ? 18 : 28);
constexpr int kPlBoxH = DVC_PLANAR_VERT ? 8 * kPlPx + 24 : 26;   // staged box limits
constexpr int kPlBufFloats = kPlBoxW4 * 4 * kPlBoxH;  // 2912 floats = 11648 B
#ifndef DVC_PLANAR_BUFS
#define DVC_PLANAR_BUFS 4
#endif
constexpr int kPlBufs = DVC_PLANAR_BUFS;              // channel planes in flight
constexpr int kPlSlots = (kPlBoxW4 * kPlBoxH + kThreads - 1) / kThreads;  // float4 per thread, 3
constexpr int kPlSmemBytes = kPlBufs * kPlBufFloats * 4;

// Box of a tile's taps -> does it fit the staging buffers?  Evaluated identically by the
// staged launch and by its complement (the gather launch), so no flag buffer is needed.
struct PlBox {
  int bx0, by0, hh, w4;   // origin, rows, width in float4 columns (cp.async variant)
  bool fits;
};
__device__ __forceinline__ PlBox planar_box(int xmin, int xmax, int ymin, int ymax, int H, int W,
                                            int tma) {
  PlBox b;
  const int bx1 = min(xmax + 1, W - 1), by1 = min(ymax + 1, H - 1);
  b.by0 = ymin;
  b.hh = by1 - ymin + 1;
  if (tma) {                       // TMA: any origin, 64 columns
    b.bx0 = xmin;
    b.w4 = 16;
    b.fits = (bx1 - xmin + 1) <= 64 && b.hh <= kPlBoxH;
  } else {                         // cp.async: 16-byte aligned origin
    b.bx0 = xmin & ~3;
    b.w4 = (bx1 - b.bx0 + 4) >> 2;
#if DVC_PLANAR_PITCH
    b.fits = b.w4 <= kPlBoxW4 && b.hh <= kPlBoxH;
#else
    b.fits = (((b.w4 & 1) ? b.w4 : b.w4 + 1) <= kPlBoxW4) && b.hh <= kPlBoxH;
#endif
  }
  return b;
}

// Per-thread state of the staged plane loop.  The loop is unrolled over the kPlBufs staging
// buffers so that every shared-memory address is `register + immediate` and the copy / gather
// of a plane cost a fixed, small instruction count (the first version re-derived tap indexes,
// buffer bases and border predicates per plane: 107 SASS instructions per plane and thread,
// 71 % of the issue slots; ncu, profiles/r02_planar.md).
struct PlanarLoop {
  const float* src[(kPlBoxW4 * kPlBoxH + kThreads - 1) / kThreads];
  uint32_t dst[(kPlBoxW4 * kPlBoxH + kThreads - 1) / kThreads];
  int i_nw[kPlPx], i_ne[kPlPx], i_sw[kPlPx], i_se[kPlPx];
  float* pk[kPlPx];
  Taps T[kPlPx];
  bool valid[kPlPx];
};

template <bool kFast>
__device__ __forceinline__ void planar_planes(PlanarLoop& L, float* __restrict__ smem, const int C,
                                              const long long im_c, const long long out_c) {
  constexpr int kSlots = (kPlBoxW4 * kPlBoxH + kThreads - 1) / kThreads;
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(smem);
  auto issue = [&](int c, int buf) {
    if (c < C) {
#pragma unroll
      for (int k = 0; k < kSlots; ++k) {
        if (L.dst[k] != 0xffffffffu)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                           s_base + L.dst[k] + (uint32_t)(buf * kPlBufFloats * 4)),
                       "l"(L.src[k])
                       : "memory");
        L.src[k] += im_c;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // empty groups keep the count uniform
  };
#pragma unroll
  for (int c = 0; c < kPlBufs - 1; ++c) issue(c, c);

#pragma unroll 1
  for (int c0 = 0; c0 < C; c0 += kPlBufs) {
#pragma unroll
    for (int u = 0; u < kPlBufs; ++u) {
      const int c = c0 + u;
      if (c < C) {                  // block-uniform
        asm volatile("cp.async.wait_group %0;" ::"n"(kPlBufs - 2) : "memory");   // plane c has landed
        __syncthreads();            // ... for every thread; and everyone is done with plane c - 1
        issue(c + kPlBufs - 1, (u + kPlBufs - 1) % kPlBufs);   // into the buffer plane c - 1 used
        const float* __restrict__ b = smem + u * kPlBufFloats;
#pragma unroll
        for (int k = 0; k < kPlPx; ++k) {
          const float vnw = b[L.i_nw[k]];
          float vne = b[L.i_ne[k]], vsw = b[L.i_sw[k]], vse = b[L.i_se[k]];
          if (!kFast) {
            const bool in_e = L.T[k].dx != 0, in_s = L.T[k].dy != 0;
            vne = in_e ? vne : 0.f;
            vsw = in_s ? vsw : 0.f;
            vse = (in_e && in_s) ? vse : 0.f;
          }
          if (kFast || L.valid[k]) __stcs(L.pk[k], blend(vnw, vne, vsw, vse, L.T[k]));
          L.pk[k] += out_c;
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// kStaged = true : tiles whose tap box fits are gathered from shared memory, others return;
// kStaged = false: the complementary launch (no shared memory -> all of L1 for the
//                  incoherent gathers) takes the tiles the first one left, others return.
// Both evaluate the same deterministic box test, so no flag buffer is needed.
#ifndef DVC_PLANAR_MINB
#define DVC_PLANAR_MINB 4
#endif
// Up to kMaxTasks planar problems in ONE launch (largest first): the half- and quarter-resolution
// context warps of a P-frame are 1.7 and 0.4 waves of CTAs on their own; laid behind the
// full-resolution one they fill its tail instead of leaving the GPU half empty three times.
struct PlanarBatch {
  WarpTask t[kMaxTasks];
  int n_tasks;
  int first[kMaxTasks + 1];   // first block of each task; first[kMaxTasks] = number of planar blocks
};

template <bool kStaged>
__global__ void __launch_bounds__(kThreads, DVC_PLANAR_MINB)
warp_planar_kernel(const __grid_constant__ PlanarBatch pb) {
  extern __shared__ __align__(16) float pl_smem[];
  __shared__ int s_red[32];   // static: the gather launch has no dynamic shared memory
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // (Letting the 3-channel frame warp ride along in the complement launch was measured and
  // dropped: 0.456-0.478 ms per NCHW step against 0.448 ms with its own warp_multi launch --
  // it runs at this kernel's 64-register occupancy and the early-exit tiles are not free.)
  int task = 0;
#pragma unroll
  for (int i = 1; i < kMaxTasks; ++i)
    if (i < pb.n_tasks && (int)blockIdx.x >= pb.first[i]) task = i;
  const WarpTask& t = pb.t[task];
  const int tile = blockIdx.x - pb.first[task];
  const int tx = tile % t.tiles_x;
  const int rest = tile / t.tiles_x;
  const int ty = rest % t.tiles_y;
  const int n = rest / t.tiles_y;
  const int H = t.g.H, W = t.g.W;
  const float* __restrict__ fl = t.flow + n * t.fl_n;

  Taps T[kPlPx];
  bool valid[kPlPx];
  int wcl[kPlPx], hcl[kPlPx];
  int xmin = 1 << 30, xmax = -1, ymin = 1 << 30, ymax = -1;
#pragma unroll
  for (int k = 0; k < kPlPx; ++k) {
    const int w = tx * kPlTileW + (DVC_PLANAR_VERT ? 0 : k * 32) + lane;
    const int h = ty * kPlTileH + wid + (DVC_PLANAR_VERT ? k * 8 : 0);
    valid[k] = (w < W) && (h < H);
    wcl[k] = min(w, W - 1);          // out-of-tile lanes replay a real pixel: the box is unaffected
    hcl[k] = min(h, H - 1);
    const int hc = hcl[k];
    const float fx = fetch_flow(fl, t.fl_h, t.fl_w, hc, wcl[k], t.flow_level);
    const float fy = fetch_flow(fl + t.fl_c, t.fl_h, t.fl_w, hc, wcl[k], t.flow_level);
    T[k] = make_taps(t.g, hc, wcl[k], fx, fy);
    xmin = min(xmin, T[k].x0); xmax = max(xmax, T[k].x0);
    ymin = min(ymin, T[k].y0); ymax = max(ymax, T[k].y0);
  }

  // ---- bounding box of the tile's taps -----------------------------------------
  xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
  ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
  if (lane == 0) {
    s_red[wid] = xmin;
    s_red[8 + wid] = xmax;
    s_red[16 + wid] = ymin;
    s_red[24 + wid] = ymax;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    xmin = min(xmin, s_red[i]);
    xmax = max(xmax, s_red[8 + i]);
    ymin = min(ymin, s_red[16 + i]);
    ymax = max(ymax, s_red[24 + i]);
  }
  const PlBox box = planar_box(xmin, xmax, ymin, ymax, H, W, kStaged ? 0 : t.pl_tma);
  const int bx0 = box.bx0, by0 = box.by0, hh = box.hh, w4 = box.w4;
#if DVC_PLANAR_PITCH
  constexpr int pitch = DVC_PLANAR_PITCH;      // rows are a fixed pitch apart
#else
  // an odd float4 pitch spreads the rows of a box over the banks (pitch = 4 mod 8 words)
  const int pitch = ((w4 & 1) ? w4 : w4 + 1) * 4;
#endif

  const float* __restrict__ im_n = t.im + n * t.im_n;
  float* __restrict__ po = t.out + n * t.out_n;

  if (box.fits != kStaged) return;
  if (!kStaged) {
    // ---- incoherent flow: global gathers (warp_tile_strided's arithmetic), both
    // pixels of the thread interleaved so 8 independent loads are in flight per channel
    const float* p_nw[kPlPx];
    int o_e[kPlPx];
    long long o_s[kPlPx];
    float* pk[kPlPx];
#pragma unroll
    for (int k = 0; k < kPlPx; ++k) {
      p_nw[k] = im_n + T[k].y0 * t.im_h + T[k].x0;
      o_e[k] = T[k].dx ? 1 : 0;
      o_s[k] = T[k].dy ? t.im_h : 0;
      pk[k] = po + hcl[k] * t.out_h + wcl[k] * t.out_w;
    }
#pragma unroll 2
    for (int c = 0; c < t.C; ++c) {
#pragma unroll
      for (int k = 0; k < kPlPx; ++k) {
        const bool in_e = T[k].dx != 0, in_s = T[k].dy != 0;
        const float vnw = __ldg(p_nw[k]);
        float vne = __ldg(p_nw[k] + o_e[k]);
        float vsw = __ldg(p_nw[k] + o_s[k]);
        float vse = __ldg(p_nw[k] + o_s[k] + o_e[k]);
        vne = in_e ? vne : 0.f;
        vsw = in_s ? vsw : 0.f;
        vse = (in_e && in_s) ? vse : 0.f;
        if (valid[k]) *pk[k] = blend(vnw, vne, vsw, vse, T[k]);
        p_nw[k] += t.im_c;
        pk[k] += t.out_c;
      }
    }
    return;
  }

  // ---- staged path ------------------------------------------------------------------
  // this thread's share of a box copy: the same (row, float4 column) slots for every
  // plane; the source pointers advance by one plane per use
  PlanarLoop L;
  const int n_vec = w4 * hh;
#pragma unroll
  for (int k = 0; k < kPlSlots; ++k) {
    const int e = threadIdx.x + k * kThreads;
    const int r = e / w4, c4 = e - r * w4;
    L.src[k] = im_n + (long long)(by0 + r) * t.im_h + bx0 + c4 * 4;
    L.dst[k] = e < n_vec ? (uint32_t)(r * pitch + c4 * 4) * 4u : 0xffffffffu;
  }
  bool all_valid = true;
#pragma unroll
  for (int k = 0; k < kPlPx; ++k) {
    const bool in_e = T[k].dx != 0, in_s = T[k].dy != 0;
    L.i_nw[k] = (T[k].y0 - by0) * pitch + (T[k].x0 - bx0);
    L.i_ne[k] = L.i_nw[k] + (in_e ? 1 : 0);
    L.i_sw[k] = L.i_nw[k] + (in_s ? pitch : 0);
    L.i_se[k] = L.i_sw[k] + (in_e ? 1 : 0);
    L.pk[k] = po + hcl[k] * t.out_h + wcl[k] * t.out_w;
    L.T[k] = T[k];
    L.valid[k] = valid[k];
    all_valid = all_valid && valid[k];
  }
  // Block-uniform fast case: the whole tile is inside the image and no tap of it touches the
  // east / south border, so neither the out-of-image masks nor the store predicate is needed.
  const bool fast = (xmax + 1 <= W - 1) && (ymax + 1 <= H - 1) && (tx + 1) * kPlTileW <= W &&
                    (ty + 1) * kPlTileH <= H;
  if (fast)
    planar_planes<true>(L, pl_smem, t.C, t.im_c, t.out_c);
  else
    planar_planes<false>(L, pl_smem, t.C, t.im_c, t.out_c);
}

// ---------------------------------------------------------------------------
// planar path, boxes staged by TMA (cp.async.bulk.tensor.4d)
//
// The cp.async variant above spends a quarter of its instructions, and -- at
// 8 LSU cycles per LDGSTS warp-instruction -- most of the SM's load/store pipe on
// copying the box (ncu: L1/shared pipe 78 %, issue slots 67 %).  A tap box is a
// rectangle of one channel plane: exactly one TMA tile.  One elected thread
// issues ONE instruction per plane; the copy engine writes the 64 x BH box into
// shared memory with a dense 64-float pitch (bank = column mod 32: the gather of a
// coherent flow is conflict-free, see DVC_PLANAR_PITCH) and signals an mbarrier.
// Three tensor maps (box heights 24 / 32 / 40 rows) keep the over-fetch down; the
// box origin needs no alignment, out-of-image parts are zero-filled and never
// read (out-of-image taps are masked, as everywhere).
// ---------------------------------------------------------------------------
constexpr int kTmaBoxW = 64;
constexpr int kTmaH0 = 24, kTmaH1 = 32, kTmaH2 = kPlBoxH;     // box heights of the three maps
constexpr int kTmaBufBytes = kTmaBoxW * kPlBoxH * 4;          // 10 240
constexpr int kTmaSmemBytes = kPlBufs * kTmaBufBytes + 128;   // + alignment slack
struct alignas(64) PlanarMaps {
  CUtensorMap m[3];
};

__device__ __forceinline__ uint32_t pl_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void pl_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 20)) __trap();   // a lost copy traps instead of hanging the device
  }
}

__global__ void __launch_bounds__(kThreads, DVC_PLANAR_MINB)
warp_planar_tma_kernel(const __grid_constant__ WarpTask t, const __grid_constant__ PlanarMaps maps) {
  extern __shared__ __align__(128) unsigned char pl_tma_smem[];
  __shared__ int s_red[32];
  __shared__ __align__(8) unsigned long long s_full[kPlBufs];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int tile = blockIdx.x;
  const int tx = tile % t.tiles_x;
  const int rest = tile / t.tiles_x;
  const int ty = rest % t.tiles_y;
  const int n = rest / t.tiles_y;
  const int H = t.g.H, W = t.g.W;
  const float* __restrict__ fl = t.flow + n * t.fl_n;

  Taps T[kPlPx];
  bool valid[kPlPx];
  int wcl[kPlPx], hcl[kPlPx];
  int xmin = 1 << 30, xmax = -1, ymin = 1 << 30, ymax = -1;
#pragma unroll
  for (int k = 0; k < kPlPx; ++k) {
    const int w = tx * kPlTileW + (DVC_PLANAR_VERT ? 0 : k * 32) + lane;
    const int h = ty * kPlTileH + wid + (DVC_PLANAR_VERT ? k * 8 : 0);
    valid[k] = (w < W) && (h < H);
    wcl[k] = min(w, W - 1);
    hcl[k] = min(h, H - 1);
    const float fx = fetch_flow(fl, t.fl_h, t.fl_w, hcl[k], wcl[k], t.flow_level);
    const float fy = fetch_flow(fl + t.fl_c, t.fl_h, t.fl_w, hcl[k], wcl[k], t.flow_level);
    T[k] = make_taps(t.g, hcl[k], wcl[k], fx, fy);
    xmin = min(xmin, T[k].x0); xmax = max(xmax, T[k].x0);
    ymin = min(ymin, T[k].y0); ymax = max(ymax, T[k].y0);
  }
  xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
  ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
  if (lane == 0) {
    s_red[wid] = xmin;
    s_red[8 + wid] = xmax;
    s_red[16 + wid] = ymin;
    s_red[24 + wid] = ymax;
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int b = 0; b < kPlBufs; ++b)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pl_smem_u32(&s_full[b])), "r"(1)
                   : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    xmin = min(xmin, s_red[i]);
    xmax = max(xmax, s_red[8 + i]);
    ymin = min(ymin, s_red[16 + i]);
    ymax = max(ymax, s_red[24 + i]);
  }
  const PlBox box = planar_box(xmin, xmax, ymin, ymax, H, W, 1);
  if (!box.fits) return;           // block-uniform: the gather launch takes this tile
  const int bx0 = box.bx0, by0 = box.by0;
  const int sel = box.hh <= kTmaH0 ? 0 : (box.hh <= kTmaH1 ? 1 : 2);
  const uint32_t bytes = (uint32_t)(kTmaBoxW * 4) *
                         (uint32_t)(sel == 0 ? kTmaH0 : (sel == 1 ? kTmaH1 : kTmaH2));
  const CUtensorMap* map = &maps.m[sel];

  // 128-byte aligned staging buffers
  const uint32_t s_base = (pl_smem_u32(pl_tma_smem) + 127u) & ~127u;
  const float* __restrict__ sbuf0 = reinterpret_cast<const float*>(
      pl_tma_smem + (s_base - pl_smem_u32(pl_tma_smem)));
  const int C = t.C;
  auto issue = [&](int c) {        // thread 0 only
    if (c < C) {
      const int b = c % kPlBufs;
      const uint32_t bar = pl_smem_u32(&s_full[b]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                   : "memory");
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
          "[%0], [%1, {%2, %3, %4, %5}], [%6];"
          ::"r"(s_base + (uint32_t)(b * kTmaBufBytes)), "l"(map), "r"(bx0), "r"(by0), "r"(c), "r"(n),
            "r"(bar)
          : "memory");
    }
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int c = 0; c < kPlBufs - 1; ++c) issue(c);
  }

  int i_nw[kPlPx], i_ne[kPlPx], i_sw[kPlPx], i_se[kPlPx];
  float* pk[kPlPx];
  float* __restrict__ po = t.out + n * t.out_n;
#pragma unroll
  for (int k = 0; k < kPlPx; ++k) {
    const bool in_e = T[k].dx != 0, in_s = T[k].dy != 0;
    i_nw[k] = (T[k].y0 - by0) * kTmaBoxW + (T[k].x0 - bx0);
    i_ne[k] = i_nw[k] + (in_e ? 1 : 0);
    i_sw[k] = i_nw[k] + (in_s ? kTmaBoxW : 0);
    i_se[k] = i_sw[k] + (in_e ? 1 : 0);
    pk[k] = po + hcl[k] * t.out_h + wcl[k] * t.out_w;
  }
#pragma unroll 1
  for (int c = 0; c < C; ++c) {
    __syncthreads();               // everyone is done with plane c - 1: its buffer is free
    if (threadIdx.x == 0) issue(c + kPlBufs - 1);
    const int b = c % kPlBufs;
    pl_mbar_wait(pl_smem_u32(&s_full[b]), (uint32_t)((c / kPlBufs) & 1));
    const float* __restrict__ sb = sbuf0 + b * (kTmaBufBytes / 4);
#pragma unroll
    for (int k = 0; k < kPlPx; ++k) {
      const bool in_e = T[k].dx != 0, in_s = T[k].dy != 0;
      const float vnw = sb[i_nw[k]];
      float vne = sb[i_ne[k]], vsw = sb[i_sw[k]], vse = sb[i_se[k]];
      vne = in_e ? vne : 0.f;
      vsw = in_s ? vsw : 0.f;
      vse = (in_e && in_s) ? vse : 0.f;
      if (valid[k]) __stcs(pk[k], blend(vnw, vne, vsw, vse, T[k]));
      pk[k] += t.out_c;
    }
  }
}

// cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// OFF by default and UNVERIFIED: on this pool's B200 boxes every cp.async.bulk.tensor
// (UTMALDG) raises cudaErrorIllegalInstruction -- our own minimal program with descriptors in
// param or global memory, ranks 2-4, with and without swizzle, and Triton's stock
// tensor-descriptor kernel alike (tools/tma_probe.cu, tools/triton_tma_probe.py,
// profiles/r02_tma_probe.md) -- so this path could not be run.  DVC_WARP_PLANAR_TMA=1 enables it.
static int planar_tma_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DVC_WARP_PLANAR_TMA");
    v = e ? atoi(e) : 0;
  }
  return v;
}

// tensor maps over im[N][C][H][W] (element strides im_n, im_c, im_h, 1); false if not encodable
static bool build_planar_maps(PlanarMaps& maps, const WarpTask& t, const dvc_warp_task& in) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || in.W < kTmaBoxW || in.H < 8) return false;
  const cuuint64_t dims[4] = {(cuuint64_t)in.W, (cuuint64_t)in.H, (cuuint64_t)in.C, (cuuint64_t)in.N};
  const cuuint64_t strides[3] = {(cuuint64_t)t.im_h * 4, (cuuint64_t)t.im_c * 4,
                                 (cuuint64_t)t.im_n * 4};
  for (int i = 0; i < 3; ++i)
    if (strides[i] == 0 || (strides[i] & 15) || strides[i] >= (1ULL << 40)) return false;
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const int heights[3] = {kTmaH0, kTmaH1, kTmaH2};
  for (int i = 0; i < 3; ++i) {
    const cuuint32_t box[4] = {(cuuint32_t)kTmaBoxW, (cuuint32_t)heights[i], 1, 1};
    CUresult r = enc(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in.im), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
  }
  return true;
}

static int planar_min_c() {   // tuning knob (not API)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DVC_WARP_PLANAR_MINC");
    v = e ? atoi(e) : 8;
  }
  return v;
}
static bool planar_ok(const WarpTask& t, const dvc_warp_task& in) {
  // unit pixel stride, 16-byte aligned rows / planes / samples, rows of whole float4
  return t.mode == kModeStrided && in.C >= planar_min_c() && t.im_w == 1 && (in.W % 4) == 0 &&
         (t.im_h % 4) == 0 && (t.im_c % 4) == 0 && (t.im_n % 4) == 0 && t.im_h >= in.W &&
         aligned16(in.im);
}

static void planar_tiles(WarpTask& t, const dvc_warp_task& in) {
  t.tiles_x = (int)((in.W + kPlTileW - 1) / kPlTileW);
  t.tiles_y = (int)((in.H + kPlTileH - 1) / kPlTileH);
  t.n_blocks = (int)((long long)t.tiles_x * t.tiles_y * in.N);
}

// one task through the TMA-staged kernel (+ its complement); only with DVC_WARP_PLANAR_TMA=1
static int launch_planar_tma(WarpTask t, const PlanarMaps& maps, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(warp_planar_tma_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaSmemBytes);
    if (e != cudaSuccess)
      return fail(DVC_ERR_CUDA, "flow_warp: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  t.pl_tma = 1;
  warp_planar_tma_kernel<<<(unsigned)t.n_blocks, kThreads, kTmaSmemBytes, stream>>>(t, maps);
  int rc = check_launch("warp_planar_tma_kernel");
  if (rc) return rc;
  PlanarBatch pb;
  pb.n_tasks = 1;
  for (int i = 0; i < kMaxTasks; ++i) pb.t[i] = t;
  pb.first[0] = 0;
  for (int i = 1; i <= kMaxTasks; ++i) pb.first[i] = t.n_blocks;
  warp_planar_kernel<false><<<(unsigned)t.n_blocks, kThreads, 0, stream>>>(pb);
  return check_launch("warp_planar_kernel<gather>");
}

static int launch_planar_batch(PlanarBatch& pb, cudaStream_t stream) {
  // largest task first
  for (int i = 1; i < pb.n_tasks; ++i)
    for (int j = i; j > 0 && pb.t[j].n_blocks > pb.t[j - 1].n_blocks; --j) {
      const WarpTask tmp = pb.t[j];
      pb.t[j] = pb.t[j - 1];
      pb.t[j - 1] = tmp;
    }
  long long total = 0;
  for (int i = 0; i < kMaxTasks; ++i) {
    pb.first[i] = (int)total;
    if (i < pb.n_tasks) total += pb.t[i].n_blocks;
    else pb.t[i] = pb.t[0];
  }
  pb.first[kMaxTasks] = (int)total;
  DVC_REQUIRE(total < 2147483647LL, "flow_warp: too many tiles");
  if (kPlSmemBytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(warp_planar_kernel<true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kPlSmemBytes);
    if (e != cudaSuccess)
      return fail(DVC_ERR_CUDA, "flow_warp: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  warp_planar_kernel<true><<<(unsigned)total, kThreads, kPlSmemBytes, stream>>>(pb);
  int rc = check_launch("warp_planar_kernel<staged>");
  if (rc) return rc;
  warp_planar_kernel<false><<<(unsigned)total, kThreads, 0, stream>>>(pb);
  return check_launch("warp_planar_kernel<gather>");
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int build_task(WarpTask& t, const dvc_warp_task& in, int flags) {
  DVC_REQUIRE(in.im && in.flow && in.out, "flow_warp: null pointer");
  DVC_REQUIRE(in.flow_downscale >= 0 && in.flow_downscale <= 2,
              "flow_warp: flow_downscale must be 0, 1 or 2");
  t.flow_level = in.flow_downscale;
  {
    // tuning knob (not API): DVC_WARP_PREFETCH = 0 | 1 | 2
    static int pf = -1;
    if (pf < 0) {
      const char* e = getenv("DVC_WARP_PREFETCH");
      pf = e ? atoi(e) : 1;
    }
    t.prefetch = pf;
  }
  DVC_REQUIRE(in.N > 0 && in.C > 0 && in.H > 0 && in.W > 0, "flow_warp: empty tensor");
  DVC_REQUIRE(in.H < (1 << 24) && in.W < (1 << 24) && in.N < 65536 && in.C < (1 << 24),
              "flow_warp: extent too large");
  t.im = in.im;
  t.flow = in.flow;
  t.out = in.out;
  t.N = (int)in.N;
  t.C = (int)in.C;
  fill_geom(t.g, in.H, in.W, flags);
  const Strides4 si = make_strides(in.im_st), sf = make_strides(in.flow_st),
                 so = make_strides(in.out_st);
  t.im_n = si.n; t.im_c = si.c; t.im_h = si.h; t.im_w = si.w;
  t.fl_n = sf.n; t.fl_c = sf.c; t.fl_h = sf.h; t.fl_w = sf.w;
  t.out_n = so.n; t.out_c = so.c; t.out_h = so.h; t.out_w = so.w;
  // float4 path: channels_last, c4 = C/4 lanes per pixel must divide a warp
  const int64_t c4 = in.C / 4;
  const bool vec = nhwc_vec4_ok(in.im, si, in.C) && nhwc_vec4_ok(in.out, so, in.C) &&
                   c4 >= 1 && c4 <= 32 && (32 % c4) == 0 &&
                   fits_int32(1, in.C, in.H, in.W, si) && si.h >= 0 && si.w >= 0 &&
                   (long long)in.H * si.h + (long long)in.W * si.w < (1LL << 32);
  if (vec) {
    t.mode = (in.C == 64 && si.w == 64) ? kModeVec4C64 : kModeVec4;
    t.c4 = (int)c4;
    t.ppb = (kThreads / 32) * (32 / t.c4);   // pixel columns per CTA
    t.rows = t.c4;                           // rows per CTA (= rows per warp tile)
    t.tiles_x = (int)((in.W + t.ppb - 1) / t.ppb);
  } else {
    t.mode = kModeStrided;
    t.c4 = 0;
    t.ppb = kStridedTileW;
    t.rows = kRows;
    t.tiles_x = (int)((in.W + kStridedTileW - 1) / kStridedTileW);
  }
  t.tiles_y = (int)((in.H + t.rows - 1) / t.rows);
  const long long nb = (long long)t.tiles_x * t.tiles_y * in.N;
  DVC_REQUIRE(nb < 2147483647LL, "flow_warp: too many tiles");
  t.n_blocks = (int)nb;
  return DVC_OK;
}

static int launch_batch(const dvc_warp_task* tasks, int n_tasks, int flags,
                        cudaStream_t stream) {
  DVC_REQUIRE(tasks && n_tasks >= 1 && n_tasks <= kMaxTasks,
              "warp_multi: n_tasks must be in [1,%d]", kMaxTasks);
  WarpBatch batch;
  PlanarBatch planar_batch;
  planar_batch.n_tasks = 0;
  long long total = 0;
  int kept = 0;
  static int planar = -1;   // tuning knob (not API): DVC_WARP_PLANAR=0 keeps NCHW on the strided path
  if (planar < 0) {
    const char* e = getenv("DVC_WARP_PLANAR");
    planar = e ? atoi(e) : 1;
  }
  for (int i = 0; i < n_tasks; ++i) {
    int rc = build_task(batch.t[kept], tasks[i], flags);
    if (rc) return rc;
    if (planar && planar_ok(batch.t[kept], tasks[i])) {   // NCHW features: the staged kernels
      WarpTask& pt = planar_batch.t[planar_batch.n_tasks];
      pt = batch.t[kept];
      pt.pl_tma = 0;
      planar_tiles(pt, tasks[i]);
      DVC_REQUIRE((long long)pt.tiles_x * pt.tiles_y * tasks[i].N < 2147483647LL,
                  "flow_warp: too many tiles");
      PlanarMaps maps;
      if (planar_tma_mode() && build_planar_maps(maps, pt, tasks[i])) {
        rc = launch_planar_tma(pt, maps, stream);
        if (rc) return rc;
      } else {
        ++planar_batch.n_tasks;
      }
      continue;
    }
    batch.t[kept].first_block = (int)total;
    total += batch.t[kept].n_blocks;
    ++kept;
  }
  if (planar_batch.n_tasks > 0) {
    int rc = launch_planar_batch(planar_batch, stream);
    if (rc) return rc;
  }
  n_tasks = kept;
  batch.n_tasks = n_tasks;
  if (n_tasks == 0) return DVC_OK;
  for (int i = n_tasks; i < kMaxTasks; ++i) batch.t[i] = batch.t[0];
  DVC_REQUIRE(total < 2147483647LL, "warp_multi: grid too large");
  // proportional interleave: quota[k] ~ n_blocks[k] * 32 / total (>= 1)
  {
    const int kPeriodTarget = 32;
    int period = 0;
    long long cycles = 1LL << 40;
    for (int i = 0; i < kMaxTasks; ++i) {
      int q = 0;
      if (i < n_tasks) {
        q = (int)((batch.t[i].n_blocks * (long long)kPeriodTarget + total / 2) / total);
        if (q < 1) q = 1;
        const long long c = batch.t[i].n_blocks / q;
        if (c < cycles) cycles = c;
      }
      batch.quota[i] = q;
      batch.qprefix[i] = period;
      period += q;
    }
    if (n_tasks == 1) cycles = 0;   // nothing to interleave
    batch.period = period;
    batch.interleaved = (int)(cycles * period);
    long long pos = batch.interleaved;
    for (int i = 0; i < kMaxTasks; ++i) {
      batch.tail_first[i] = (int)pos;
      batch.tail_tile0[i] = (int)(cycles * batch.quota[i]);
      if (i < n_tasks) pos += batch.t[i].n_blocks - cycles * batch.quota[i];
    }
  }
  warp_multi_kernel<<<(unsigned)total, kThreads, 0, stream>>>(batch);
  return check_launch("warp_multi_kernel");
}

// ---------------------------------------------------------------------------
// bilinear 2x downscale (align_corners=False), forward
// ---------------------------------------------------------------------------
struct DownP {
  const float* x;
  float* y;
  int N, C, H, W, Ho, Wo;
  long long xs_n, xs_c, xs_h, xs_w, ys_n, ys_c, ys_h, ys_w;
  float scale_h, scale_w, post;
  int exact2;  // H, W even: source lambdas are exactly 1/2
};

// UpSample.cuh area_pixel_compute_source_index(align_corners=false, cubic=false)
__device__ __forceinline__ float area_src(float scale, int dst) {
  float s = fmaf(scale, (float)dst + 0.5f, -0.5f);
  return s < 0.f ? 0.f : s;
}

__global__ void __launch_bounds__(256) bilinear_down2_kernel(const DownP p) {
  const long long total = (long long)p.N * p.C * p.Ho * p.Wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int wo = (int)(i % p.Wo);
    long long r = i / p.Wo;
    const int ho = (int)(r % p.Ho);
    r /= p.Ho;
    const int c = (int)(r % p.C);
    const int n = (int)(r / p.C);
    const float* __restrict__ xp = p.x + n * p.xs_n + c * p.xs_c;
    float v;
    if (p.exact2) {
      const float* q = xp + (2 * ho) * p.xs_h + (2 * wo) * p.xs_w;
      const float a = __ldg(q), b = __ldg(q + p.xs_w);
      const float cc = __ldg(q + p.xs_h), d = __ldg(q + p.xs_h + p.xs_w);
      // 0.5*(0.5a+0.5b) + 0.5*(0.5c+0.5d): power-of-two scalings commute with
      // rounding, so this is exactly ((a+b)+(c+d))/4
      v = mul_rn(add_rn(add_rn(a, b), add_rn(cc, d)), 0.25f);
    } else {
      const float h1r = area_src(p.scale_h, ho), w1r = area_src(p.scale_w, wo);
      const int h1 = (int)h1r, w1 = (int)w1r;
      const int h1p = (h1 < p.H - 1) ? 1 : 0, w1p = (w1 < p.W - 1) ? 1 : 0;
      const float h1l = h1r - (float)h1, h0l = 1.f - h1l;
      const float w1l = w1r - (float)w1, w0l = 1.f - w1l;
      const float* q = xp + h1 * p.xs_h + w1 * p.xs_w;
      const float a = __ldg(q), b = __ldg(q + w1p * p.xs_w);
      const float cc = __ldg(q + h1p * p.xs_h), d = __ldg(q + h1p * p.xs_h + w1p * p.xs_w);
      // upsample_bilinear2d_out_frame as nvcc contracts it
      v = fmaf(h0l, fmaf(w0l, a, w1l * b), h1l * fmaf(w0l, cc, w1l * d));
    }
    p.y[n * p.ys_n + c * p.ys_c + ho * p.ys_h + wo * p.ys_w] = mul_rn(v, p.post);
  }
}

// two pyramid levels in one pass: one thread per level-3 sample reads a 4x4
// block of mv, emits 2x2 level-2 samples and 1 level-3 sample.
struct PyrP {
  const float* mv;
  float* mv2;
  float* mv3;
  int N, H3, W3;
  long long a_n, a_c, a_h, a_w, b_n, b_c, b_h, b_w, c_n, c_c, c_h, c_w;
};

__global__ void __launch_bounds__(256) flow_pyramid_kernel(const PyrP p) {
  const long long total = (long long)p.N * 2 * p.H3 * p.W3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int w3 = (int)(i % p.W3);
    long long r = i / p.W3;
    const int h3 = (int)(r % p.H3);
    r /= p.H3;
    const int c = (int)(r & 1);
    const int n = (int)(r >> 1);
    const float* __restrict__ src = p.mv + n * p.a_n + c * p.a_c + (4 * h3) * p.a_h + (4 * w3) * p.a_w;
    float m2[2][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float* q = src + (2 * j) * p.a_h + (2 * k) * p.a_w;
        const float a = __ldg(q), b = __ldg(q + p.a_w);
        const float cc = __ldg(q + p.a_h), d = __ldg(q + p.a_h + p.a_w);
        // bilinear /2 then "/ 2" (video_model.py:499): (.)*0.25*0.5
        m2[j][k] = mul_rn(mul_rn(add_rn(add_rn(a, b), add_rn(cc, d)), 0.25f), 0.5f);
        p.mv2[n * p.b_n + c * p.b_c + (2 * h3 + j) * p.b_h + (2 * w3 + k) * p.b_w] = m2[j][k];
      }
    const float v3 = mul_rn(
        mul_rn(add_rn(add_rn(m2[0][0], m2[0][1]), add_rn(m2[1][0], m2[1][1])), 0.25f), 0.5f);
    p.mv3[n * p.c_n + c * p.c_c + h3 * p.c_h + w3 * p.c_w] = v3;
  }
}

static unsigned grid_for(long long total, int threads, int per_sm) {
  long long want = (total + threads - 1) / threads;
  long long cap = (long long)sm_count() * per_sm;
  if (want < 1) want = 1;
  return (unsigned)(want < cap ? want : cap);
}

}  // namespace dvc

using namespace dvc;

extern "C" {

int dvc_flow_warp_fwd(const float* im, const float* flow, float* out, int64_t N, int64_t C,
                      int64_t H, int64_t W, const int64_t im_st[4], const int64_t flow_st[4],
                      const int64_t out_st[4], int flags, dvc_stream_t stream) {
  DVC_REQUIRE(im_st && flow_st && out_st, "flow_warp: null strides");
  dvc_warp_task t;
  t.im = im; t.flow = flow; t.out = out;
  t.N = N; t.C = C; t.H = H; t.W = W;
  t.flow_downscale = 0;
  for (int i = 0; i < 4; ++i) {
    t.im_st[i] = im_st[i];
    t.flow_st[i] = flow_st[i];
    t.out_st[i] = out_st[i];
  }
  return launch_batch(&t, 1, flags, (cudaStream_t)stream);
}

int dvc_warp_multi_fwd(const dvc_warp_task* tasks, int n_tasks, int flags,
                       dvc_stream_t stream) {
  return launch_batch(tasks, n_tasks, flags, (cudaStream_t)stream);
}

int dvc_bilinear_down2_fwd(const float* x, float* y, int64_t N, int64_t C, int64_t H,
                           int64_t W, const int64_t x_st[4], const int64_t y_st[4],
                           float post_scale, dvc_stream_t stream) {
  DVC_REQUIRE(x && y && x_st && y_st, "bilinear_down2: null pointer");
  DVC_REQUIRE(N > 0 && C > 0 && H >= 2 && W >= 2, "bilinear_down2: needs H,W >= 2");
  DownP p;
  p.x = x; p.y = y;
  p.N = (int)N; p.C = (int)C; p.H = (int)H; p.W = (int)W;
  p.Ho = (int)(H / 2); p.Wo = (int)(W / 2);
  p.xs_n = x_st[0]; p.xs_c = x_st[1]; p.xs_h = x_st[2]; p.xs_w = x_st[3];
  p.ys_n = y_st[0]; p.ys_c = y_st[1]; p.ys_h = y_st[2]; p.ys_w = y_st[3];
  // UpSample.h area_pixel_compute_scale(align_corners=false): in / out in fp32
  p.scale_h = (float)H / (float)p.Ho;
  p.scale_w = (float)W / (float)p.Wo;
  p.post = post_scale;
  p.exact2 = ((H % 2) == 0 && (W % 2) == 0) ? 1 : 0;
  const long long total = (long long)N * C * p.Ho * p.Wo;
  bilinear_down2_kernel<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("bilinear_down2_kernel");
}

int dvc_flow_pyramid_fwd(const float* mv, float* mv2, float* mv3, int64_t N, int64_t H,
                         int64_t W, const int64_t mv_st[4], const int64_t mv2_st[4],
                         const int64_t mv3_st[4], dvc_stream_t stream) {
  DVC_REQUIRE(mv && mv2 && mv3 && mv_st && mv2_st && mv3_st, "flow_pyramid: null pointer");
  DVC_REQUIRE(N > 0 && H >= 4 && W >= 4 && (H % 4) == 0 && (W % 4) == 0,
              "flow_pyramid: H and W must be positive multiples of 4 (got %lld x %lld)",
              (long long)H, (long long)W);
  PyrP p;
  p.mv = mv; p.mv2 = mv2; p.mv3 = mv3;
  p.N = (int)N; p.H3 = (int)(H / 4); p.W3 = (int)(W / 4);
  p.a_n = mv_st[0]; p.a_c = mv_st[1]; p.a_h = mv_st[2]; p.a_w = mv_st[3];
  p.b_n = mv2_st[0]; p.b_c = mv2_st[1]; p.b_h = mv2_st[2]; p.b_w = mv2_st[3];
  p.c_n = mv3_st[0]; p.c_c = mv3_st[1]; p.c_h = mv3_st[2]; p.c_w = mv3_st[3];
  const long long total = (long long)N * 2 * p.H3 * p.W3;
  flow_pyramid_kernel<<<grid_for(total, 256, 8), 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("flow_pyramid_kernel");
}

}  // extern "C"
