// dvc_coder.cu -- entropy-coder inputs and range-ANS bit streams on the GPU
// (sm_100a).
//
// Replaces, for the reference's real-bitstream path (dmc/test.py:187-188 ->
// DMC.encode_inter / decode_inter, dmc/models/video_model.py:586-614):
//   GaussianConditional.build_indexes          call sites video_model.py:248-249, 272, 282,
//                                              422-423, 447, 457
//   EntropyModel.quantize(.., "symbols")       inside compress
//   GaussianConditional.compress / decompress  call sites video_model.py:250-251, 273, 283,
//                                              424-425, 448, 458
//   EntropyBottleneck.compress / decompress    call sites video_model.py:238-239, 257, 411-412, 431
// all of which live in CompressAI (Python + a C++ rANS extension working on
// Python lists on the CPU: one device->host copy, two .tolist() and one
// sequential coder call per tensor).
//
// Bit-stream arithmetic = CompressAI's (ryg_rans rans64: 64-bit state, lower
// bound 2^31, 32-bit renormalisation words, 16-bit probabilities, 4-bit bypass
// nibbles for out-of-table symbols), restated in oracle/c/rans_ref.c.  A rANS
// stream is one serial dependency chain, so a single stock stream cannot use a
// GPU.  The B200 layout therefore cuts the symbols of a sample (NCHW order)
// into SUB-STREAMS of `stream_symbols` symbols; each sub-stream is a complete
// stock rans64 stream (the stock decoder decodes it given that slice of the
// indexes) and one WARP owns one sub-stream:
//   phase A (32 lanes, data parallel): symbol = round(x - mean), index from the
//            scale table (binary search = the reference's 63 compare-and-
//            subtract passes), CDF look-up, exact reciprocal of the frequency
//            (Alverson / ryg Rans64EncSymbolInit) -> records in shared memory;
//   phase B (serial chain): renormalise, q = mulhi(x, rcp) >> shift,
//            x += bias + q * (2^16 - freq): ~20 dependent instructions per
//            symbol instead of a 64-bit division.
// Container of one sample (little-endian u32 words):
//   [0] 'DVC1'  [1] n_symbols  [2] stream_symbols  [3] n_streams
//   [4 .. 4+n_streams)  words in each sub-stream      then the sub-streams.
// stream_symbols = 0 selects ONE raw stock stream without header, byte-
// identical to CompressAI's RansEncoder.encode_with_indexes (interop mode,
// serial: one warp per sample).
//
// That is the 'DVC1' layout (lanes = 1).  The default since round 2 is the
// lane-interleaved layout further down ('DVC3' / 'DVS3', lanes = 32): 32 stock
// rans64 coders per warp with a shared word stream, implied zeros for the
// near-deterministic table rows, the data-parallel half in its own kernel.
#include <stddef.h>

#include "dvc_common.cuh"

namespace dvc {

constexpr uint32_t kMagic = 0x31435644u;  // "DVC1"
constexpr int kChunk = 256;               // symbols staged per phase-A round
constexpr int kCoderWarps = 4;            // warps (= sub-streams) per CTA
constexpr int kMaxTable = 256;            // scale-table entries held in shared memory

struct CTS {  // element strides of an [N,C,H,W] operand (0 = broadcast)
  long long n, c, h, w;
};
static inline CTS cts(const int64_t s[4]) {
  CTS r;
  if (s) { r.n = s[0]; r.c = s[1]; r.h = s[2]; r.w = s[3]; }
  else { r.n = r.c = r.h = r.w = 0; }
  return r;
}

struct Tables {
  const int32_t* cdf;       // [n_cdf][cdf_stride]
  const int32_t* cdf_size;  // [n_cdf]
  const int32_t* offset;    // [n_cdf]
  int n_cdf, cdf_stride;
};

struct Source {             // where symbols and indexes come from
  const int32_t* symbols;   // [opt] contiguous [N][L]
  const float* x;           // [opt] strided; symbol = int(round(x - mean))
  const float* means;       // [opt] strided (broadcast strides allowed)
  const int32_t* indexes;   // [opt] contiguous [N][L]
  const float* scales;      // [opt] strided: index = build_indexes(scale)
  const float* scale_table; // device float[T], ascending
  int T;
  float scale_bound;
  CTS xs, ms, ss;
  int C, H, W, HW;
  long long L;              // symbols per sample = C*H*W
  // checkerboard pairing of the scale planes (decoder, video_model.py:268, :281):
  // positions whose (h + w) parity differs from cb_parity read `cb_alt`
  // elements further (the other channel half); cb_parity < 0: off
  int cb_parity;
  long long cb_alt;
};

// GaussianConditional.build_indexes for one scale: s = max(scale, bound);
// index = (T-1) - #{k < T-1 : s <= table[k]} = first k in [0, T-1) with
// s <= table[k], else T-1 (table ascending; NaN compares false -> T-1).
__device__ __forceinline__ int scale_index(float s, float bound, const float* tab, int T) {
  s = (s < bound) ? bound : s;
  int lo = 0, hi = T - 1;  // answer in [lo, hi]
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (s <= tab[mid]) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}

__device__ __forceinline__ void split_chw(const Source& s, long long e, int& c, int& h, int& w) {
  const unsigned ue = (unsigned)e;   // L < 2^31 (fill_source): 32-bit divisions
  c = (int)(ue / (unsigned)s.HW);
  const unsigned r = ue - (unsigned)c * (unsigned)s.HW;
  h = (int)(r / (unsigned)s.W);
  w = (int)(r - (unsigned)h * (unsigned)s.W);
}

__device__ __forceinline__ int fetch_symbol(const Source& s, int n, long long e, int c, int h,
                                            int w) {
  if (s.symbols) return __ldg(s.symbols + n * s.L + e);
  float v = __ldg(s.x + n * s.xs.n + c * s.xs.c + h * s.xs.h + w * s.xs.w);
  if (s.means) v = sub_rn(v, __ldg(s.means + n * s.ms.n + c * s.ms.c + h * s.ms.h + w * s.ms.w));
  return (int)rintf(v);  // torch.round(...).int()
}

__device__ __forceinline__ int fetch_index(const Source& s, int n, long long e, int c, int h,
                                           int w, const float* tab) {
  if (s.indexes) return __ldg(s.indexes + n * s.L + e);
  if (s.scales) {
    long long o = n * s.ss.n + c * s.ss.c + h * s.ss.h + w * s.ss.w;
    if (s.cb_parity >= 0 && ((h + w) & 1) != s.cb_parity) o += s.cb_alt;
    return scale_index(__ldg(s.scales + o), s.scale_bound, tab, s.T);
  }
  return c;  // EntropyBottleneck._build_indexes: the channel
}

// ---------------------------------------------------------------------------
// f1: symbols + indexes as tensors (GaussianConditional.build_indexes API)
// ---------------------------------------------------------------------------
struct SymIdxP {
  Source src;
  int32_t* out_symbols;
  int32_t* out_indexes;
  int N;
};

__global__ void __launch_bounds__(256) symbols_indexes_kernel(const SymIdxP p) {
  __shared__ float tab[kMaxTable];
  for (int i = threadIdx.x; i < p.src.T; i += blockDim.x) tab[i] = __ldg(p.src.scale_table + i);
  __syncthreads();
  const long long total = p.src.L * p.N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / p.src.L);
    const long long e = i - n * p.src.L;
    int c, h, w;
    split_chw(p.src, e, c, h, w);
    if (p.out_symbols) p.out_symbols[i] = fetch_symbol(p.src, n, e, c, h, w);
    if (p.out_indexes) p.out_indexes[i] = fetch_index(p.src, n, e, c, h, w, tab);
  }
}

// ---------------------------------------------------------------------------
// f2: encoder
// ---------------------------------------------------------------------------
struct IlvEncChunk;
struct IlvDecChunk;

struct EncRec {                 // per-warp shared staging, structure of arrays
  unsigned long long rcp[kChunk];
  uint32_t bias[kChunk];        // start (+ 2^16 - 1 when freq == 1)
  uint32_t fs[kChunk];          // freq | rcp_shift << 17 | escape << 31
  uint32_t raw[kChunk];         // escape payload
};

struct EncP {
  Source src;
  Tables tb;
  const uint8_t* skip;          // [opt] lane-interleaved layout: byte per table row, != 0: value 0 implied
  const IlvEncChunk* ilv_enc;   // lane-interleaved layout: coder records of every chunk (prepare kernel)
  int chunks_per_sample;
  uint32_t* stream_words;       // scratch: [N][n_streams]
  uint32_t* stream_data;        // scratch: [N][n_streams][cap]
  long long S;                  // symbols per sub-stream
  int n_streams, cap, N;
  int* status;                  // [opt] device flag: bad index seen
};

__device__ __forceinline__ void enc_put_bits(unsigned long long& x, uint32_t*& ptr, uint32_t val) {
  // Rans64EncPutBits(nbits = 4): freq = 2^12, x_max = 2^47 * 2^12
  if (x >= (1ull << 59)) {
    *--ptr = (uint32_t)x;
    x >>= 32;
  }
  x = (x << 4) | val;
}

__global__ void __launch_bounds__(kCoderWarps * 32) rans_encode_kernel(const EncP p) {
  __shared__ EncRec recs[kCoderWarps];
  __shared__ float tab[kMaxTable];
  for (int i = threadIdx.x; i < p.src.T; i += blockDim.x) tab[i] = __ldg(p.src.scale_table + i);
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int j = blockIdx.x * kCoderWarps + wid;  // sub-stream
  const int n = blockIdx.y;
  if (j >= p.n_streams) return;
  EncRec& R = recs[wid];
  const long long s_begin = (long long)j * p.S;
  const long long s_end = min(p.src.L, s_begin + p.S);
  uint32_t* const cap_end = p.stream_data + ((long long)n * p.n_streams + j + 1) * p.cap;
  uint32_t* ptr = cap_end;
  unsigned long long x = 1ull << 31;  // Rans64EncInit
  bool bad = false;

  for (long long ce = s_end; ce > s_begin; ce -= kChunk) {
    const long long cb = max(s_begin, ce - kChunk);
    const int cnt = (int)(ce - cb);
    // ---- phase A: records of this chunk, data parallel -----------------------
    for (int i = lane; i < cnt; i += 32) {
      const long long e = cb + i;
      int c, h, w;
      split_chw(p.src, e, c, h, w);
      const int sym = fetch_symbol(p.src, n, e, c, h, w);
      int ci = fetch_index(p.src, n, e, c, h, w, tab);
      if (ci < 0 || ci >= p.tb.n_cdf) { bad = true; ci = 0; }
      const int32_t* __restrict__ row = p.tb.cdf + (long long)ci * p.tb.cdf_stride;
      const int max_value = __ldg(p.tb.cdf_size + ci) - 2;
      int value = sym - __ldg(p.tb.offset + ci);
      uint32_t raw = 0, esc = 0;
      if (value < 0) {
        raw = (uint32_t)(-2 * value - 1);
        value = max_value;
      } else if (value >= max_value) {
        raw = (uint32_t)(2 * (value - max_value));
        value = max_value;
      }
      if (value == max_value) esc = 1u;
      const uint32_t start = (uint32_t)__ldg(row + value);
      const uint32_t freq = (uint32_t)__ldg(row + value + 1) - start;
      // Rans64EncSymbolInit: q = floor(x / freq) = mulhi(x, rcp) >> shift, exact
      unsigned long long rcp;
      uint32_t shift, bias;
      if (freq < 2) {
        rcp = ~0ull;
        shift = 0;
        bias = start + (1u << 16) - 1u;
      } else {
        uint32_t sh = 32u - (uint32_t)__clz((int)(freq - 1u));  // ceil(log2(freq))
        const unsigned long long x1 = 1ull << (sh + 31);
        const unsigned long long t1 = x1 / freq;
        const unsigned long long x0 = (unsigned long long)(freq - 1u) + ((x1 % freq) << 32);
        rcp = x0 / freq + (t1 << 32);
        shift = sh - 1u;
        bias = start;
      }
      R.rcp[i] = rcp;
      R.bias[i] = bias;
      R.fs[i] = freq | (shift << 17) | (esc << 31);
      R.raw[i] = raw;
    }
    __syncwarp();
    // ---- phase B: the serial chain, last symbol first -------------------------
    // One lane; the record of symbol i-1 is fetched while symbol i is in the
    // chain, so only the state update itself is serial: renormalisation test
    // x >= freq << 47  <=>  (x >> 47) >= freq  (one shift of the high word),
    // q = mulhi(x, rcp) >> shift, x += bias + q * (2^16 - freq).
    if (lane == 0) {
      int i = cnt - 1;
      uint32_t fs = R.fs[i], bias = R.bias[i], raw = R.raw[i];
      unsigned long long rcp = R.rcp[i];
      while (true) {
        const int in = i > 0 ? i - 1 : 0;
        const uint32_t fs_n = R.fs[in], bias_n = R.bias[in], raw_n = R.raw[in];
        const unsigned long long rcp_n = R.rcp[in];
        if (fs >> 31) {  // bypass: raw nibbles (high first), then their count
          int nb = 0;
          while (nb < 8 && (raw >> (nb * 4)) != 0) ++nb;
          for (int k = nb - 1; k >= 0; --k) enc_put_bits(x, ptr, (raw >> (k * 4)) & 15u);
          enc_put_bits(x, ptr, (uint32_t)nb);  // nb <= 8 < 15: a single count nibble
        }
        const uint32_t freq = fs & 0x1ffffu;
        const uint32_t shift = (fs >> 17) & 31u;
        if ((uint32_t)(x >> 47) >= freq) {  // Rans64EncPut renormalisation
          *--ptr = (uint32_t)x;
          x >>= 32;
        }
        const unsigned long long q = __umul64hi(x, rcp) >> shift;
        x = x + bias + q * (unsigned long long)((1u << 16) - freq);
        if (i == 0) break;
        --i;
        fs = fs_n; bias = bias_n; raw = raw_n; rcp = rcp_n;
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    ptr -= 2;  // Rans64EncFlush
    ptr[0] = (uint32_t)x;
    ptr[1] = (uint32_t)(x >> 32);
    p.stream_words[(long long)n * p.n_streams + j] = (uint32_t)(cap_end - ptr);
  }
  if (p.status && __any_sync(0xffffffffu, bad) && lane == 0) atomicExch(p.status, 1);
}

// gather the sub-streams of a sample into its container
struct PackP {
  const uint32_t* stream_words;
  const uint32_t* stream_data;
  uint8_t* out;
  long long out_stride;   // bytes between samples
  long long* out_bytes;   // [N]: container size, or -(needed) if out_stride is too small
  long long L, S;
  int n_streams, cap, header;  // header = 1: container, 0: raw stock stream
  uint32_t magic;
  // adaptive implied zeros: magic_alt is written instead when *flagged > flagged_max
  uint32_t magic_alt;
  const unsigned* flagged;
  long long flagged_max;
};

__global__ void __launch_bounds__(128) rans_pack_kernel(const PackP p) {
  __shared__ unsigned long long red[4];
  const int j = blockIdx.x, n = blockIdx.y;
  const uint32_t* cnt = p.stream_words + (long long)n * p.n_streams;
  unsigned long long before = 0, total = 0;
  for (int i = threadIdx.x; i < p.n_streams; i += blockDim.x) {
    const uint32_t cw = cnt[i];
    total += cw;
    if (i < j) before += cw;
  }
  // two block sums (before, total)
  for (int pass = 0; pass < 2; ++pass) {
    unsigned long long v = pass ? total : before;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    v = red[0] + red[1] + red[2] + red[3];
    __syncthreads();
    if (pass) total = v; else before = v;
  }
  const long long head = p.header ? (4 + p.n_streams) : 0;
  const long long need = (head + (long long)total) * 4;
  const bool fits = need <= p.out_stride;
  if (j == 0 && blockIdx.z == 0 && threadIdx.x == 0) p.out_bytes[n] = fits ? need : -need;
  if (!fits) return;
  uint32_t* dst = reinterpret_cast<uint32_t*>(p.out + n * p.out_stride);
  const uint32_t mine = cnt[j];
  if (p.header && blockIdx.z == 0) {
    if (j == 0 && threadIdx.x < 4) {
      const uint32_t magic = (p.flagged && (long long)*p.flagged > p.flagged_max) ? p.magic_alt : p.magic;
      const uint32_t hdr[4] = {magic, (uint32_t)p.L, (uint32_t)p.S, (uint32_t)p.n_streams};
      dst[threadIdx.x] = hdr[threadIdx.x];
    }
    if (threadIdx.x == 0) dst[4 + j] = mine;
  }
  const uint32_t* src = p.stream_data + ((long long)n * p.n_streams + j + 1) * p.cap - mine;
  uint32_t* d = dst + head + before;
  // gridDim.z CTAs share the copy of a long sub-stream
  for (uint32_t i = blockIdx.z * blockDim.x + threadIdx.x; i < mine; i += gridDim.z * blockDim.x)
    d[i] = src[i];
}

// ---------------------------------------------------------------------------
// f2: decoder.  One warp per sub-stream; every lane carries the same state so
// the CDF search is a 32-wide ballot (first probe centred on the table's mode,
// then 32-ary refinement).
// ---------------------------------------------------------------------------
struct DecStage {
  int idx[kChunk];
  int size[kChunk];
  int off[kChunk];
  int val[kChunk];
};

struct DecP {
  Source src;                 // indexes / scales / means (symbols, x unused)
  Tables tb;
  const uint8_t* skip;        // [opt] lane-interleaved layout ('DVS3'): byte per table row
  const uint16_t* pack;       // [opt] lane-interleaved layout: packed look-up + CDF rows
  int pack_entries;
  const IlvDecChunk* ilv_dec; // lane-interleaved layout: pass-1 items of every chunk (prepare kernel)
  const uint16_t* ilv_ci;     //   table row of every position
  int32_t* ilv_sym;           //   decoded symbols [N][L]
  int chunks_per_sample;
  const uint8_t* in;
  long long in_stride;        // bytes between samples
  const long long* in_bytes;  // [N] device
  float* out_f;               // [opt] strided: float(symbol) + mean
  int32_t* out_sym;           // [opt] contiguous [N][L]
  CTS os;
  long long S;
  int n_streams, header, N;
  int* status;                // [opt] device flag: 2 = malformed container
};

struct Reader {
  const uint32_t* ptr;
  const uint32_t* end;
  __device__ __forceinline__ uint32_t next() {
    const uint32_t v = (ptr < end) ? __ldg(ptr) : 0u;  // a corrupt stream reads zeros, never out of bounds
    ++ptr;
    return v;
  }
};

__device__ __forceinline__ uint32_t dec_get_bits(unsigned long long& x, Reader& rd) {
  const uint32_t val = (uint32_t)(x & 15u);  // Rans64DecGetBits(4)
  x >>= 4;
  if (x < (1ull << 31)) x = (x << 32) | rd.next();
  return val;
}

__global__ void __launch_bounds__(kCoderWarps * 32) rans_decode_kernel(const DecP p) {
  __shared__ DecStage stage[kCoderWarps];
  __shared__ float tab[kMaxTable];
  for (int i = threadIdx.x; i < p.src.T; i += blockDim.x) tab[i] = __ldg(p.src.scale_table + i);
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int j = blockIdx.x * kCoderWarps + wid;
  const int n = blockIdx.y;
  if (j >= p.n_streams) return;
  DecStage& G = stage[wid];
  const uint32_t* words = reinterpret_cast<const uint32_t*>(p.in + n * p.in_stride);
  const long long total_words = __ldg(p.in_bytes + n) >> 2;
  Reader rd;
  if (p.header) {
    bool ok = total_words >= 4 + p.n_streams && __ldg(words) == kMagic &&
              __ldg(words + 1) == (uint32_t)p.src.L && __ldg(words + 2) == (uint32_t)p.S &&
              __ldg(words + 3) == (uint32_t)p.n_streams;
    unsigned long long before = 0;
    if (ok)
      for (int i = lane; i < j; i += 32) before += __ldg(words + 4 + i);
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    const unsigned long long mine = ok ? __ldg(words + 4 + j) : 0ull;
    const long long first = 4 + p.n_streams + (long long)before;
    if (!ok || first + (long long)mine > total_words) {
      if (p.status && lane == 0) atomicExch(p.status, 2);
      rd.ptr = rd.end = words;  // decode zeros: defined, flagged
    } else {
      rd.ptr = words + first;
      rd.end = rd.ptr + mine;
    }
  } else {
    rd.ptr = words;
    rd.end = words + total_words;
  }
  unsigned long long x = (unsigned long long)rd.next();  // Rans64DecInit
  x |= (unsigned long long)rd.next() << 32;

  const long long s_begin = (long long)j * p.S;
  const long long s_end = min(p.src.L, s_begin + p.S);
  bool bad = false;
  for (long long cb = s_begin; cb < s_end; cb += kChunk) {
    const int cnt = (int)min((long long)kChunk, s_end - cb);
    // ---- phase A: table rows of this chunk ------------------------------------
    for (int i = lane; i < cnt; i += 32) {
      const long long e = cb + i;
      int c, h, w;
      split_chw(p.src, e, c, h, w);
      int ci = fetch_index(p.src, n, e, c, h, w, tab);
      if (ci < 0 || ci >= p.tb.n_cdf) { bad = true; ci = 0; }
      G.idx[i] = ci;
      G.size[i] = __ldg(p.tb.cdf_size + ci);
      G.off[i] = __ldg(p.tb.offset + ci);
    }
    __syncwarp();
    // ---- phase B: the chain; all lanes in lock step ---------------------------
    for (int i = 0; i < cnt; ++i) {
      const int size = G.size[i], off = G.off[i];
      const int32_t* __restrict__ row = p.tb.cdf + (long long)G.idx[i] * p.tb.cdf_stride;
      const uint32_t cum = (uint32_t)(x & 0xffffu);  // Rans64DecGet
      // s = max j in [0, size-1) with row[j] <= cum  (row[0] = 0, row[size-1] = 2^16)
      int lo, hi;
      uint32_t start = 0, next = 0;
      bool found = false;
      {
        int j0 = -off - 15;  // window centred on the mode (symbol 0 sits at -offset)
        j0 = max(0, min(j0, size - 32));
        const int jj = j0 + lane;
        const uint32_t v = (jj < size) ? (uint32_t)__ldg(row + jj) : 0xffffffffu;
        const unsigned b = __ballot_sync(0xffffffffu, v <= cum);
        if (b == 0u) { lo = 0; hi = j0; }
        else if (b == 0xffffffffu) { lo = j0 + 31; hi = size - 1; start = __shfl_sync(0xffffffffu, v, 31); }
        else {
          const int k = __popc(b);  // lanes [0,k) hold entries <= cum
          lo = j0 + k - 1;
          hi = lo + 1;
          start = __shfl_sync(0xffffffffu, v, k - 1);
          next = __shfl_sync(0xffffffffu, v, k);
          found = true;
        }
      }
      while (!found) {  // invariant: row[lo] <= cum < row[hi]
        const int span = hi - lo;
        if (span <= 1) {
          start = (uint32_t)__ldg(row + lo);
          next = (uint32_t)__ldg(row + lo + 1);
          break;
        }
        const int step = (span + 31) >> 5;
        const int jj = lo + lane * step;
        const uint32_t v = (jj < hi) ? (uint32_t)__ldg(row + jj) : 0xffffffffu;
        const unsigned b = __ballot_sync(0xffffffffu, v <= cum);
        const int k = __popc(b);  // >= 1: lane 0 probes row[lo]
        lo = lo + (k - 1) * step;
        hi = min(hi, lo + step);
      }
      const int s = lo;
      {  // Rans64DecAdvance
        const uint32_t freq = next - start;
        x = (unsigned long long)freq * (x >> 16) + (x & 0xffffu) - start;
        if (x < (1ull << 31)) x = (x << 32) | rd.next();
      }
      int value = s;
      if (s == size - 2) {  // bypass
        int val = (int)dec_get_bits(x, rd);
        int nb = val;
        while (val == 15 && rd.ptr <= rd.end) {
          val = (int)dec_get_bits(x, rd);
          nb += val;
        }
        uint32_t raw = 0;
        for (int k = 0; k < nb; ++k) {
          val = (int)dec_get_bits(x, rd);
          if (k < 8) raw |= (uint32_t)val << (k * 4);
        }
        value = (int)(raw >> 1);
        if (raw & 1u) value = -value - 1;
        else value += size - 2;
      }
      if (lane == 0) G.val[i] = value + off;
    }
    __syncwarp();
    // ---- write back, coalesced --------------------------------------------------
    for (int i = lane; i < cnt; i += 32) {
      const long long e = cb + i;
      const int v = G.val[i];
      if (p.out_sym) p.out_sym[n * p.src.L + e] = v;
      if (p.out_f) {
        int c, h, w;
        split_chw(p.src, e, c, h, w);
        float f = (float)v;  // EntropyModel.dequantize
        if (p.src.means)
          f = add_rn(f, __ldg(p.src.means + n * p.src.ms.n + c * p.src.ms.c + h * p.src.ms.h +
                              w * p.src.ms.w));
        p.out_f[n * p.os.n + c * p.os.c + h * p.os.h + w * p.os.w] = f;
      }
    }
    __syncwarp();
  }
  if (p.status && __any_sync(0xffffffffu, bad) && lane == 0) atomicExch(p.status, 1);
}

// ---------------------------------------------------------------------------
// f2, lane-interleaved container ('DVC3' / 'DVS3').
//
// A GPU lane walks a range-coder chain ~10x slower than a CPU core, and every
// independent chain costs its flush bytes, so the 'DVC1' layout (one chain per
// warp, 12 bytes per chain) has to choose between bytes and time.  Here a
// sub-stream is coded by the 32 lanes of ONE warp, each lane a stock rans64
// state (same per-symbol arithmetic as above, bypass coding included); item k
// of the sub-stream's item list belongs to lane k % 32 and is coded in round
// k / 32.  The lanes share one word stream: in every step the lanes that
// renormalise take consecutive words in lane order (ballot + popc), so the
// words of a round are contiguous and the decoder reads them with the same
// ballot.  A step is the t-th rans operation of a lane's item (t = 0: the
// symbol, t = 1: bypass count nibble, t >= 2: bypass data nibbles).  Flush: a
// mask word (bit l: lane l's state needs a high word), then the 32 states,
// 1 or 2 words each -- 132..260 bytes per sub-stream for 32 chains, i.e.
// 4..8 bytes per chain instead of 12.
//
// 'DVS3' adds implied zeros.  At low rates most latents of a DMC-style codec sit
// in the narrowest table rows (sigma at the 0.11 floor), where value 0 holds
// >= 65528/65536 of the mass: coding one costs ~1e-4 bit but a full chain step.
// `skip_rows` marks those rows (one byte per row, derived by both sides from the
// tables).  Positions are handled in chunks of 1024 = 32 groups of 32; pass 1 of
// a chunk codes, group by group, the regular symbols and -- for a group that has
// marked positions -- ONE flag "some marked symbol of this group is not 0"
// (P(flag) = 8 k / 65536 for k marked positions, an upper bound of what the
// tables say); pass 2 codes the marked symbols of the flagged groups with their
// ordinary table rows.  The chain is up to 32x shorter on the marked symbols, the
// bits are those of the stock coder (a flag costs what its zeros would have cost)
// plus ~8 bits per flagged group, and nothing is lossy: y_hat is bit-identical.
// ---------------------------------------------------------------------------
constexpr uint32_t kMagic3 = 0x33435644u;    // "DVC3"
constexpr uint32_t kMagic3S = 0x33535644u;   // "DVS3": + implied-zero groups
constexpr int kIlvChunk = 1024;              // positions per chunk (32 groups x 32 lanes)
constexpr int kIlvItems = kIlvChunk + 32;    // + one flag per group
constexpr int kRowCache = 256;               // table rows whose size/offset/mark live in shared memory
// Inverse-CDF look-up of the decoder: `cum` (16 bits) -> a key -> the first table position a
// symbol with such a `cum` can have (u16 per key and row).  The 416 keys are monotone in `cum`:
// 88 logarithmic keys for the lower end (less than 2048 counts from 0: eight per octave of the
// distance d from the end -- that is where a row has many symbols per count), 240 central keys
// of 256 counts (cum >> 8), 88 logarithmic keys for the upper end.  The logarithmic key is the
// exponent and the three leading mantissa bits of float(d | 1): one conversion, one shift.
// Because the keys are monotone, entry k + 1 bounds the bracket of entry k from above (the
// bisection of the rare case uses it; entry 416 is the sentinel size - 2).
// With 1024-count central keys and four keys per octave (the first version) the wide rows of a
// 0.11 ... 256 scale table had more than four positions per key for 4.5 % (sigma = 40) ... 75 %
// (sigma = 256) of the cum values, and the decoder fell into its bisection in most rounds of such
// content (profiles/r02_coder.md).
constexpr int kLutEnd = 88;
constexpr int kLutKeys = 2 * kLutEnd + 240;    // 416
constexpr int kLutStride = kLutKeys + 2;       // + sentinel, + pad (4-byte aligned rows)
__host__ __device__ inline int lut_key(uint32_t cum) {
  const uint32_t up = cum >> 15;                         // 1: upper half
  const uint32_t d = up ? 65535u - cum : cum;            // distance from the nearer end
#ifdef __CUDA_ARCH__
  const uint32_t t = (__float_as_uint(__uint2float_rn(d | 1u)) >> 20) - (127u << 3);
#else
  uint32_t e = 0;
  while (((d | 1u) >> (e + 1)) != 0u) ++e;
  const uint32_t t = (e << 3) + ((((d | 1u) << 3) >> e) & 7u);
#endif
  const uint32_t end = up ? (uint32_t)(kLutKeys - 1) - t : t;
  return (int)(d >= 2048u ? (uint32_t)(kLutEnd - 8) + (cum >> 8) : end);
}
constexpr int kPackMaxBytes = 124 * 1024;    // largest packed table the decoder stages in shared memory
constexpr int kIlvRing = 5;                  // chunks in flight between the copy engine and the chain
constexpr unsigned kFull = 0xffffffffu;
constexpr uint32_t kIdleItem = 0xffffu;       // position field of an idle item (decoder)

// The decoder's packed tables (`cdf_pack`, built by the caller from the CDF tables, see
// include/dvc_b200.h): u16 lut[n][418], u32 row_start[n], u16 cdf[total] holding (value - 1) mod
// 2^16 -- "cum >= value" is "cum > stored", 65536 fits, and value = (stored + 1) [mod 2^16 for the
// row's leading 0] -- each row followed by 4 entries 0xffff, so that
// the four probes of the decoder never leave it.
struct PackLayout {
  int start_off, tbl_off, bytes;   // byte offsets of row_start / cdf, total size (multiple of 16)
};
__host__ __device__ inline PackLayout pack_layout(int n_cdf, int total) {
  PackLayout q;
  q.start_off = n_cdf * kLutStride * 2;   // kLutStride is even: 4-byte aligned
  q.tbl_off = q.start_off + 4 * n_cdf;
  q.bytes = ((q.tbl_off + 2 * total + 15) / 16) * 16;
  return q;
}

// ---- PTX wrappers: mbarrier + 1-D bulk copy (the copy engine fills the chain's stages) -------
__device__ __forceinline__ uint32_t ilv_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void ilv_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ilv_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void ilv_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
    if (!ok && ++spins > (1u << 20)) __trap();   // a lost copy traps instead of hanging the GPU
  }
}
__device__ __forceinline__ void ilv_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes,
                                             uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

struct IlvRows {                   // per CTA, loaded once
  int2 so[kRowCache];              // (cdf_size, offset) of a table row
  uint8_t mark[kRowCache];
  float tab[kMaxTable];
};
__device__ __forceinline__ void ilv_load_rows(IlvRows& R, const Source& src, const Tables& tb,
                                              const uint8_t* skip) {
  for (int i = threadIdx.x; i < src.T; i += blockDim.x) R.tab[i] = __ldg(src.scale_table + i);
  for (int i = threadIdx.x; i < min(tb.n_cdf, kRowCache); i += blockDim.x) {
    R.so[i] = make_int2(__ldg(tb.cdf_size + i), __ldg(tb.offset + i));
    R.mark[i] = skip ? __ldg(skip + i) : (uint8_t)0;
  }
}
__device__ __forceinline__ int2 ilv_size_off(const IlvRows& R, const Tables& tb, int ci) {
  return ci < kRowCache ? R.so[ci] : make_int2(__ldg(tb.cdf_size + ci), __ldg(tb.offset + ci));
}
__device__ __forceinline__ bool ilv_marked(const IlvRows& R, const uint8_t* skip, int ci) {
  if (!skip) return false;
  return (ci < kRowCache ? R.mark[ci] : __ldg(skip + ci)) != 0;
}
__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total) {
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  total = __shfl_sync(kFull, inc, 31);
  return inc - v;
}

// What the data-parallel half (the prepare kernels, one CTA of 1024 threads per chunk, the
// whole GPU) leaves in global memory for the serial half (the chain kernels, one warp per
// sub-stream).  Everything is laid out per chunk; sub-streams are whole chunks.
struct IlvMeta {                   // 16 bytes per chunk
  int n1, n2;                      // items of pass 1 / pass 2 (decoder: n2 is found while decoding)
  int flags;                       // bit 0: some item is bypass-coded (encoder), bit 1: marked positions exist
  int pad;
};
struct IlvEncChunk {               // coder records in ITEM order: pass 1 at [0, n1), pass 2 at [n1, n1 + n2)
  unsigned long long rcp[kIlvItems];   // Rans64EncSymbolInit's exact reciprocal
  uint32_t bias[kIlvItems];        // start (+ 2^16 - 1 when freq == 1)
  uint32_t fs[kIlvItems];          // freq | rcp_shift << 17 | operations << 22 (1, or 2 + nibbles)
  uint32_t raw[kIlvItems];         // bypass payload
  IlvMeta meta;
};
struct IlvDecChunk {               // pass-1 items in order
  uint4 item[kIlvItems];           // x: row offset in the CDF table (flag: 65536 - 8k), y: cdf_size,
                                   // z: offset, w: position (flag: kIlvChunk + g) | table row << 16
  uint32_t skm[32];                // marked positions of each group
  IlvMeta meta;
  uint32_t pad[4];
};
static_assert(sizeof(IlvEncChunk) % 16 == 0 && sizeof(IlvDecChunk) % 16 == 0, "bulk-copy granularity");

struct IlvPrepP {
  Source src;
  Tables tb;
  const uint8_t* skip;
  IlvEncChunk* enc;                // encoder: [N * chunks_per_sample]
  IlvDecChunk* dec;                // decoder
  uint16_t* pos_ci;                // decoder: table row of every position [N][chunks_per_sample * 1024]
  int32_t* sym_out;                // decoder: [N][L]; implied zeros are written here
  const uint32_t* row_start;       // decoder [opt]: row offsets of the packed table
  // encoder, adaptive implied zeros: groups that would be flagged (counted by ilv_count_kernel);
  // above flagged_max the marks are ignored and the container is a plain 'DVC3'
  unsigned* flagged;
  long long flagged_max;
  int chunks_per_sample;
  int* status;
};

// ---- adaptive implied zeros: how many groups would carry a set flag ------------------------------
// A set flag costs ~8-13 bits on top of coding the group's symbols; on data the tables describe
// that is a fraction of a per cent of a per cent, on data they do not (marked rows full of
// non-zero symbols) it would inflate the stream, so the encoder counts first and only uses the
// marks when the flags stay within the caller's budget.
__global__ void __launch_bounds__(kIlvChunk) ilv_count_kernel(const IlvPrepP p) {
  __shared__ IlvRows R;
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x / p.chunks_per_sample;
  const int c = blockIdx.x - n * p.chunks_per_sample;
  ilv_load_rows(R, p.src, p.tb, p.skip);
  __syncthreads();
  const long long e = (long long)c * kIlvChunk + threadIdx.x;
  bool nz = false;
  if (e < p.src.L) {
    int ch, h, w;
    split_chw(p.src, e, ch, h, w);
    const int ci = fetch_index(p.src, n, e, ch, h, w, R.tab);
    if (ci >= 0 && ci < p.tb.n_cdf && ilv_marked(R, p.skip, ci))
      nz = fetch_symbol(p.src, n, e, ch, h, w) != 0;
  }
  const int any = __syncthreads_count(__any_sync(kFull, nz) && lane == 0);   // flagged groups of the chunk
  if (threadIdx.x == 0 && any) atomicAdd(p.flagged, (unsigned)any);
}

// ---- prepare: one chunk per CTA, warp g = group g, lane = position in the group ---------------
template <bool kEncoder>
__global__ void __launch_bounds__(kIlvChunk) ilv_prepare_kernel(const IlvPrepP p) {
  __shared__ IlvRows R;
  __shared__ int cnt1[32], cnt2[32];
  __shared__ int any_esc, any_mark;
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int n = blockIdx.x / p.chunks_per_sample;
  const int c = blockIdx.x - n * p.chunks_per_sample;
  if (threadIdx.x == 0) { any_esc = 0; any_mark = 0; }
  const bool use_marks = !kEncoder || !p.flagged || (long long)*p.flagged <= p.flagged_max;
  ilv_load_rows(R, p.src, p.tb, use_marks ? p.skip : nullptr);
  __syncthreads();
  const long long e = (long long)c * kIlvChunk + threadIdx.x;
  const bool valid = e < p.src.L;
  bool sk = false, bad = false;
  int sym = 0, ci = 0, ch = 0, h = 0, w = 0;
  if (valid) {
    split_chw(p.src, e, ch, h, w);
    if (kEncoder) sym = fetch_symbol(p.src, n, e, ch, h, w);
    ci = fetch_index(p.src, n, e, ch, h, w, R.tab);
    if (ci < 0 || ci >= p.tb.n_cdf) { bad = true; ci = 0; }
    sk = ilv_marked(R, use_marks ? p.skip : nullptr, ci);
  }
  const uint32_t m_sk = __ballot_sync(kFull, sk);
  const uint32_t m_rg = __ballot_sync(kFull, valid && !sk);
  const bool fl = kEncoder && __ballot_sync(kFull, sk && sym != 0) != 0u;
  if (lane == 0) {
    cnt1[g] = __popc(m_rg) + (m_sk != 0u ? 1 : 0);
    cnt2[g] = fl ? __popc(m_sk) : 0;
    if (m_sk) any_mark = 1;
  }
  __syncthreads();
  int n1, n2;
  const int s1 = __shfl_sync(kFull, warp_excl_scan(cnt1[lane], lane, n1), g);
  const int s2 = __shfl_sync(kFull, warp_excl_scan(cnt2[lane], lane, n2), g);
  const unsigned lt = (1u << lane) - 1u;
  // slot of this position in the item order (-1: an implied zero, not coded)
  int slot = -1;
  if (valid && !sk) slot = s1 + __popc(m_rg & lt);
  else if (valid && fl) slot = n1 + s2 + __popc(m_sk & lt);
  const int flag_slot = s1 + __popc(m_rg);       // of the group's flag, when it has one
  const int2 so = ilv_size_off(R, p.tb, ci);
  if (kEncoder) {
    IlvEncChunk& O = p.enc[blockIdx.x];
    auto put = [&](int at, uint32_t start, uint32_t freq, uint32_t nops, uint32_t raw) {
      unsigned long long rcp;
      uint32_t shift, bias;
      if (freq < 2u) {
        rcp = ~0ull; shift = 0; bias = start + (1u << 16) - 1u;
      } else {
        const uint32_t sh = 32u - (uint32_t)__clz((int)(freq - 1u));  // ceil(log2(freq))
        const unsigned long long x1 = 1ull << (sh + 31);
        const unsigned long long t1 = x1 / freq;
        const unsigned long long x0 = (unsigned long long)(freq - 1u) + ((x1 % freq) << 32);
        rcp = x0 / freq + (t1 << 32);
        shift = sh - 1u;
        bias = start;
      }
      O.rcp[at] = rcp; O.bias[at] = bias; O.fs[at] = freq | (shift << 17) | (nops << 22); O.raw[at] = raw;
    };
    bool esc = false;
    if (slot >= 0) {
      const int32_t* __restrict__ row = p.tb.cdf + (long long)ci * p.tb.cdf_stride;
      const int max_value = so.x - 2;
      int value = sym - so.y;
      uint32_t raw = 0, nops = 1;
      if (value < 0) {
        raw = (uint32_t)(-2 * value - 1);
        value = max_value;
      } else if (value >= max_value) {
        raw = (uint32_t)(2 * (value - max_value));
        value = max_value;
      }
      if (max_value < 0) { bad = true; value = 0; }
      else if (value == max_value) {  // bypass: count nibble + data nibbles
        uint32_t nb = 0;
        while (nb < 8 && (raw >> (nb * 4)) != 0) ++nb;
        nops = 2 + nb;
        esc = true;
      }
      const uint32_t start = (uint32_t)__ldg(row + value);
      uint32_t freq = (uint32_t)__ldg(row + value + 1) - start;
      if (freq == 0u || freq > (1u << 16)) { bad = true; freq = 1u; }
      put(slot, start, freq, nops, raw);
    }
    if (lane == 0 && m_sk) {   // the group's flag: P(1) = 8k / 65536
      const uint32_t f1 = 8u * (uint32_t)__popc(m_sk);
      put(flag_slot, fl ? (1u << 16) - f1 : 0u, fl ? f1 : (1u << 16) - f1, 1u, 0u);
    }
    if (esc) any_esc = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
      IlvMeta m;
      m.n1 = n1; m.n2 = n2; m.flags = (any_esc ? 1 : 0) | (any_mark ? 2 : 0); m.pad = 0;
      O.meta = m;
    }
  } else {
    IlvDecChunk& O = p.dec[blockIdx.x];
    if (valid) p.pos_ci[(long long)blockIdx.x * kIlvChunk + threadIdx.x] = (uint16_t)ci;
    if (valid && sk) p.sym_out[(long long)n * p.src.L + e] = 0;   // unless its group turns out flagged
    if (slot >= 0)
      O.item[slot] = make_uint4(p.row_start ? __ldg(p.row_start + ci)
                                            : (uint32_t)((long long)ci * p.tb.cdf_stride),
                                (uint32_t)max(so.x, 2), (uint32_t)so.y,
                                (uint32_t)threadIdx.x | ((uint32_t)ci << 16));
    if (lane == 0) {
      O.skm[g] = m_sk;
      if (m_sk)
        O.item[flag_slot] = make_uint4((1u << 16) - 8u * (uint32_t)__popc(m_sk), 3u, 0u,
                                       (uint32_t)(kIlvChunk + g));
    }
    // idle items up to a whole round of 32: the chain warp loads its items without a bound test
    if (threadIdx.x < 32 && n1 + (int)threadIdx.x < ((n1 + 31) & ~31))
      O.item[n1 + threadIdx.x] = make_uint4(0u, 2u, 0u, kIdleItem);
    if (threadIdx.x == 0) {
      IlvMeta m;
      m.n1 = n1; m.n2 = 0; m.flags = any_mark ? 2 : 0; m.pad = 0;
      O.meta = m;
    }
  }
  if (bad && p.status) atomicOr(p.status, 1);
}

// ---- encoder chain ------------------------------------------------------------
struct IlvEncState {
  unsigned long long x;
  uint32_t* reg;     // scratch region of this sub-stream
  int wpos;          // next free word is reg[wpos - 1] (words grow downwards)
};
struct IlvEncRec {
  unsigned long long rcp;
  uint32_t bias, fs, raw;
};
constexpr uint32_t kIdleFs = 0x1ffffu;   // record of an idle lane: never renormalises, q = 0
__device__ __forceinline__ IlvEncRec ilv_enc_rec(const IlvEncChunk& T, int base, int k, int n_items) {
  IlvEncRec r;
  r.rcp = 0ull; r.bias = 0u; r.fs = kIdleFs; r.raw = 0u;
  if (k >= 0 && k < n_items) {
    r.rcp = T.rcp[base + k]; r.bias = T.bias[base + k]; r.fs = T.fs[base + k]; r.raw = T.raw[base + k];
  }
  return r;
}

// One pass of a chunk, rounds last to first; the record of the round after this one is fetched
// while this one is in the chain.  kBypass = false: no item of the chunk is bypass-coded, a
// round is one straight-line Rans64EncPut per lane.
template <bool kBypass>
__device__ __forceinline__ void ilv_encode_pass(const IlvEncChunk& T, int base, int n_items,
                                                IlvEncState& E, int lane) {
  if (n_items <= 0) return;
  const unsigned gt = lane == 31 ? 0u : (kFull << (lane + 1));
  int r0 = ((n_items - 1) / 32) * 32;
  IlvEncRec cur = ilv_enc_rec(T, base, r0 + lane, n_items);
  for (; r0 >= 0; r0 -= 32) {
    const IlvEncRec nxt = ilv_enc_rec(T, base, r0 - 32 + lane, n_items);
    const uint32_t freq = cur.fs & 0x1ffffu, shift = (cur.fs >> 17) & 31u;
    if (kBypass) {
      const int nops = (int)(cur.fs >> 22);          // 0 for an idle lane
      const int tmax = __reduce_max_sync(kFull, nops);
      for (int t = tmax - 1; t >= 1; --t) {          // Rans64EncPutBits(4): count, data nibbles
        const bool doing = t < nops;
        const bool emit = doing && E.x >= (1ull << 59);
        const unsigned m = __ballot_sync(kFull, emit);
        if (emit) {
          E.reg[E.wpos - 1 - __popc(m & gt)] = (uint32_t)E.x;
          E.x >>= 32;
        }
        E.wpos -= __popc(m);
        if (doing) {
          const uint32_t v = t == 1 ? (uint32_t)(nops - 2) : (cur.raw >> ((t - 2) * 4)) & 15u;
          E.x = (E.x << 4) | v;
        }
      }
    }
    {  // Rans64EncPut: renormalise when x >= freq << 47, then x = (x / freq) << 16 + x % freq + start
      const bool emit = (uint32_t)(E.x >> 47) >= freq;     // idle lane: freq = 0x1ffff, never
      const unsigned m = __ballot_sync(kFull, emit);
      if (emit) {
        E.reg[E.wpos - 1 - __popc(m & gt)] = (uint32_t)E.x;
        E.x >>= 32;
      }
      E.wpos -= __popc(m);
      const unsigned long long q = __umul64hi(E.x, cur.rcp) >> shift;   // idle lane: rcp = 0
      E.x = E.x + cur.bias + q * (unsigned long long)((1u << 16) - freq);
    }
    cur = nxt;
  }
}

struct IlvEncShared {
  IlvEncChunk stage[kIlvRing];
  unsigned long long bar[kIlvRing];
};

// One warp per sub-stream.  Lane 0 keeps kIlvRing - 1 chunk copies in flight; the warp does
// nothing but walk the chains.
__global__ void __launch_bounds__(32) rans_ilv_encode_kernel(const EncP p) {
  extern __shared__ __align__(128) unsigned char ilv_smem[];
  IlvEncShared& S = *reinterpret_cast<IlvEncShared*>(ilv_smem);
  const int lane = threadIdx.x;
  const int j = blockIdx.x, n = blockIdx.y;
  const long long s_begin = (long long)j * p.S;
  const long long s_end = min(p.src.L, s_begin + p.S);
  const int n_chunks = (int)((s_end - s_begin + kIlvChunk - 1) / kIlvChunk);
  // chunks last to first
  const IlvEncChunk* last = p.ilv_enc + (long long)n * p.chunks_per_sample + s_begin / kIlvChunk +
                            (n_chunks - 1);
  if (lane == 0) {
    for (int s = 0; s < kIlvRing; ++s) ilv_mbar_init(ilv_smem_u32(&S.bar[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < min(kIlvRing - 1, n_chunks); ++i) {
      ilv_mbar_expect_tx(ilv_smem_u32(&S.bar[i]), (uint32_t)sizeof(IlvEncChunk));
      ilv_bulk_g2s(ilv_smem_u32(&S.stage[i]), last - i, (uint32_t)sizeof(IlvEncChunk),
                   ilv_smem_u32(&S.bar[i]));
    }
  }
  __syncwarp();
  IlvEncState E;
  E.x = 1ull << 31;  // Rans64EncInit, every lane
  E.reg = p.stream_data + ((long long)n * p.n_streams + j) * p.cap;
  E.wpos = p.cap;
#ifdef DVC_ILV_PROF
  long long pf_wait = 0, pf_pass = 0, pf_t0 = clock64(), pf_a, pf_b;
#define PF(acc) pf_b = clock64(); acc += pf_b - pf_a; pf_a = pf_b;
#else
#define PF(acc)
#endif
  for (int i = 0; i < n_chunks; ++i) {
    const int s = i % kIlvRing;
#ifdef DVC_ILV_PROF
    pf_a = clock64();
#endif
    __syncwarp();   // every lane is through with the stage that is refilled now
    if (lane == 0 && i + kIlvRing - 1 < n_chunks) {
      const int f = (i + kIlvRing - 1) % kIlvRing;
      ilv_mbar_expect_tx(ilv_smem_u32(&S.bar[f]), (uint32_t)sizeof(IlvEncChunk));
      ilv_bulk_g2s(ilv_smem_u32(&S.stage[f]), last - (i + kIlvRing - 1), (uint32_t)sizeof(IlvEncChunk),
                   ilv_smem_u32(&S.bar[f]));
    }
    ilv_mbar_wait(ilv_smem_u32(&S.bar[s]), (uint32_t)((i / kIlvRing) & 1));
    const IlvEncChunk& T = S.stage[s];
    const int n1 = T.meta.n1, n2 = T.meta.n2;
    PF(pf_wait)
    // last item first (pass 2 is decoded after pass 1)
    if (T.meta.flags & 1) {
      ilv_encode_pass<true>(T, n1, n2, E, lane);
      ilv_encode_pass<true>(T, 0, n1, E, lane);
    } else {
      ilv_encode_pass<false>(T, n1, n2, E, lane);
      ilv_encode_pass<false>(T, 0, n1, E, lane);
    }
    PF(pf_pass)
  }
#ifdef DVC_ILV_PROF
  if (lane == 0 && j == 0 && n == 0)
    printf("enc chain: chunks %d total %lld wait %lld pass %lld\n", n_chunks, clock64() - pf_t0,
           pf_wait, pf_pass);
#endif
  // ---- flush: mask, then the states (1 or 2 words each), lane 0 first -------------
  const bool wide = (E.x >> 32) != 0ull;
  const unsigned mask = __ballot_sync(kFull, wide);
  const int total = 33 + __popc(mask);
  const int at = E.wpos - total + 1 + lane + __popc(mask & ((1u << lane) - 1u));
  E.reg[at] = (uint32_t)E.x;
  if (wide) E.reg[at + 1] = (uint32_t)(E.x >> 32);
  if (lane == 0) {
    E.reg[E.wpos - total] = mask;
    p.stream_words[(long long)n * p.n_streams + j] = (uint32_t)(p.cap - (E.wpos - total));
  }
}

// ---- decoder chain ------------------------------------------------------------
struct IlvDecShared {
  IlvDecChunk stage[kIlvRing];
  unsigned long long bar[kIlvRing + 1];   // + the packed tables
  IlvRows rows;                    // pass 2 only
  uint4 item2[kIlvChunk];          // pass 2, built from the decoded flags
  uint32_t flag[32];
  uint32_t wnext[2][32];           // look-ahead of the word window, filled by cp.async
  alignas(16) uint16_t pack[8];    // the packed tables (kPack)
};
static_assert(offsetof(IlvDecShared, pack) % 16 == 0, "bulk-copy destination");
static_assert(sizeof(IlvDecShared) + kPackMaxBytes <= 227 * 1024, "chain CTA exceeds the 227 KB of an SM");

struct IlvDecState {
  unsigned long long x;
  const uint32_t* wp;   // words of this sub-stream
  int wn, base;         // their number; next unread word
  uint32_t window;      // words [base, base + 32), one per lane: a renormalising lane takes its
                        // word with a shuffle instead of a dependent load
  int wbase;            // packed-table passes: `window` / `w1` hold words [wbase, wbase + 64),
  uint32_t w1;          // 0 <= base - wbase < 32 between rounds; words [wbase + 64, wbase + 96)
  int wslot;            // are on their way into IlvDecShared::wnext[wslot] (cp.async)
  bool malformed;
#ifdef DVC_ILV_PROF
  long long prof[3];
#endif
  __device__ __forceinline__ uint32_t word(int i) const {
    return (i < wn) ? __ldg(wp + i) : 0u;   // a corrupt stream reads zeros, never out of bounds
  }
};

// One pass of a chunk.  Decoded values go straight to global memory (`out` = the chunk's slice
// of the symbol tensor; in a chunk without marked positions item k is position k: coalesced).
// This version bisects the CDF rows in global memory (tables too large to stage, or no
// `cdf_pack` given); the packed-table version follows.
template <bool kFlags>
__device__ __forceinline__ void ilv_decode_pass(const DecP& p, IlvDecShared& S, const uint4* items,
                                                int n_items, int32_t* out, IlvDecState& D,
                                                int lane) {
  if (n_items <= 0) return;
  const unsigned lt = (1u << lane) - 1u;
  const uint4 idle = make_uint4(0u, 2u, 0u, kIdleItem);
  uint4 cur = lane < n_items ? items[lane] : idle;
#ifdef DVC_ILV_PROF
  long long q0, q1;
#define QF(k) q1 = clock64(); D.prof[k] += q1 - q0; q0 = q1;
#else
#define QF(k)
#endif
  for (int r0 = 0; r0 < n_items; r0 += 32) {
#ifdef DVC_ILV_PROF
    q0 = clock64();
#endif
    const uint4 nxt = r0 + 32 + lane < n_items ? items[r0 + 32 + lane] : idle;
    const uint32_t it = cur.w & 0xffffu;
    const bool act = it != kIdleItem;
    const bool isflag = kFlags && act && it >= (uint32_t)kIlvChunk;
    const int size = (int)cur.y;
    const uint32_t cum = (uint32_t)(D.x & 0xffffu);  // Rans64DecGet
    // s = max j in [0, size-1) with row[j] <= cum  (row[0] = 0, row[size-1] = 2^16).
    // invariant row[lo] <= cum < row[hi]; next = 0 while row[hi] has not been read
    int lo = 0;
    uint32_t start = 0, next = 1u << 16;
    if (isflag) {          // the row {0, 65536 - 8k, 65536}
      if (cum >= cur.x) { lo = 1; start = cur.x; }
      else next = cur.x;
    } else if (act) {
      int hi = size - 1;
      next = 0u;
      const int32_t* __restrict__ row = p.tb.cdf + cur.x;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        const uint32_t v = (uint32_t)__ldg(row + mid);
        if (cum >= v) { lo = mid; start = v; }
        else { hi = mid; next = v; }
      }
      if (next == 0u) next = (uint32_t)__ldg(row + hi);
    }
    QF(0)
    // Rans64DecAdvance
    if (act) D.x = (unsigned long long)(next - start) * (D.x >> 16) + cum - start;
    {
      const bool need = act && D.x < (1ull << 31);
      const unsigned m = __ballot_sync(kFull, need);
      const uint32_t wv = __shfl_sync(kFull, D.window, __popc(m & lt));
      if (need) D.x = (D.x << 32) | wv;
      D.base += __popc(m);
    }
    QF(1)
    // bypass (Rans64DecGetBits(4) chain): one operation per lane and step, lanes in order
    const bool esc = act && !isflag && lo == size - 2;
    uint32_t raw = 0;
    if (__any_sync(kFull, esc)) {
      int ph = esc ? 1 : 0;  // 1: count nibbles, 2: data nibbles
      int nb = 0, kk = 0;
      do {
        const bool doing = ph != 0;
        const uint32_t val = (uint32_t)(D.x & 15u);
        if (doing) D.x >>= 4;
        const bool need = doing && D.x < (1ull << 31);
        const unsigned m = __ballot_sync(kFull, need);
        if (need) D.x = (D.x << 32) | D.word(D.base + __popc(m & lt));
        D.base += __popc(m);
        if (ph == 1) {
          nb += (int)val;
          if (val != 15u || nb > 64) {   // the encoder never writes more than 8 data nibbles
            if (nb > 64) { D.malformed = true; nb = 64; }
            ph = nb > 0 ? 2 : 0;
          }
        } else if (ph == 2) {
          if (kk < 8) raw |= val << (kk * 4);
          if (++kk == nb) ph = 0;
        }
      } while (__any_sync(kFull, ph != 0));
    }
    D.window = D.word(D.base + lane);
    if (act) {
      if (isflag) {
        S.flag[it - kIlvChunk] = (uint32_t)lo;
      } else {
        int value = lo;
        if (esc) {
          value = (int)(raw >> 1);
          if (raw & 1u) value = -value - 1;
          else value += size - 2;
        }
        out[it] = value + (int)cur.z;
      }
    }
    cur = nxt;
    QF(2)
  }
}

// Look-ahead of the word window: lane l's word `idx` goes global -> shared with a 4-byte cp.async
// (zero past the end of the sub-stream), so no register waits for it: a register load whose
// value the compiler copies at the join of the slide branch stalled the chain for an L2 round trip
// at every slide (21 % of the kernel's stall samples, profiles/r02_coder.md).
__device__ __forceinline__ void ilv_fetch_word(uint32_t* dst, const IlvDecState& D, int idx) {
  const bool ok = idx < D.wn;
  const uint32_t* src = D.wp + (ok ? idx : 0);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;"
               ::"r"(ilv_smem_u32(dst)), "l"(src), "r"(ok ? 4 : 0) : "memory");
}
__device__ __forceinline__ void ilv_fetch_wait() {
  asm volatile("cp.async.wait_all;" ::: "memory");
}

// The same pass for the packed tables, written as ONE basic block per round.  The chain warp is
// alone on its SM: nothing hides an instruction's latency except the independent instructions
// the scheduler finds next to it, and every divergent branch of the version above (idle lanes,
// the probe selection, the flag rows) both serialised its arms and stopped the compiler from
// moving the round's off-chain work (next item, stores, word window) into the shadow of the
// chain's loads and shuffles.  Here
//   * an idle lane runs the chain arithmetic on the row {0, 2^16}: x' = 2^16 (x >> 16) + cum = x,
//     so no predicate guards the search, the advance or the renormalisation;
//   * the four probes are settled with a depth-2 tree of selects;
//   * the two rare cases (symbol beyond the probes, bypass-coded value) sit behind warp-uniform
//     votes;
//   * the word window is 64 words deep (two registers per lane) and its next 32 words travel
//     global -> shared by cp.async a whole slide (~4 rounds) ahead (ilv_fetch_word).
// Same arithmetic, same word order: byte-for-byte the container of the encoder above.
template <bool kFlags>
__device__ __forceinline__ void ilv_decode_pass_packed(IlvDecShared& S, const uint16_t* lut,
                                                       const uint16_t* tbl, const uint4* items,
                                                       int n_items, int32_t* out, IlvDecState& D,
                                                       int lane) {
  if (n_items <= 0) return;
  const unsigned lt = (1u << lane) - 1u;
  // the item list is padded with idle items to a whole number of rounds (by whoever built it)
  uint4 cur = items[lane];
  int wbase = D.wbase, wslot = D.wslot;
  uint32_t w0 = D.window, w1 = D.w1;
  asm volatile("" : "+l"(out));   // keep the base in registers (otherwise re-derived every round)
#pragma unroll 1
  for (int r0 = 0; r0 < n_items; r0 += 32) {
    uint4 nxt = cur;
    if (r0 + 32 < n_items) nxt = items[r0 + 32 + lane];
    const uint32_t it = cur.w & 0xffffu;
    const bool act = it != kIdleItem;
    const bool isflag = kFlags && act && it >= (uint32_t)kIlvChunk;
    const bool regular = act && !isflag;
    const int size = (int)cur.y;
    const uint32_t cum = (uint32_t)D.x & 0xffffu;  // Rans64DecGet
    // ---- search: every lane probes a valid row (idle lanes and flags: row 0 from its start)
    const uint32_t row_at = (kFlags && isflag) ? 0u : cur.x;
    const uint16_t* const lut_at = lut + (cur.w >> 16) * kLutStride + lut_key(cum);
    const uint32_t lo0 = lut_at[0];
    const uint16_t* r = tbl + row_at + lo0;
    const uint32_t b0 = r[0], b1 = r[1], b2 = r[2], b3 = r[3], b4 = r[4];
    const bool c1 = cum > b1, c2 = cum > b2, c3 = cum > b3;
    const uint32_t s01 = c1 ? b1 : b0, s23 = c3 ? b3 : b2;
    const uint32_t n01 = c1 ? b2 : b1, n23 = c3 ? b4 : b3;
    uint32_t sm1 = c2 ? s23 : s01, nm1 = c2 ? n23 : n01;
    int lo = (int)lo0 + (c1 ? 1 : 0) + (c2 ? 1 : 0) + (c3 ? 1 : 0);
    if (__builtin_expect(__any_sync(kFull, regular && cum > b4), 0)) {   // rare: bisect (lo0 + 4, hi)
      if (regular && cum > b4) {
        int hi = (int)lut_at[1] + 1;   // the next key's first symbol ends above every cum of this key
        lo = (int)lo0 + 4; sm1 = b4; nm1 = 0u;
        bool have_next = false;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          const uint32_t v = tbl[cur.x + mid];
          if (cum > v) { lo = mid; sm1 = v; }
          else { hi = mid; nm1 = v; have_next = true; }
        }
        if (!have_next) nm1 = tbl[cur.x + hi];
      }
    }
    uint32_t start = (sm1 + 1u) & 0xffffu, next = nm1 + 1u;
    if (kFlags) {          // the row {0, 65536 - 8k, 65536}
      const bool set = cum >= cur.x;
      if (isflag) { lo = set ? 1 : 0; start = set ? cur.x : 0u; next = set ? (1u << 16) : cur.x; }
    }
    if (!act) { start = 0u; next = 1u << 16; }
    // ---- Rans64DecAdvance + renormalisation (ballot-ordered words from the window)
    D.x = (unsigned long long)(next - start) * (D.x >> 16) + (cum - start);
    {
      const bool need = D.x < (1ull << 31);
      const unsigned m = __ballot_sync(kFull, need);
      const int at = D.base - wbase + __popc(m & lt);          // < 64
      const uint32_t v0 = __shfl_sync(kFull, w0, at & 31);
      const uint32_t v1 = __shfl_sync(kFull, w1, at & 31);
      if (need) D.x = (D.x << 32) | (at < 32 ? v0 : v1);
      D.base += __popc(m);
    }
    // ---- bypass (Rans64DecGetBits(4) chain): one operation per lane and step, lanes in order
    const bool esc = regular && lo == size - 2;
    int value = lo;
    if (__builtin_expect(__any_sync(kFull, esc), 0)) {
      uint32_t raw = 0;
      int ph = esc ? 1 : 0;  // 1: count nibbles, 2: data nibbles
      int nb = 0, kk = 0;
      do {
        const bool doing = ph != 0;
        const uint32_t val = (uint32_t)(D.x & 15u);
        if (doing) D.x >>= 4;
        const bool need = doing && D.x < (1ull << 31);
        const unsigned m = __ballot_sync(kFull, need);
        if (need) D.x = (D.x << 32) | D.word(D.base + __popc(m & lt));
        D.base += __popc(m);
        if (ph == 1) {
          nb += (int)val;
          if (val != 15u || nb > 64) {   // the encoder never writes more than 8 data nibbles
            if (nb > 64) { D.malformed = true; nb = 64; }
            ph = nb > 0 ? 2 : 0;
          }
        } else if (ph == 2) {
          if (kk < 8) raw |= val << (kk * 4);
          if (++kk == nb) ph = 0;
        }
      } while (__any_sync(kFull, ph != 0));
      if (esc) {
        value = (int)(raw >> 1);
        if (raw & 1u) value = -value - 1;
        else value += size - 2;
      }
      ilv_fetch_wait();
      wbase = D.base;                    // the words were read past the window: start it over
      w0 = D.word(wbase + lane);
      w1 = D.word(wbase + 32 + lane);
      ilv_fetch_word(&S.wnext[wslot][lane], D, wbase + 64 + lane);
    } else if (D.base - wbase >= 32) {   // warp-uniform: slide the window by 32 words
      wbase += 32;
      w0 = w1;
      ilv_fetch_wait();                  // issued a whole slide ago
      w1 = S.wnext[wslot][lane];
      wslot ^= 1;                        // the other slot was read a slide ago
      ilv_fetch_word(&S.wnext[wslot][lane], D, wbase + 64 + lane);
    }
    if (kFlags && isflag) S.flag[it - kIlvChunk] = (uint32_t)lo;
    if (regular) out[it] = value + (int)cur.z;
    cur = nxt;
  }
  D.wbase = wbase; D.wslot = wslot; D.window = w0; D.w1 = w1;
}

template <bool kPack>
__global__ void __launch_bounds__(32) rans_ilv_decode_kernel(const DecP p) {
  extern __shared__ __align__(128) unsigned char ilv_smem[];
  IlvDecShared& S = *reinterpret_cast<IlvDecShared*>(ilv_smem);
  const int lane = threadIdx.x;
  const int j = blockIdx.x, n = blockIdx.y;
  const long long s_begin = (long long)j * p.S;
  const long long s_end = min(p.src.L, s_begin + p.S);
  const int n_chunks = (int)((s_end - s_begin + kIlvChunk - 1) / kIlvChunk);
  const long long first_chunk = (long long)n * p.chunks_per_sample + s_begin / kIlvChunk;
  const IlvDecChunk* first = p.ilv_dec + first_chunk;
  const PackLayout pl = pack_layout(p.tb.n_cdf, p.pack_entries);
  if (lane == 0) {
    for (int s = 0; s <= kIlvRing; ++s) ilv_mbar_init(ilv_smem_u32(&S.bar[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (kPack) {
      ilv_mbar_expect_tx(ilv_smem_u32(&S.bar[kIlvRing]), (uint32_t)pl.bytes);
      ilv_bulk_g2s(ilv_smem_u32(S.pack), p.pack, (uint32_t)pl.bytes, ilv_smem_u32(&S.bar[kIlvRing]));
    }
    for (int i = 0; i < min(kIlvRing - 1, n_chunks); ++i) {
      ilv_mbar_expect_tx(ilv_smem_u32(&S.bar[i]), (uint32_t)sizeof(IlvDecChunk));
      ilv_bulk_g2s(ilv_smem_u32(&S.stage[i]), first + i, (uint32_t)sizeof(IlvDecChunk),
                   ilv_smem_u32(&S.bar[i]));
    }
  }
  ilv_load_rows(S.rows, p.src, p.tb, p.skip);
  __syncwarp();
  const uint16_t* const lut = S.pack;
  const uint16_t* const tbl = S.pack + pl.tbl_off / 2;
  const uint32_t* words = reinterpret_cast<const uint32_t*>(p.in + n * p.in_stride);
  const long long total_words = __ldg(p.in_bytes + n) >> 2;
  IlvDecState D;
  D.malformed = false;
#ifdef DVC_ILV_PROF
  D.prof[0] = D.prof[1] = D.prof[2] = 0;
#endif
  {
    bool ok = total_words >= 4 + p.n_streams && __ldg(words) == (p.skip ? kMagic3S : kMagic3) &&
              __ldg(words + 1) == (uint32_t)p.src.L && __ldg(words + 2) == (uint32_t)p.S &&
              __ldg(words + 3) == (uint32_t)p.n_streams;
    unsigned long long before = 0;
    if (ok)
      for (int i = lane; i < j; i += 32) before += __ldg(words + 4 + i);
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(kFull, before, o);
    const unsigned long long mine = ok ? __ldg(words + 4 + j) : 0ull;
    const long long first_word = 4 + p.n_streams + (long long)before;
    if (!ok || mine < 33ull || mine > 0x7fffffffull || first_word + (long long)mine > total_words) {
      D.malformed = true;
      D.wp = words;   // decode zeros: defined, flagged
      D.wn = 0;
    } else {
      D.wp = words + first_word;
      D.wn = (int)mine;
    }
  }
  {  // the 32 states
    const uint32_t mask = D.word(0);
    const int at = 1 + lane + __popc(mask & ((1u << lane) - 1u));
    D.x = (unsigned long long)D.word(at);
    if ((mask >> lane) & 1u) D.x |= (unsigned long long)D.word(at + 1) << 32;
    D.base = 33 + __popc(mask);
    D.window = D.word(D.base + lane);
    D.wbase = D.base;
    D.w1 = D.word(D.base + 32 + lane);
    D.wslot = 0;
    if (kPack) ilv_fetch_word(&S.wnext[0][lane], D, D.base + 64 + lane);
  }
  int32_t* const sym_out = p.ilv_sym + (long long)n * p.src.L;
  if (kPack) ilv_mbar_wait(ilv_smem_u32(&S.bar[kIlvRing]), 0u);
#ifdef DVC_ILV_PROF
  long long pf_wait = 0, pf_pass = 0, pf_t0 = clock64(), pf_a, pf_b;
#endif
  for (int i = 0; i < n_chunks; ++i) {
    const int s = i % kIlvRing;
#ifdef DVC_ILV_PROF
    pf_a = clock64();
#endif
    __syncwarp();   // every lane is through with the stage that is refilled now
    if (lane == 0 && i + kIlvRing - 1 < n_chunks) {
      const int f = (i + kIlvRing - 1) % kIlvRing;
      ilv_mbar_expect_tx(ilv_smem_u32(&S.bar[f]), (uint32_t)sizeof(IlvDecChunk));
      ilv_bulk_g2s(ilv_smem_u32(&S.stage[f]), first + (i + kIlvRing - 1), (uint32_t)sizeof(IlvDecChunk),
                   ilv_smem_u32(&S.bar[f]));
    }
    ilv_mbar_wait(ilv_smem_u32(&S.bar[s]), (uint32_t)((i / kIlvRing) & 1));
    const IlvDecChunk& T = S.stage[s];
    const long long cb = s_begin + (long long)i * kIlvChunk;
    PF(pf_wait)
    // pass 1: regular symbols and group flags
    if (!(T.meta.flags & 2)) {
      if (kPack) ilv_decode_pass_packed<false>(S, lut, tbl, T.item, T.meta.n1, sym_out + cb, D, lane);
      else ilv_decode_pass<false>(p, S, T.item, T.meta.n1, sym_out + cb, D, lane);
    } else {
      if (kPack) ilv_decode_pass_packed<true>(S, lut, tbl, T.item, T.meta.n1, sym_out + cb, D, lane);
      else ilv_decode_pass<true>(p, S, T.item, T.meta.n1, sym_out + cb, D, lane);
      __syncwarp();
      // pass 2: the marked symbols of the groups whose flag came out set (lane g <-> group g);
      // the others keep the zeros the prepare kernel wrote
      const uint32_t skm = T.skm[lane];
      unsigned fm = __ballot_sync(kFull, skm != 0u && S.flag[lane] != 0u);
      if (fm) {
        int n2 = 0;
        while (fm) {
          const int g = __ffs((int)fm) - 1;
          fm &= fm - 1u;
          const uint32_t sm = __shfl_sync(kFull, skm, g);
          if ((sm >> lane) & 1u) {
            const int pos = 32 * g + lane;
            const int ci = __ldg(p.ilv_ci + (first_chunk + i) * kIlvChunk + pos);
            const int2 so = ilv_size_off(S.rows, p.tb, ci);
            const uint32_t at = kPack ? reinterpret_cast<const uint32_t*>(S.pack + pl.start_off / 2)[ci]
                                      : (uint32_t)((long long)ci * p.tb.cdf_stride);
            S.item2[n2 + __popc(sm & ((1u << lane) - 1u))] =
                make_uint4(at, (uint32_t)max(so.x, 2), (uint32_t)so.y, (uint32_t)pos | ((uint32_t)ci << 16));
          }
          n2 += __popc(sm);
        }
        if (n2 + lane < ((n2 + 31) & ~31)) S.item2[n2 + lane] = make_uint4(0u, 2u, 0u, kIdleItem);
        __syncwarp();
        if (kPack) ilv_decode_pass_packed<false>(S, lut, tbl, S.item2, n2, sym_out + cb, D, lane);
        else ilv_decode_pass<false>(p, S, S.item2, n2, sym_out + cb, D, lane);
      }
    }
    PF(pf_pass)
  }
#ifdef DVC_ILV_PROF
  if (lane == 0 && j == 0 && n == 0)
    printf("dec chain: chunks %d total %lld wait %lld pass %lld | search %lld advance %lld tail %lld\n",
           n_chunks, clock64() - pf_t0, pf_wait, pf_pass, D.prof[0], D.prof[1], D.prof[2]);
#endif
  // a well-formed sub-stream ends with every lane back at the encoder's initial state and
  // every word consumed
  if (kPack) ilv_fetch_wait();   // no copy outlives the CTA
  if (D.x != (1ull << 31) || D.base != D.wn) D.malformed = true;
  if (p.status && __any_sync(kFull, D.malformed) && lane == 0) atomicOr(p.status, 2);
}

// float(symbol) + means -> out (EntropyModel.dequantize), data parallel
struct IlvFinishP {
  Source src;
  const int32_t* sym;      // [N][L]
  float* out_f;
  CTS os;
  int N;
};
__global__ void __launch_bounds__(256) ilv_finish_kernel(const IlvFinishP p) {
  const long long total = p.src.L * p.N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / p.src.L);
    const long long e = i - n * p.src.L;
    int c, h, w;
    split_chw(p.src, e, c, h, w);
    float f = (float)__ldg(p.sym + i);
    if (p.src.means)
      f = add_rn(f, __ldg(p.src.means + n * p.src.ms.n + c * p.src.ms.c + h * p.src.ms.h +
                          w * p.src.ms.w));
    p.out_f[n * p.os.n + c * p.os.c + h * p.os.h + w * p.os.w] = f;
  }
}

// ---------------------------------------------------------------------------
// Decoder side of the checkerboard dual prior (video_model.py:259-289 ==
// :433-464): element-wise glue between the two decoding passes and the
// spatial-prior conv.  q0 / q1 are the decoded symbol planes [N, C/2, H, W]
// (int32, contiguous).  Selection replaces the reference's multiply-by-mask and
// add-of-zero (equal values; zero signs may differ).
// ---------------------------------------------------------------------------
struct DecStageP {
  const int32_t* q0;
  const int32_t* q1;
  const float* means;   // [N, C, H, W]
  const float* scales;  // stage a: [N, C, H, W]
  const float* prior;   // stage b: [N, 2C, H, W] = (means_0, scales_0, means_1, scales_1)
  float* out;           // stage a: params [N, 3C, H, W]; stage b: y_hat [N, C, H, W]
  CTS ms, ss, ps, os;
  int N, C, H, W;
};

// params = cat((q0 + means_0) * mask_0, (q0 + means_1) * mask_1, means, scales)
__global__ void __launch_bounds__(256) decode_stage_a_kernel(const DecStageP p) {
  const int half = p.C >> 1, HW = p.H * p.W;
  const long long per = (long long)p.C * HW, total = per * p.N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / per);
    const int r = (int)(i - n * per);
    const int c = r / HW, r2 = r - c * HW, h = r2 / p.W, w = r2 - h * p.W;
    const float m = __ldg(p.means + n * p.ms.n + c * p.ms.c + h * p.ms.h + w * p.ms.w);
    const float s = __ldg(p.scales + n * p.ss.n + c * p.ss.c + h * p.ss.h + w * p.ss.w);
    const int cq = c < half ? c : c - half;
    const bool mine = (((h + w) & 1) == 0) == (c < half);  // first half: even cells, second: odd
    float v = 0.f;
    if (mine) v = add_rn((float)__ldg(p.q0 + ((long long)(n * half + cq) * p.H + h) * p.W + w), m);
    float* o = p.out + n * p.os.n + h * p.os.h + w * p.os.w;
    o[c * p.os.c] = v;
    o[(p.C + c) * p.os.c] = m;
    o[(2 * p.C + c) * p.os.c] = s;
  }
}

// y_hat = cat(y_hat_0_0 + y_hat_0_1, y_hat_1_1 + y_hat_1_0)
__global__ void __launch_bounds__(256) decode_stage_b_kernel(const DecStageP p) {
  const int half = p.C >> 1, HW = p.H * p.W;
  const long long per = (long long)p.C * HW, total = per * p.N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / per);
    const int r = (int)(i - n * per);
    const int c = r / HW, r2 = r - c * HW, h = r2 / p.W, w = r2 - h * p.W;
    const int cq = c < half ? c : c - half;
    const bool anchor = (((h + w) & 1) == 0) == (c < half);
    const long long qo = ((long long)(n * half + cq) * p.H + h) * p.W + w;
    float v;
    if (anchor) {
      v = add_rn((float)__ldg(p.q0 + qo),
                 __ldg(p.means + n * p.ms.n + c * p.ms.c + h * p.ms.h + w * p.ms.w));
    } else {
      // spatial-prior means: chunk 0 (first half) or chunk 2 (second half) of `prior`
      const int pc = c < half ? cq : 2 * half + cq;
      v = add_rn((float)__ldg(p.q1 + qo),
                 __ldg(p.prior + n * p.ps.n + pc * p.ps.c + h * p.ps.h + w * p.ps.w));
    }
    p.out[n * p.os.n + c * p.os.c + h * p.os.h + w * p.os.w] = v;
  }
}

static int fill_source(Source& s, const int32_t* symbols, const float* x, const float* means,
                       const int32_t* indexes, const float* scales, const float* scale_table,
                       int64_t T, float scale_bound, int64_t C, int64_t H, int64_t W,
                       const int64_t x_st[4], const int64_t means_st[4],
                       const int64_t scales_st[4], const char* who) {
  DVC_REQUIRE(C > 0 && H > 0 && W > 0, "%s: empty tensor", who);
  DVC_REQUIRE((long long)C * H * W < 2147483647LL, "%s: C*H*W too large", who);
  DVC_REQUIRE(!x || x_st, "%s: x without strides", who);
  DVC_REQUIRE(!means || means_st, "%s: means without strides", who);
  DVC_REQUIRE(!scales || (scales_st && scale_table && T >= 1 && T <= kMaxTable),
              "%s: scales need strides and a scale table of 1..%d entries", who, kMaxTable);
  s.symbols = symbols; s.x = x; s.means = means; s.indexes = indexes; s.scales = scales;
  s.scale_table = scale_table;
  s.T = scales ? (int)T : 0;
  s.scale_bound = scale_bound;
  s.xs = cts(x_st); s.ms = cts(means_st); s.ss = cts(scales_st);
  s.C = (int)C; s.H = (int)H; s.W = (int)W; s.HW = (int)(H * W);
  s.L = (long long)C * H * W;
  s.cb_parity = -1;
  s.cb_alt = 0;
  return DVC_OK;
}

static int fill_tables(Tables& t, const int32_t* cdf, const int32_t* cdf_size,
                       const int32_t* offset, int64_t n_cdf, int64_t cdf_stride, const char* who) {
  DVC_REQUIRE(cdf && cdf_size && offset, "%s: null CDF tables (run update() first)", who);
  DVC_REQUIRE(n_cdf >= 1 && cdf_stride >= 3 && n_cdf < (1 << 24) && cdf_stride < (1 << 24),
              "%s: bad CDF table extents", who);
  t.cdf = cdf; t.cdf_size = cdf_size; t.offset = offset;
  t.n_cdf = (int)n_cdf; t.cdf_stride = (int)cdf_stride;
  return DVC_OK;
}

struct Partition {
  long long S;
  int n_streams, cap, header;
};
static int make_partition(Partition& q, long long L, int64_t stream_symbols, int lanes,
                          const char* who) {
  DVC_REQUIRE(stream_symbols >= 0, "%s: stream_symbols must be >= 0", who);
  DVC_REQUIRE(lanes == 1 || (lanes == 32 && stream_symbols > 0 && stream_symbols % 1024 == 0),
              "%s: lanes must be 1, or 32 together with stream_symbols > 0 and a multiple of 1024 "
              "(a raw stream is a stock stream)", who);
  q.header = stream_symbols > 0 ? 1 : 0;
  q.S = q.header ? stream_symbols : L;
  const long long ns = (L + q.S - 1) / q.S;
  DVC_REQUIRE(ns >= 1 && ns <= (1 << 20), "%s: too many sub-streams (%lld)", who, ns);
  DVC_REQUIRE(2 * q.S + 4 < 2147483647LL, "%s: sub-stream too long", who);
  q.n_streams = (int)ns;
  // <= 52 bits per symbol + initial state + flush; lane-interleaved: 32 states (<= 65 words
  // with the mask) and <= 16 bits per group flag (one per 32 symbols)
  q.cap = lanes == 32 ? (int)(2 * q.S + 160) : (int)(2 * q.S + 4);
  return DVC_OK;
}

}  // namespace dvc

using namespace dvc;

extern "C" {

int dvc_symbols_indexes_fwd(const float* x, const float* means, const float* scales,
                            const float* table, int64_t T, int32_t* symbols, int32_t* indexes,
                            int64_t N, int64_t C, int64_t H, int64_t W, const int64_t x_st[4],
                            const int64_t means_st[4], const int64_t scales_st[4],
                            float scale_bound, dvc_stream_t stream) {
  DVC_REQUIRE(N > 0, "symbols_indexes: empty tensor");
  DVC_REQUIRE(!symbols || x, "symbols_indexes: symbols requested without x");
  DVC_REQUIRE(!indexes || scales, "symbols_indexes: indexes requested without scales");
  DVC_REQUIRE(symbols || indexes, "symbols_indexes: nothing to do");
  SymIdxP p;
  int rc = fill_source(p.src, nullptr, x, means, nullptr, scales, table, T, scale_bound, C, H, W,
                       x_st, means_st, scales_st, "symbols_indexes");
  if (rc) return rc;
  p.out_symbols = symbols;
  p.out_indexes = indexes;
  p.N = (int)N;
  const long long total = p.src.L * N;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  symbols_indexes_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("symbols_indexes_kernel");
}

// HOST function (setup time, once per model): CompressAI's
// _CXX.pmf_to_quantized_cdf as used by EntropyModel._pmf_to_cdf.  Scale the pmf
// to 2^precision, renormalise by the integer total, accumulate, pin the last
// entry, then give every zero-width symbol one count taken from the narrowest
// symbol that can spare it.
int dvc_pmf_to_quantized_cdf(const float* pmf, int64_t n, int precision, int32_t* cdf) {
  DVC_REQUIRE(pmf && cdf && n >= 1 && n < (1 << 24), "pmf_to_quantized_cdf: bad arguments");
  DVC_REQUIRE(precision >= 1 && precision <= 16, "pmf_to_quantized_cdf: precision in [1,16]");
  const uint32_t one = 1u << precision;
  uint32_t* c = reinterpret_cast<uint32_t*>(cdf);
  uint32_t total = 0;
  c[0] = 0;
  for (int64_t i = 0; i < n; ++i) {
    const float pr = pmf[i];
    if (!(pr >= 0.f) || pr > 3.0e38f)
      return fail(DVC_ERR_INVALID_ARGUMENT,
                  "Invalid `pmf`, non-finite or negative element found");
    const uint32_t f = (uint32_t)roundf(pr * (float)one);
    c[i + 1] = f;
    total += f;
  }
  if (total == 0)
    return fail(DVC_ERR_INVALID_ARGUMENT,
                "Invalid `pmf`: at least one element must have a non-zero probability.");
  uint32_t run = 0;
  for (int64_t i = 1; i <= n; ++i) {
    run += (uint32_t)(((uint64_t)one * c[i]) / total);
    c[i] = run;
  }
  c[n] = one;
  for (int64_t i = 0; i < n; ++i) {
    if (c[i] != c[i + 1]) continue;
    int64_t donor = -1;
    uint32_t narrowest = 0xffffffffu;
    for (int64_t k = 0; k < n; ++k) {
      const uint32_t width = c[k + 1] - c[k];
      if (width > 1 && width < narrowest) {
        narrowest = width;
        donor = k;
      }
    }
    if (donor < 0)
      return fail(DVC_ERR_INVALID_ARGUMENT, "pmf_to_quantized_cdf: more symbols than counts");
    if (donor < i) {
      for (int64_t k = donor + 1; k <= i; ++k) c[k] -= 1;
    } else {
      for (int64_t k = i + 1; k <= donor; ++k) c[k] += 1;
    }
  }
  return DVC_OK;
}

int64_t dvc_rans_scratch_bytes(int64_t N, int64_t L, int64_t stream_symbols, int lanes) {
  Partition q;
  if (N < 1 || L < 1 || make_partition(q, L, stream_symbols, lanes, "rans_scratch_bytes"))
    return -1;
  int64_t bytes = N * (int64_t)q.n_streams * (int64_t)(q.cap + 1) * 4;
  if (lanes == 32)   // + the coder records of every chunk (prepare kernel -> chain kernel)
    bytes = ((bytes + 15) / 16) * 16 + N * ((L + kIlvChunk - 1) / kIlvChunk) * (int64_t)sizeof(IlvEncChunk) + 16;
  return bytes;
}

int64_t dvc_rans_decode_scratch_bytes(int64_t N, int64_t L, int64_t stream_symbols, int lanes) {
  Partition q;
  if (N < 1 || L < 1 || make_partition(q, L, stream_symbols, lanes, "rans_decode_scratch_bytes"))
    return -1;
  if (lanes != 32) return 0;
  const int64_t chunks = N * ((L + kIlvChunk - 1) / kIlvChunk);
  // pass-1 items of every chunk, table row of every position, decoded symbols
  return chunks * (int64_t)sizeof(IlvDecChunk) + chunks * kIlvChunk * 2 + ((N * L * 4 + 15) / 16) * 16;
}

int64_t dvc_rans_max_bytes(int64_t L, int64_t stream_symbols, int lanes) {
  Partition q;
  if (L < 1 || make_partition(q, L, stream_symbols, lanes, "rans_max_bytes")) return -1;
  return ((q.header ? 4 + (int64_t)q.n_streams : 0) + (int64_t)q.n_streams * q.cap) * 4;
}

int dvc_rans_encode(const float* x, const float* means, const int32_t* symbols,
                    const int32_t* indexes, const float* scales, const float* scale_table,
                    int64_t T, float scale_bound, const int32_t* cdf, const int32_t* cdf_size,
                    const int32_t* offset, int64_t n_cdf, int64_t cdf_stride, uint8_t* out,
                    int64_t out_stride_bytes, int64_t* out_bytes, void* scratch, int* status,
                    int64_t N, int64_t C, int64_t H, int64_t W, const int64_t x_st[4],
                    const int64_t means_st[4], const int64_t scales_st[4],
                    int64_t stream_symbols, int lanes, const uint8_t* skip_rows,
                    int64_t skip_max_flagged, dvc_stream_t stream) {
  DVC_REQUIRE((x != nullptr) != (symbols != nullptr),
              "rans_encode: give exactly one of x / symbols");
  DVC_REQUIRE(!skip_rows || lanes == 32, "rans_encode: skip_rows needs the lane-interleaved layout");
  DVC_REQUIRE(out && out_bytes && scratch, "rans_encode: null output / scratch");
  DVC_REQUIRE(N > 0 && N <= 65535, "rans_encode: N must be in [1, 65535]");
  DVC_REQUIRE((out_stride_bytes % 4) == 0 && (reinterpret_cast<uintptr_t>(out) & 3u) == 0,
              "rans_encode: out must be 4-byte aligned with a 4-byte multiple stride");
  EncP p;
  int rc = fill_source(p.src, symbols, x, means, indexes, scales, scale_table, T, scale_bound, C,
                       H, W, x_st, means_st, scales_st, "rans_encode");
  if (rc) return rc;
  rc = fill_tables(p.tb, cdf, cdf_size, offset, n_cdf, cdf_stride, "rans_encode");
  if (rc) return rc;
  Partition q;
  rc = make_partition(q, p.src.L, stream_symbols, lanes, "rans_encode");
  if (rc) return rc;
  DVC_REQUIRE(lanes == 1 || n_cdf <= 65535, "rans_encode: more than 65535 table rows");
  p.S = q.S; p.n_streams = q.n_streams; p.cap = q.cap; p.N = (int)N;
  p.stream_words = reinterpret_cast<uint32_t*>(scratch);
  p.stream_data = p.stream_words + N * (int64_t)q.n_streams;
  p.status = status;
  p.skip = skip_rows;
  const unsigned* flagged = nullptr;
  long long flagged_max = 0;
  if (lanes == 32) {
    DVC_REQUIRE(n_cdf * cdf_stride < 2147483647LL, "rans_encode: CDF table too large");
    const int64_t cps = (p.src.L + kIlvChunk - 1) / kIlvChunk;
    const int64_t head = ((N * (int64_t)q.n_streams * (int64_t)(q.cap + 1) * 4 + 15) / 16) * 16;
    IlvEncChunk* chunks = reinterpret_cast<IlvEncChunk*>(reinterpret_cast<uint8_t*>(scratch) + head);
    DVC_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 15u) == 0, "rans_encode: scratch must be 16-byte aligned");
    IlvPrepP k;
    k.src = p.src; k.tb = p.tb; k.skip = skip_rows; k.enc = chunks; k.dec = nullptr; k.pos_ci = nullptr;
    k.sym_out = nullptr; k.row_start = nullptr;
    k.flagged = nullptr; k.flagged_max = 0;
    k.chunks_per_sample = (int)cps; k.status = status;
    if (skip_rows && skip_max_flagged >= 0) {
      // the counter lives behind the chunk records (dvc_rans_scratch_bytes reserves it)
      k.flagged = reinterpret_cast<unsigned*>(chunks + N * cps);
      k.flagged_max = skip_max_flagged;
      cudaError_t e0 = cudaMemsetAsync(k.flagged, 0, sizeof(unsigned), (cudaStream_t)stream);
      if (e0 != cudaSuccess)
        return fail(DVC_ERR_CUDA, "rans_encode: cudaMemsetAsync: %s", cudaGetErrorString(e0));
      ilv_count_kernel<<<(unsigned)(N * cps), kIlvChunk, 0, (cudaStream_t)stream>>>(k);
      rc = check_launch("ilv_count_kernel");
      if (rc) return rc;
    }
    flagged = k.flagged; flagged_max = k.flagged_max;
    ilv_prepare_kernel<true><<<(unsigned)(N * cps), kIlvChunk, 0, (cudaStream_t)stream>>>(k);
    rc = check_launch("ilv_prepare_kernel");
    if (rc) return rc;
    p.ilv_enc = chunks; p.chunks_per_sample = (int)cps;
    cudaError_t e = cudaFuncSetAttribute(rans_ilv_encode_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(IlvEncShared));
    if (e != cudaSuccess)
      return fail(DVC_ERR_CUDA, "rans_encode: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    rans_ilv_encode_kernel<<<dim3((unsigned)q.n_streams, (unsigned)N), 32, sizeof(IlvEncShared),
                             (cudaStream_t)stream>>>(p);
    rc = check_launch("rans_ilv_encode_kernel");
  } else {
    dim3 grid((unsigned)((q.n_streams + kCoderWarps - 1) / kCoderWarps), (unsigned)N);
    rans_encode_kernel<<<grid, kCoderWarps * 32, 0, (cudaStream_t)stream>>>(p);
    rc = check_launch("rans_encode_kernel");
  }
  if (rc) return rc;
  PackP k;
  k.stream_words = p.stream_words; k.stream_data = p.stream_data;
  k.out = out; k.out_stride = out_stride_bytes;
  k.out_bytes = reinterpret_cast<long long*>(out_bytes);
  k.L = p.src.L; k.S = q.S; k.n_streams = q.n_streams; k.cap = q.cap; k.header = q.header;
  k.magic = lanes == 32 ? (skip_rows ? kMagic3S : kMagic3) : kMagic;
  k.magic_alt = kMagic3; k.flagged = flagged; k.flagged_max = flagged_max;
  // long sub-streams (few of them) are copied by several CTAs each
  const long long per_stream = (long long)q.S < 4096 ? 1 : ((long long)q.S + 4095) / 4096;
  const unsigned slices = (unsigned)(per_stream > 32 ? 32 : per_stream);
  rans_pack_kernel<<<dim3((unsigned)q.n_streams, (unsigned)N, slices), 128, 0, (cudaStream_t)stream>>>(k);
  return check_launch("rans_pack_kernel");
}

int dvc_rans_decode(const uint8_t* in, int64_t in_stride_bytes, const int64_t* in_bytes,
                    const int32_t* indexes, const float* scales, const float* scale_table,
                    int64_t T, float scale_bound, const int32_t* cdf, const int32_t* cdf_size,
                    const int32_t* offset, int64_t n_cdf, int64_t cdf_stride, const float* means,
                    float* out, int32_t* out_symbols, int* status, int64_t N, int64_t C,
                    int64_t H, int64_t W, const int64_t scales_st[4], const int64_t means_st[4],
                    const int64_t out_st[4], int64_t stream_symbols, int cb_parity,
                    int64_t cb_alt, int lanes, const uint8_t* skip_rows,
                    const uint16_t* cdf_pack, int64_t cdf_pack_entries, void* scratch,
                    dvc_stream_t stream) {
  DVC_REQUIRE(in && in_bytes, "rans_decode: null input");
  DVC_REQUIRE(!skip_rows || lanes == 32, "rans_decode: skip_rows needs the lane-interleaved layout");
  DVC_REQUIRE(cb_parity < 0 || (cb_parity <= 1 && scales),
              "rans_decode: cb_parity must be -1, or 0/1 together with scales");
  DVC_REQUIRE(out || out_symbols, "rans_decode: nothing to write");
  DVC_REQUIRE(!out || out_st, "rans_decode: out without strides");
  DVC_REQUIRE(N > 0 && N <= 65535, "rans_decode: N must be in [1, 65535]");
  DVC_REQUIRE((in_stride_bytes % 4) == 0 && (reinterpret_cast<uintptr_t>(in) & 3u) == 0,
              "rans_decode: in must be 4-byte aligned with a 4-byte multiple stride");
  DecP p;
  int rc = fill_source(p.src, nullptr, nullptr, means, indexes, scales, scale_table, T,
                       scale_bound, C, H, W, nullptr, means_st, scales_st, "rans_decode");
  if (rc) return rc;
  p.src.cb_parity = cb_parity < 0 ? -1 : cb_parity;
  p.src.cb_alt = cb_alt;
  rc = fill_tables(p.tb, cdf, cdf_size, offset, n_cdf, cdf_stride, "rans_decode");
  if (rc) return rc;
  Partition q;
  rc = make_partition(q, p.src.L, stream_symbols, lanes, "rans_decode");
  if (rc) return rc;
  DVC_REQUIRE(lanes == 1 || n_cdf <= 65535, "rans_decode: more than 65535 table rows");
  p.in = in; p.in_stride = in_stride_bytes;
  p.in_bytes = reinterpret_cast<const long long*>(in_bytes);
  p.out_f = out; p.out_sym = out_symbols; p.os = cts(out_st);
  p.S = q.S; p.n_streams = q.n_streams; p.header = q.header; p.N = (int)N;
  p.status = status;
  p.skip = skip_rows;
  p.pack = nullptr;
  p.pack_entries = 0;
  if (lanes == 32) {
    DVC_REQUIRE(n_cdf * cdf_stride < 2147483647LL, "rans_decode: CDF table too large");
    DVC_REQUIRE(scratch && (reinterpret_cast<uintptr_t>(scratch) & 15u) == 0,
                "rans_decode: the lane-interleaved layout needs 16-byte aligned scratch "
                "(dvc_rans_decode_scratch_bytes)");
    const int64_t cps = (p.src.L + kIlvChunk - 1) / kIlvChunk;
    uint8_t* sp = reinterpret_cast<uint8_t*>(scratch);
    IlvDecChunk* chunks = reinterpret_cast<IlvDecChunk*>(sp);
    sp += N * cps * (int64_t)sizeof(IlvDecChunk);
    uint16_t* pos_ci = reinterpret_cast<uint16_t*>(sp);
    sp += N * cps * kIlvChunk * 2;
    int32_t* syms = reinterpret_cast<int32_t*>(sp);
    // the packed tables are staged in shared memory when they fit
    if (cdf_pack && n_cdf <= kRowCache && cdf_pack_entries > 0 && cdf_pack_entries < (1 << 20) &&
        pack_layout((int)n_cdf, (int)cdf_pack_entries).bytes <= kPackMaxBytes) {
      DVC_REQUIRE((reinterpret_cast<uintptr_t>(cdf_pack) & 15u) == 0, "rans_decode: cdf_pack must be 16-byte aligned");
      p.pack = cdf_pack;
      p.pack_entries = (int)cdf_pack_entries;
    }
    const PackLayout pl = pack_layout((int)n_cdf, p.pack_entries);
    p.ilv_sym = out_symbols ? out_symbols : syms;
    IlvPrepP k;
    k.src = p.src; k.tb = p.tb; k.skip = skip_rows; k.enc = nullptr; k.dec = chunks; k.pos_ci = pos_ci;
    k.sym_out = p.ilv_sym;
    k.flagged = nullptr; k.flagged_max = 0;
    k.row_start = p.pack ? reinterpret_cast<const uint32_t*>(
                               reinterpret_cast<const uint8_t*>(p.pack) + pl.start_off) : nullptr;
    k.chunks_per_sample = (int)cps; k.status = status;
    ilv_prepare_kernel<false><<<(unsigned)(N * cps), kIlvChunk, 0, (cudaStream_t)stream>>>(k);
    rc = check_launch("ilv_prepare_kernel");
    if (rc) return rc;
    p.ilv_dec = chunks; p.ilv_ci = pos_ci; p.chunks_per_sample = (int)cps;
    const size_t smem = sizeof(IlvDecShared) + (p.pack ? (size_t)pl.bytes : 0);
    auto kern = p.pack ? rans_ilv_decode_kernel<true> : rans_ilv_decode_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess)
      return fail(DVC_ERR_CUDA, "rans_decode: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    kern<<<dim3((unsigned)q.n_streams, (unsigned)N), 32, smem, (cudaStream_t)stream>>>(p);
    rc = check_launch("rans_ilv_decode_kernel");
    if (rc || !out) return rc;
    IlvFinishP f;
    f.src = p.src; f.sym = p.ilv_sym; f.out_f = out; f.os = p.os; f.N = (int)N;
    const long long total = p.src.L * N;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    ilv_finish_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(f);
    return check_launch("ilv_finish_kernel");

  }
  dim3 grid((unsigned)((q.n_streams + kCoderWarps - 1) / kCoderWarps), (unsigned)N);
  rans_decode_kernel<<<grid, kCoderWarps * 32, 0, (cudaStream_t)stream>>>(p);
  return check_launch("rans_decode_kernel");
}

static int launch_dec_stage(DecStageP& p, int64_t N, int64_t C, int64_t H, int64_t W, bool stage_a,
                            dvc_stream_t stream, const char* who) {
  DVC_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, "%s: empty tensor", who);
  DVC_REQUIRE((C % 2) == 0 && (H % 2) == 0 && (W % 2) == 0, "%s: C, H, W must be even", who);
  DVC_REQUIRE((long long)C * H * W < 2147483647LL / 3, "%s: C*H*W too large", who);
  p.N = (int)N; p.C = (int)C; p.H = (int)H; p.W = (int)W;
  const long long total = (long long)N * C * H * W;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (stage_a) decode_stage_a_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  else decode_stage_b_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch(who);
}

int dvc_dual_prior_decode_stage_a(const int32_t* q0, const float* means, const float* scales,
                                  float* params, int64_t N, int64_t C, int64_t H, int64_t W,
                                  const int64_t means_st[4], const int64_t scales_st[4],
                                  const int64_t params_st[4], dvc_stream_t stream) {
  DVC_REQUIRE(q0 && means && scales && params && means_st && scales_st && params_st,
              "dual_prior_decode_stage_a: null pointer");
  DecStageP p;
  p.q0 = q0; p.q1 = nullptr; p.means = means; p.scales = scales; p.prior = nullptr; p.out = params;
  p.ms = cts(means_st); p.ss = cts(scales_st); p.ps = cts(nullptr); p.os = cts(params_st);
  return launch_dec_stage(p, N, C, H, W, true, stream, "dual_prior_decode_stage_a");
}

int dvc_dual_prior_decode_stage_b(const int32_t* q0, const int32_t* q1, const float* means,
                                  const float* prior, float* y_hat, int64_t N, int64_t C,
                                  int64_t H, int64_t W, const int64_t means_st[4],
                                  const int64_t prior_st[4], const int64_t y_hat_st[4],
                                  dvc_stream_t stream) {
  DVC_REQUIRE(q0 && q1 && means && prior && y_hat && means_st && prior_st && y_hat_st,
              "dual_prior_decode_stage_b: null pointer");
  DecStageP p;
  p.q0 = q0; p.q1 = q1; p.means = means; p.scales = nullptr; p.prior = prior; p.out = y_hat;
  p.ms = cts(means_st); p.ss = cts(nullptr); p.ps = cts(prior_st); p.os = cts(y_hat_st);
  return launch_dec_stage(p, N, C, H, W, false, stream, "dual_prior_decode_stage_b");
}

}  // extern "C"
