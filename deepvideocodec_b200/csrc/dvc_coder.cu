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
  c = (int)(e / s.HW);
  const int r = (int)(e - (long long)c * s.HW);
  h = r / s.W;
  w = r - h * s.W;
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
struct EncRec {                 // per-warp shared staging, structure of arrays
  unsigned long long rcp[kChunk];
  uint32_t bias[kChunk];        // start (+ 2^16 - 1 when freq == 1)
  uint32_t fs[kChunk];          // freq | rcp_shift << 17 | escape << 31
  uint32_t raw[kChunk];         // escape payload
};

struct EncP {
  Source src;
  Tables tb;
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
  int n_streams, cap, header;  // header = 1: DVC1 container, 0: raw stock stream
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
  if (j == 0 && threadIdx.x == 0) p.out_bytes[n] = fits ? need : -need;
  if (!fits) return;
  uint32_t* dst = reinterpret_cast<uint32_t*>(p.out + n * p.out_stride);
  const uint32_t mine = cnt[j];
  if (p.header) {
    if (j == 0 && threadIdx.x < 4) {
      const uint32_t hdr[4] = {kMagic, (uint32_t)p.L, (uint32_t)p.S, (uint32_t)p.n_streams};
      dst[threadIdx.x] = hdr[threadIdx.x];
    }
    if (threadIdx.x == 0) dst[4 + j] = mine;
  }
  const uint32_t* src = p.stream_data + ((long long)n * p.n_streams + j + 1) * p.cap - mine;
  uint32_t* d = dst + head + before;
  for (uint32_t i = threadIdx.x; i < mine; i += blockDim.x) d[i] = src[i];
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
static int make_partition(Partition& q, long long L, int64_t stream_symbols, const char* who) {
  DVC_REQUIRE(stream_symbols >= 0, "%s: stream_symbols must be >= 0", who);
  q.header = stream_symbols > 0 ? 1 : 0;
  q.S = q.header ? stream_symbols : L;
  const long long ns = (L + q.S - 1) / q.S;
  DVC_REQUIRE(ns >= 1 && ns <= (1 << 20), "%s: too many sub-streams (%lld)", who, ns);
  DVC_REQUIRE(2 * q.S + 4 < 2147483647LL, "%s: sub-stream too long", who);
  q.n_streams = (int)ns;
  q.cap = (int)(2 * q.S + 4);  // <= 52 bits per symbol + initial state + flush
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

int64_t dvc_rans_scratch_bytes(int64_t N, int64_t L, int64_t stream_symbols) {
  Partition q;
  if (N < 1 || L < 1 || make_partition(q, L, stream_symbols, "rans_scratch_bytes")) return -1;
  return N * (int64_t)q.n_streams * (int64_t)(q.cap + 1) * 4;
}

int64_t dvc_rans_max_bytes(int64_t L, int64_t stream_symbols) {
  Partition q;
  if (L < 1 || make_partition(q, L, stream_symbols, "rans_max_bytes")) return -1;
  return ((q.header ? 4 + (int64_t)q.n_streams : 0) + (int64_t)q.n_streams * q.cap) * 4;
}

int dvc_rans_encode(const float* x, const float* means, const int32_t* symbols,
                    const int32_t* indexes, const float* scales, const float* scale_table,
                    int64_t T, float scale_bound, const int32_t* cdf, const int32_t* cdf_size,
                    const int32_t* offset, int64_t n_cdf, int64_t cdf_stride, uint8_t* out,
                    int64_t out_stride_bytes, int64_t* out_bytes, void* scratch, int* status,
                    int64_t N, int64_t C, int64_t H, int64_t W, const int64_t x_st[4],
                    const int64_t means_st[4], const int64_t scales_st[4],
                    int64_t stream_symbols, dvc_stream_t stream) {
  DVC_REQUIRE((x != nullptr) != (symbols != nullptr),
              "rans_encode: give exactly one of x / symbols");
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
  rc = make_partition(q, p.src.L, stream_symbols, "rans_encode");
  if (rc) return rc;
  p.S = q.S; p.n_streams = q.n_streams; p.cap = q.cap; p.N = (int)N;
  p.stream_words = reinterpret_cast<uint32_t*>(scratch);
  p.stream_data = p.stream_words + N * (int64_t)q.n_streams;
  p.status = status;
  dim3 grid((unsigned)((q.n_streams + kCoderWarps - 1) / kCoderWarps), (unsigned)N);
  rans_encode_kernel<<<grid, kCoderWarps * 32, 0, (cudaStream_t)stream>>>(p);
  rc = check_launch("rans_encode_kernel");
  if (rc) return rc;
  PackP k;
  k.stream_words = p.stream_words; k.stream_data = p.stream_data;
  k.out = out; k.out_stride = out_stride_bytes;
  k.out_bytes = reinterpret_cast<long long*>(out_bytes);
  k.L = p.src.L; k.S = q.S; k.n_streams = q.n_streams; k.cap = q.cap; k.header = q.header;
  rans_pack_kernel<<<dim3((unsigned)q.n_streams, (unsigned)N), 128, 0, (cudaStream_t)stream>>>(k);
  return check_launch("rans_pack_kernel");
}

int dvc_rans_decode(const uint8_t* in, int64_t in_stride_bytes, const int64_t* in_bytes,
                    const int32_t* indexes, const float* scales, const float* scale_table,
                    int64_t T, float scale_bound, const int32_t* cdf, const int32_t* cdf_size,
                    const int32_t* offset, int64_t n_cdf, int64_t cdf_stride, const float* means,
                    float* out, int32_t* out_symbols, int* status, int64_t N, int64_t C,
                    int64_t H, int64_t W, const int64_t scales_st[4], const int64_t means_st[4],
                    const int64_t out_st[4], int64_t stream_symbols, int cb_parity,
                    int64_t cb_alt, dvc_stream_t stream) {
  DVC_REQUIRE(in && in_bytes, "rans_decode: null input");
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
  rc = make_partition(q, p.src.L, stream_symbols, "rans_decode");
  if (rc) return rc;
  p.in = in; p.in_stride = in_stride_bytes;
  p.in_bytes = reinterpret_cast<const long long*>(in_bytes);
  p.out_f = out; p.out_sym = out_symbols; p.os = cts(out_st);
  p.S = q.S; p.n_streams = q.n_streams; p.header = q.header; p.N = (int)N;
  p.status = status;
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
