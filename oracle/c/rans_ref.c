/*
 * rans_ref.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C CPU restatement of the entropy-coding arithmetic the reference
 * reaches through CompressAI:
 *
 *   GaussianConditional.build_indexes     call sites dmc/models/video_model.py:248-249,
 *                                         272, 282, 422-423, 447, 457
 *   EntropyModel.compress / decompress    call sites video_model.py:238-239, 250-251,
 *                                         257, 273, 283, 411-412, 424-425, 431, 448, 458
 *   EntropyModel._pmf_to_cdf              via GaussianConditional.update_scale_table /
 *                                         EntropyBottleneck.update, video_model.py:669-677
 *
 * The arithmetic lives in the third-party package `compressai` (InterDigital
 * CompressAI, version pinned nowhere by the reference; >= 1.2 inferred,
 * SURVEY.md 8c), which is ABSENT from /root/reference and not installed:
 *   compressai/cpp_exts/rans/rans_interface.cpp   (BufferedRansEncoder, RansDecoder)
 *   compressai/cpp_exts/ops/ops.cpp               (pmf_to_quantized_cdf)
 *   third_party/ryg_rans/rans64.h                 (Rans64Enc*, Rans64Dec*, + PutBits/GetBits)
 * This file restates their PUBLISHED algorithm: 64-bit rANS state, lower bound
 * L = 2^31, 32-bit renormalisation words, 16-bit probability precision, 4-bit
 * bypass nibbles for out-of-range symbols.  PARITY UNPINNED: no CompressAI
 * build and no golden bitstream exist in the reference; what the tests can and
 * do pin is (i) encode -> decode round trips, (ii) hand-computed known-answer
 * streams for tiny inputs (tests/test_oracle_rans.py), (iii) the invariants of
 * a quantised CDF (strictly increasing, cdf[0] = 0, cdf[-1] = 2^16).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RANS64_L (1ull << 31)
#define PRECISION 16
#define BYPASS_PRECISION 4
#define MAX_BYPASS_VAL ((1 << BYPASS_PRECISION) - 1)

/* ------------------------------------------------------------------------ */
/* ops.cpp: pmf_to_quantized_cdf(pmf, precision) -> cdf[n + 1]               */
/* returns 0, or -1 (negative / non-finite pmf), -2 (all-zero pmf),          */
/* -3 (no frequency left to steal)                                           */
/* ------------------------------------------------------------------------ */
int dvcref_pmf_to_quantized_cdf(const float* pmf, int n, int precision, uint32_t* cdf) {
  for (int i = 0; i < n; ++i)
    if (pmf[i] < 0 || !isfinite(pmf[i])) return -1;
  cdf[0] = 0;
  for (int i = 0; i < n; ++i)
    cdf[i + 1] = (uint32_t)roundf(pmf[i] * (float)(1 << precision)); /* std::round(float) */
  uint32_t total = 0;
  for (int i = 0; i <= n; ++i) total += cdf[i];
  if (total == 0) return -2;
  for (int i = 0; i <= n; ++i)
    cdf[i] = (uint32_t)((((uint64_t)1 << precision) * cdf[i]) / total);
  for (int i = 1; i <= n; ++i) cdf[i] += cdf[i - 1]; /* partial_sum */
  cdf[n] = 1u << precision;
  for (int i = 0; i < n; ++i) {
    if (cdf[i] == cdf[i + 1]) {
      /* steal one count from the lowest-frequency symbol that has > 1 */
      uint32_t best_freq = ~0u;
      int best_steal = -1;
      for (int j = 0; j < n; ++j) {
        uint32_t freq = cdf[j + 1] - cdf[j];
        if (freq > 1 && freq < best_freq) {
          best_freq = freq;
          best_steal = j;
        }
      }
      if (best_steal == -1) return -3;
      if (best_steal < i) {
        for (int j = best_steal + 1; j <= i; ++j) cdf[j]--;
      } else {
        for (int j = i + 1; j <= best_steal; ++j) cdf[j]++;
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* GaussianConditional.build_indexes: s = max(scales, bound);                */
/* idx = (T-1) - sum_{k < T-1} (s <= table[k])                               */
/* ------------------------------------------------------------------------ */
void dvcref_build_indexes(const float* scales, int64_t n, const float* table, int T,
                          float bound, int32_t* out) {
  for (int64_t i = 0; i < n; ++i) {
    float s = scales[i];
    if (s < bound) s = bound; /* torch.max(x, bound): NaN stays NaN */
    int32_t idx = T - 1;
    for (int k = 0; k < T - 1; ++k) idx -= (s <= table[k]) ? 1 : 0;
    out[i] = idx;
  }
}

/* ------------------------------------------------------------------------ */
/* rans64.h                                                                  */
/* ------------------------------------------------------------------------ */
static void enc_put(uint64_t* r, uint32_t** pptr, uint32_t start, uint32_t freq,
                    uint32_t scale_bits) {
  uint64_t x = *r;
  uint64_t x_max = ((RANS64_L >> scale_bits) << 32) * freq;
  if (x >= x_max) {
    *pptr -= 1;
    **pptr = (uint32_t)x;
    x >>= 32;
  }
  *r = ((x / freq) << scale_bits) + (x % freq) + start;
}

static void enc_put_bits(uint64_t* r, uint32_t** pptr, uint32_t val, uint32_t nbits) {
  uint64_t x = *r;
  uint32_t freq = 1u << (16 - nbits);
  uint64_t x_max = ((RANS64_L >> 16) << 32) * freq;
  if (x >= x_max) {
    *pptr -= 1;
    **pptr = (uint32_t)x;
    x >>= 32;
  }
  *r = (x << nbits) | val;
}

static uint32_t dec_get_bits(uint64_t* r, const uint32_t** pptr, uint32_t nbits) {
  uint64_t x = *r;
  uint32_t val = (uint32_t)(x & ((1u << nbits) - 1));
  x >>= nbits;
  if (x < RANS64_L) {
    x = (x << 32) | **pptr;
    *pptr += 1;
  }
  *r = x;
  return val;
}

typedef struct {
  uint16_t start, range;
  uint8_t bypass;
} sym_t;

/* number of rANS operations symbol `s` coded with table `ci` expands to */
static int64_t expand(int32_t s, int32_t ci, const int32_t* cdfs, int64_t cdf_stride,
                      const int32_t* cdf_sizes, const int32_t* offsets, sym_t* out) {
  const int32_t* cdf = cdfs + (int64_t)ci * cdf_stride;
  const int32_t max_value = cdf_sizes[ci] - 2;
  int32_t value = s - offsets[ci];
  uint32_t raw_val = 0;
  int64_t k = 0;
  /* upstream holds raw_val in 32 bits: |symbol| beyond ~2^30 is outside its
   * domain (and its nibble-count loop would shift by >= 32); widen the
   * intermediate so the restatement is defined on the whole tested range */
  if (value < 0) {
    raw_val = (uint32_t)(-2 * (int64_t)value - 1);
    value = max_value;
  } else if (value >= max_value) {
    raw_val = (uint32_t)(2 * ((int64_t)value - max_value));
    value = max_value;
  }
  if (out) {
    out[k].start = (uint16_t)cdf[value];
    out[k].range = (uint16_t)(cdf[value + 1] - cdf[value]);
    out[k].bypass = 0;
  }
  ++k;
  if (value == max_value) {
    int32_t n_bypass = 0;
    while (n_bypass < 8 && (raw_val >> (n_bypass * BYPASS_PRECISION)) != 0) ++n_bypass;
    int32_t val = n_bypass;
    while (val >= MAX_BYPASS_VAL) {
      if (out) { out[k].start = MAX_BYPASS_VAL; out[k].range = MAX_BYPASS_VAL + 1; out[k].bypass = 1; }
      ++k;
      val -= MAX_BYPASS_VAL;
    }
    if (out) { out[k].start = (uint16_t)val; out[k].range = (uint16_t)(val + 1); out[k].bypass = 1; }
    ++k;
    for (int32_t j = 0; j < n_bypass; ++j) {
      const int32_t v = (raw_val >> (j * BYPASS_PRECISION)) & MAX_BYPASS_VAL;
      if (out) { out[k].start = (uint16_t)v; out[k].range = (uint16_t)(v + 1); out[k].bypass = 1; }
      ++k;
    }
  }
  return k;
}

/* rans_interface.cpp: RansEncoder::encode_with_indexes (= BufferedRansEncoder
 * encode_with_indexes + flush).  Returns the byte count written to `out`
 * (stream occupies out[0 .. nbytes)), -1 if out_cap is too small, -2 on a bad
 * index. */
int64_t dvcref_rans_encode_with_indexes(const int32_t* symbols, const int32_t* indexes,
                                        int64_t n, const int32_t* cdfs, int64_t cdf_stride,
                                        const int32_t* cdf_sizes, const int32_t* offsets,
                                        int32_t n_cdfs, uint8_t* out, int64_t out_cap) {
  int64_t n_syms = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (indexes[i] < 0 || indexes[i] >= n_cdfs) return -2;
    n_syms += expand(symbols[i], indexes[i], cdfs, cdf_stride, cdf_sizes, offsets, NULL);
  }
  sym_t* syms = (sym_t*)malloc(sizeof(sym_t) * (size_t)(n_syms > 0 ? n_syms : 1));
  int64_t k = 0;
  for (int64_t i = 0; i < n; ++i)
    k += expand(symbols[i], indexes[i], cdfs, cdf_stride, cdf_sizes, offsets, syms + k);
  const int64_t n_words = n_syms + 2; /* one word per op at most + the 2-word flush */
  uint32_t* buf = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n_words);
  uint32_t* ptr = buf + n_words;
  uint64_t rans = RANS64_L; /* Rans64EncInit */
  while (k > 0) {
    const sym_t s = syms[--k];
    if (!s.bypass) enc_put(&rans, &ptr, s.start, s.range, PRECISION);
    else enc_put_bits(&rans, &ptr, s.start, BYPASS_PRECISION);
  }
  ptr -= 2; /* Rans64EncFlush */
  ptr[0] = (uint32_t)(rans >> 0);
  ptr[1] = (uint32_t)(rans >> 32);
  const int64_t nbytes = (int64_t)((buf + n_words) - ptr) * 4;
  int64_t rc = nbytes;
  if (nbytes > out_cap) rc = -1;
  else memcpy(out, ptr, (size_t)nbytes);
  free(buf);
  free(syms);
  return rc;
}

/* rans_interface.cpp: RansDecoder::decode_with_indexes.  Returns the number of
 * bytes consumed, or -2 on a bad index. */
int64_t dvcref_rans_decode_with_indexes(const uint8_t* enc, int64_t nbytes,
                                        const int32_t* indexes, int64_t n, const int32_t* cdfs,
                                        int64_t cdf_stride, const int32_t* cdf_sizes,
                                        const int32_t* offsets, int32_t n_cdfs, int32_t* out) {
  (void)nbytes;
  const uint32_t* base = (const uint32_t*)enc;
  const uint32_t* ptr = base;
  uint64_t rans = (uint64_t)ptr[0] | ((uint64_t)ptr[1] << 32); /* Rans64DecInit */
  ptr += 2;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t ci = indexes[i];
    if (ci < 0 || ci >= n_cdfs) return -2;
    const int32_t* cdf = cdfs + (int64_t)ci * cdf_stride;
    const int32_t max_value = cdf_sizes[ci] - 2;
    const int32_t offset = offsets[ci];
    const uint32_t cum_freq = (uint32_t)(rans & ((1u << PRECISION) - 1)); /* Rans64DecGet */
    int32_t j = 0; /* std::find_if(cdf, cdf + size, v > cum_freq) */
    while (j < cdf_sizes[ci] && !((uint32_t)cdf[j] > cum_freq)) ++j;
    const uint32_t s = (uint32_t)(j - 1);
    { /* Rans64DecAdvance */
      const uint32_t start = (uint32_t)cdf[s], freq = (uint32_t)(cdf[s + 1] - cdf[s]);
      const uint64_t mask = (1ull << PRECISION) - 1;
      uint64_t x = rans;
      x = freq * (x >> PRECISION) + (x & mask) - start;
      if (x < RANS64_L) {
        x = (x << 32) | *ptr;
        ptr += 1;
      }
      rans = x;
    }
    int32_t value = (int32_t)s;
    if (value == max_value) {
      int32_t val = (int32_t)dec_get_bits(&rans, &ptr, BYPASS_PRECISION);
      int32_t n_bypass = val;
      while (val == MAX_BYPASS_VAL) {
        val = (int32_t)dec_get_bits(&rans, &ptr, BYPASS_PRECISION);
        n_bypass += val;
      }
      uint32_t raw_val = 0;
      for (int32_t b = 0; b < n_bypass; ++b) {
        val = (int32_t)dec_get_bits(&rans, &ptr, BYPASS_PRECISION);
        if (b < 8) raw_val |= (uint32_t)val << (b * BYPASS_PRECISION);
      }
      value = (int32_t)(raw_val >> 1);
      if (raw_val & 1) value = -value - 1;
      else value += max_value;
    }
    out[i] = value + offset;
  }
  return (int64_t)(ptr - base) * 4;
}
