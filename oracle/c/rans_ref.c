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

/* ------------------------------------------------------------------------ */
/* The lane-interleaved sub-stream of deepvideocodec_b200/csrc/dvc_coder.cu   */
/* ('DVC3' / 'DVS3' containers, include/dvc_b200.h), restated sequentially.   */
/*                                                                            */
/* This is the repository's OWN layout -- nothing in CompressAI corresponds   */
/* to it.  Every lane is a stock rans64 coder (the operations above); what is */
/* restated here is the schedule: which lane codes which item, and in which   */
/* order the lanes' renormalisation words appear in the shared stream.        */
/*   positions in chunks of 1024 = 32 groups of 32                            */
/*   pass 1 of a chunk: per group, the symbols of unmarked rows, then (if the */
/*     group has k >= 1 marked positions) the flag "a marked symbol != 0",    */
/*     coded as start/freq = (0, 65536-8k) or (65536-8k, 8k)                  */
/*   pass 2: the marked symbols of the flagged groups                         */
/*   item i of a pass -> lane i % 32, round i / 32                            */
/*   step t of a round = the t-th rans operation of every lane that has one,  */
/*     lanes ascending; decode order = chunks, pass 1, pass 2, rounds, steps  */
/*   stream = mask, 32 states (1 or 2 words; bit l of mask: 2), round words   */
/* ------------------------------------------------------------------------ */
#define ILV_LANES 32
#define ILV_CHUNK 1024

typedef struct {
  int32_t lane;
  sym_t op;
} lane_op_t;

typedef struct {
  lane_op_t* v;
  int64_t n, cap;
} opvec_t;

static void opvec_push(opvec_t* q, int32_t lane, sym_t op) {
  if (q->n == q->cap) {
    q->cap = q->cap ? 2 * q->cap : 1024;
    q->v = (lane_op_t*)realloc(q->v, sizeof(lane_op_t) * (size_t)q->cap);
  }
  q->v[q->n].lane = lane;
  q->v[q->n].op = op;
  q->n += 1;
}

/* the operations of one pass, appended in decode order.  items: position within the chunk, or
 * -(g + 1) for the flag of group g */
static void ilv_schedule_pass(opvec_t* q, const int32_t* items, int64_t n_items,
                              const int32_t* sym, const int32_t* idx, const int32_t* marked_in_group,
                              const uint8_t* flagged, const int32_t* cdfs, int64_t cdf_stride,
                              const int32_t* cdf_sizes, const int32_t* offsets) {
  sym_t ops[ILV_LANES][24];
  int64_t nops[ILV_LANES];
  for (int64_t r0 = 0; r0 < n_items; r0 += ILV_LANES) {
    const int64_t cnt = n_items - r0 < ILV_LANES ? n_items - r0 : ILV_LANES;
    int64_t tmax = 0;
    for (int64_t l = 0; l < cnt; ++l) {
      const int32_t it = items[r0 + l];
      if (it < 0) {
        const int32_t g = -it - 1;
        const uint32_t f1 = 8u * (uint32_t)marked_in_group[g];
        ops[l][0].bypass = 0;
        ops[l][0].start = (uint16_t)(flagged[g] ? 65536u - f1 : 0u);
        /* range can be 65536 - 8k < 65536: fits 16 bits because k >= 1 */
        ops[l][0].range = (uint16_t)(flagged[g] ? f1 : 65536u - f1);
        nops[l] = 1;
      } else {
        nops[l] = expand(sym[it], idx[it], cdfs, cdf_stride, cdf_sizes, offsets, ops[l]);
      }
      if (nops[l] > tmax) tmax = nops[l];
    }
    for (int64_t t = 0; t < tmax; ++t)
      for (int64_t l = 0; l < cnt; ++l)
        if (t < nops[l]) opvec_push(q, (int32_t)l, ops[l][t]);
  }
}

/* group analysis of one chunk; returns the number of pass-1 items */
static int64_t ilv_lists(const int32_t* idx, int64_t n_valid, const uint8_t* marks,
                         int32_t* marked_in_group, int32_t* items1) {
  int64_t n1 = 0;
  for (int32_t g = 0; g < ILV_CHUNK / ILV_LANES; ++g) {
    marked_in_group[g] = 0;
    for (int32_t l = 0; l < ILV_LANES; ++l) {
      const int64_t pos = (int64_t)g * ILV_LANES + l;
      if (pos >= n_valid) break;
      if (marks && marks[idx[pos]]) marked_in_group[g] += 1;
      else items1[n1++] = (int32_t)pos;
    }
    if (marked_in_group[g]) items1[n1++] = -(g + 1);
  }
  return n1;
}

/* Encodes symbols[0..n) into out[0..n_words); returns n_words, -1 if out_cap_words is too
 * small, -2 on a bad index. */
int64_t dvcref_ilv_encode(const int32_t* symbols, const int32_t* indexes, int64_t n,
                          const int32_t* cdfs, int64_t cdf_stride, const int32_t* cdf_sizes,
                          const int32_t* offsets, int32_t n_cdfs, const uint8_t* marks,
                          uint32_t* out, int64_t out_cap_words) {
  for (int64_t i = 0; i < n; ++i)
    if (indexes[i] < 0 || indexes[i] >= n_cdfs) return -2;
  opvec_t q = {NULL, 0, 0};
  int32_t items1[ILV_CHUNK + ILV_LANES], items2[ILV_CHUNK], marked[ILV_CHUNK / ILV_LANES];
  uint8_t flagged[ILV_CHUNK / ILV_LANES];
  for (int64_t cb = 0; cb < n; cb += ILV_CHUNK) {
    const int64_t n_valid = n - cb < ILV_CHUNK ? n - cb : ILV_CHUNK;
    const int32_t* sym = symbols + cb;
    const int32_t* idx = indexes + cb;
    const int64_t n1 = ilv_lists(idx, n_valid, marks, marked, items1);
    int64_t n2 = 0;
    for (int32_t g = 0; g < ILV_CHUNK / ILV_LANES; ++g) {
      flagged[g] = 0;
      if (!marked[g]) continue;
      for (int32_t l = 0; l < ILV_LANES && (int64_t)g * ILV_LANES + l < n_valid; ++l) {
        const int64_t pos = (int64_t)g * ILV_LANES + l;
        if (marks[idx[pos]] && sym[pos] != 0) flagged[g] = 1;
      }
      if (flagged[g])
        for (int32_t l = 0; l < ILV_LANES && (int64_t)g * ILV_LANES + l < n_valid; ++l) {
          const int64_t pos = (int64_t)g * ILV_LANES + l;
          if (marks[idx[pos]]) items2[n2++] = (int32_t)pos;
        }
    }
    ilv_schedule_pass(&q, items1, n1, sym, idx, marked, flagged, cdfs, cdf_stride, cdf_sizes, offsets);
    ilv_schedule_pass(&q, items2, n2, sym, idx, marked, flagged, cdfs, cdf_stride, cdf_sizes, offsets);
  }
  /* the coders run backwards over the schedule; words are written downwards */
  const int64_t n_words = q.n + 2 * ILV_LANES + 1;
  uint32_t* buf = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n_words);
  uint32_t* ptr = buf + n_words;
  uint64_t x[ILV_LANES];
  for (int l = 0; l < ILV_LANES; ++l) x[l] = RANS64_L;
  for (int64_t i = q.n - 1; i >= 0; --i) {
    const sym_t s = q.v[i].op;
    uint64_t* r = &x[q.v[i].lane];
    if (!s.bypass) {
      /* range == 0 encodes 65536 in 16 bits only for a flag with k = 0, which never exists */
      enc_put(r, &ptr, s.start, s.range, PRECISION);
    } else {
      enc_put_bits(r, &ptr, s.start, BYPASS_PRECISION);
    }
  }
  uint32_t mask = 0;
  for (int l = ILV_LANES - 1; l >= 0; --l) {
    if (x[l] >> 32) {
      mask |= 1u << l;
      *--ptr = (uint32_t)(x[l] >> 32);
    }
    *--ptr = (uint32_t)x[l];
  }
  *--ptr = mask;
  const int64_t used = (buf + n_words) - ptr;
  int64_t rc = used;
  if (used > out_cap_words) rc = -1;
  else memcpy(out, ptr, sizeof(uint32_t) * (size_t)used);
  free(buf);
  free(q.v);
  return rc;
}

typedef struct {
  const uint32_t* w;
  int64_t n, at;
  int overrun;
} wreader_t;

static uint32_t wr_next(wreader_t* r) {
  if (r->at >= r->n) {
    r->overrun = 1;
    r->at += 1;
    return 0;
  }
  return r->w[r->at++];
}

/* one pass of the decoder: lanes step in lock step, words taken in lane order */
static void ilv_decode_pass(wreader_t* rd, uint64_t* x, const int32_t* items, int64_t n_items,
                            const int32_t* idx, const int32_t* marked_in_group, uint8_t* flagged,
                            int32_t* val, const int32_t* cdfs, int64_t cdf_stride,
                            const int32_t* cdf_sizes, const int32_t* offsets) {
  for (int64_t r0 = 0; r0 < n_items; r0 += ILV_LANES) {
    const int64_t cnt = n_items - r0 < ILV_LANES ? n_items - r0 : ILV_LANES;
    int ph[ILV_LANES] = {0}, nb[ILV_LANES] = {0}, kk[ILV_LANES] = {0};
    uint32_t raw[ILV_LANES] = {0};
    int32_t sidx[ILV_LANES] = {0};
    for (int64_t l = 0; l < cnt; ++l) { /* step 0: the symbol / the flag */
      const int32_t it = items[r0 + l];
      const uint32_t cum = (uint32_t)(x[l] & 0xffffu);
      uint32_t start, freq;
      if (it < 0) {
        const uint32_t f0 = 65536u - 8u * (uint32_t)marked_in_group[-it - 1];
        const int f = cum >= f0;
        flagged[-it - 1] = (uint8_t)f;
        start = f ? f0 : 0u;
        freq = f ? 65536u - f0 : f0;
      } else {
        const int32_t ci = idx[it];
        const int32_t* cdf = cdfs + (int64_t)ci * cdf_stride;
        int32_t j = 0;
        while (j < cdf_sizes[ci] && !((uint32_t)cdf[j] > cum)) ++j;
        sidx[l] = j - 1;
        start = (uint32_t)cdf[j - 1];
        freq = (uint32_t)(cdf[j] - cdf[j - 1]);
        if (sidx[l] == cdf_sizes[ci] - 2) ph[l] = 1;
      }
      x[l] = (uint64_t)freq * (x[l] >> PRECISION) + cum - start;
      if (x[l] < RANS64_L) x[l] = (x[l] << 32) | wr_next(rd);
    }
    for (;;) { /* bypass nibbles */
      int any = 0;
      for (int64_t l = 0; l < cnt; ++l) {
        if (!ph[l]) continue;
        any = 1;
        const uint32_t v = (uint32_t)(x[l] & MAX_BYPASS_VAL);
        x[l] >>= BYPASS_PRECISION;
        if (x[l] < RANS64_L) x[l] = (x[l] << 32) | wr_next(rd);
        if (ph[l] == 1) {
          nb[l] += (int)v;
          if (v != MAX_BYPASS_VAL || nb[l] > 64) {
            if (nb[l] > 64) { nb[l] = 64; rd->overrun = 1; }
            ph[l] = nb[l] > 0 ? 2 : 0;
          }
        } else {
          if (kk[l] < 8) raw[l] |= v << (kk[l] * BYPASS_PRECISION);
          if (++kk[l] == nb[l]) ph[l] = 0;
        }
      }
      if (!any) break;
    }
    for (int64_t l = 0; l < cnt; ++l) {
      const int32_t it = items[r0 + l];
      if (it < 0) continue;
      const int32_t ci = idx[it];
      const int32_t max_value = cdf_sizes[ci] - 2;
      int32_t value = sidx[l];
      if (value == max_value) {
        value = (int32_t)(raw[l] >> 1);
        if (raw[l] & 1) value = -value - 1;
        else value += max_value;
      }
      val[it] = value + offsets[ci];
    }
  }
}

/* Decodes n symbols; returns the words consumed, -2 on a bad index, -3 if the stream is not a
 * well-formed sub-stream for these indexes (overrun, words left, lanes not back at 2^31). */
int64_t dvcref_ilv_decode(const uint32_t* words, int64_t n_words, const int32_t* indexes, int64_t n,
                          const int32_t* cdfs, int64_t cdf_stride, const int32_t* cdf_sizes,
                          const int32_t* offsets, int32_t n_cdfs, const uint8_t* marks,
                          int32_t* out) {
  for (int64_t i = 0; i < n; ++i)
    if (indexes[i] < 0 || indexes[i] >= n_cdfs) return -2;
  wreader_t rd = {words, n_words, 0, 0};
  const uint32_t mask = wr_next(&rd);
  uint64_t x[ILV_LANES];
  for (int l = 0; l < ILV_LANES; ++l) {
    x[l] = wr_next(&rd);
    if ((mask >> l) & 1u) x[l] |= (uint64_t)wr_next(&rd) << 32;
  }
  int32_t items1[ILV_CHUNK + ILV_LANES], items2[ILV_CHUNK], marked[ILV_CHUNK / ILV_LANES];
  uint8_t flagged[ILV_CHUNK / ILV_LANES];
  for (int64_t cb = 0; cb < n; cb += ILV_CHUNK) {
    const int64_t n_valid = n - cb < ILV_CHUNK ? n - cb : ILV_CHUNK;
    const int32_t* idx = indexes + cb;
    int32_t* val = out + cb;
    const int64_t n1 = ilv_lists(idx, n_valid, marks, marked, items1);
    memset(flagged, 0, sizeof(flagged));
    ilv_decode_pass(&rd, x, items1, n1, idx, marked, flagged, val, cdfs, cdf_stride, cdf_sizes, offsets);
    int64_t n2 = 0;
    for (int32_t g = 0; g < ILV_CHUNK / ILV_LANES; ++g) {
      if (!marked[g]) continue;
      for (int32_t l = 0; l < ILV_LANES && (int64_t)g * ILV_LANES + l < n_valid; ++l) {
        const int64_t pos = (int64_t)g * ILV_LANES + l;
        if (!marks[idx[pos]]) continue;
        if (flagged[g]) items2[n2++] = (int32_t)pos;
        else val[pos] = 0;
      }
    }
    ilv_decode_pass(&rd, x, items2, n2, idx, marked, flagged, val, cdfs, cdf_stride, cdf_sizes, offsets);
  }
  int ok = !rd.overrun && rd.at == n_words;
  for (int l = 0; l < ILV_LANES; ++l) ok = ok && x[l] == RANS64_L;
  return ok ? rd.at : -3;
}
