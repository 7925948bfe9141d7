/* synth_spec.h — counter-based synthetic input generators.
 *
 * Plain C, usable from host C/C++ and from CUDA device code.  Every value is a pure
 * function of (seed, stream, row, col), built only from exactly-rounded operations
 * (integer hash, int->float, mul, sub), so gcc -ffp-contract=off and nvcc -fmad=false
 * produce identical bits.  These are INPUT generators for tests and benches; they are
 * not part of the retrieval path and restate nothing from the reference except the
 * value range of MockEmbedder (reference src/embed.rs:136: uniform in [-1, 1)).
 */
#ifndef TRR_SYNTH_SPEC_H
#define TRR_SYNTH_SPEC_H

#include <stdint.h>

#if defined(__CUDACC__)
#define TRR_HD __host__ __device__ __forceinline__
#else
#define TRR_HD static inline
#endif

/* stream ids */
#define TRR_STREAM_CORPUS 0u
#define TRR_STREAM_QUERY 1u
#define TRR_STREAM_QNOISE 2u
#define TRR_STREAM_DOCTOK 3u
#define TRR_STREAM_DOCLEN 4u
#define TRR_STREAM_QTOK 5u
#define TRR_STREAM_QLEN 6u
#define TRR_STREAM_PLANT 7u
#define TRR_STREAM_DUP 8u

TRR_HD uint64_t trr_mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

TRR_HD uint64_t trr_hash4(uint64_t seed, uint32_t stream, uint64_t row, uint64_t col) {
  uint64_t h = trr_mix64(seed ^ ((uint64_t)stream * 0xD6E8FEB86659FD93ULL));
  h = trr_mix64(h + row);
  h = trr_mix64(h + col);
  return h;
}

/* uniform in [-1, 1) with 24 random bits; all three ops are exact in f32 */
TRR_HD float trr_uniform_pm1(uint64_t h) {
  float u = (float)(uint32_t)(h >> 40) * 5.9604644775390625e-08f; /* 2^-24 */
  return u * 2.0f - 1.0f;
}

/* float -> bf16 bits, round-to-nearest-even (finite inputs) */
TRR_HD uint16_t trr_f32_to_bf16_bits(float f) {
  union { float f; uint32_t u; } v;
  v.f = f;
  uint32_t lsb = (v.u >> 16) & 1u;
  v.u += 0x7FFFu + lsb;
  return (uint16_t)(v.u >> 16);
}

TRR_HD float trr_bf16_bits_to_f32(uint16_t b) {
  union { float f; uint32_t u; } v;
  v.u = ((uint32_t)b) << 16;
  return v.f;
}

/* clipped-Zipf sampling: cdf is a non-decreasing u64 table of n entries with
 * cdf[n-1] == UINT64_MAX; returns the first index i with cdf[i] >= h. */
TRR_HD uint32_t trr_cdf_lookup(const uint64_t* cdf, uint32_t n, uint64_t h) {
  uint32_t lo = 0, hi = n - 1;
  while (lo < hi) {
    uint32_t mid = lo + ((hi - lo) >> 1);
    if (cdf[mid] >= h) hi = mid; else lo = mid + 1;
  }
  return lo;
}

/* document length in tokens: U{32..64} */
TRR_HD uint32_t trr_doc_len(uint64_t seed, uint64_t doc) {
  return 32u + (uint32_t)(trr_hash4(seed, TRR_STREAM_DOCLEN, doc, 0) % 33u);
}

/* query length in terms: U{8..32} */
TRR_HD uint32_t trr_query_len(uint64_t seed, uint64_t q) {
  return 8u + (uint32_t)(trr_hash4(seed, TRR_STREAM_QLEN, q, 0) % 25u);
}

/* 1 % of queries are planted next to a corpus row; returns 1 and the row if planted */
TRR_HD int trr_query_planted(uint64_t seed, uint64_t q, uint64_t n_rows, uint64_t* row) {
  uint64_t h = trr_hash4(seed, TRR_STREAM_PLANT, q, 0);
  if (h % 100u != 0u || n_rows == 0) return 0;
  *row = (h >> 8) % n_rows;
  return 1;
}

/* tie-stress corpus: ~0.1 % of rows duplicate an earlier row's base content */
TRR_HD uint64_t trr_dup_source(uint64_t seed, uint64_t row, int enable) {
  if (!enable || row == 0) return row;
  uint64_t h = trr_hash4(seed, TRR_STREAM_DUP, row, 0);
  if (h % 1000u != 0u) return row;
  return (h >> 10) % row;
}

#endif /* TRR_SYNTH_SPEC_H */
