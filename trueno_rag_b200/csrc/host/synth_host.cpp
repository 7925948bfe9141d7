// synth_host.cpp — host-side generators of the synthetic hybrid workload (tests / benches only; not reference
// behaviour).  Same counter-based recipe as the device generator (csrc/synth_spec.h): queries,
// query terms, and the BM25 postings of a document shard as a CSR with local doc ids.  OpenMP over documents.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../../include/trueno_rag_b200.h"
#include "../synth_spec.h"

int trr_fail(int status, const std::string& msg);

static void base_row(uint64_t seed, uint32_t stream, uint64_t row, uint32_t d, float* out) {
  for (uint32_t j = 0; j < d; ++j) out[j] = trr_uniform_pm1(trr_hash4(seed, stream, row, j));
}
static void normalize_row(float* x, uint32_t d) {
  float s = 0.0f;
  for (uint32_t j = 0; j < d; ++j) s = s + x[j] * x[j];
  const float nrm = sqrtf(s);
  if (nrm > 0.0f) for (uint32_t j = 0; j < d; ++j) x[j] = x[j] / nrm;
}

extern "C" int trr_synth_queries(uint64_t seed, uint64_t q0, uint64_t n, uint32_t dim, uint64_t n_corpus, int corpus_bf16,
                                 int dups, int round_to_bf16, float* out) {
  if (!out && n) return trr_fail(TRR_ERR_INVALID_ARG, "trr_synth_queries: NULL output");
  std::vector<float> base(dim);
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t q = q0 + i;
    uint64_t row;
    float* x = out + i * dim;
    if (trr_query_planted(seed, q, n_corpus, &row)) {
      base_row(seed, TRR_STREAM_CORPUS, trr_dup_source(seed, row, dups), dim, base.data());
      normalize_row(base.data(), dim);
      if (corpus_bf16) for (uint32_t j = 0; j < dim; ++j) base[j] = trr_bf16_bits_to_f32(trr_f32_to_bf16_bits(base[j]));
      for (uint32_t j = 0; j < dim; ++j) x[j] = base[j] + 0.1f * trr_uniform_pm1(trr_hash4(seed, TRR_STREAM_QNOISE, q, j));
    } else {
      base_row(seed, TRR_STREAM_QUERY, q, dim, x);
    }
    normalize_row(x, dim);
    if (round_to_bf16) for (uint32_t j = 0; j < dim; ++j) x[j] = trr_bf16_bits_to_f32(trr_f32_to_bf16_bits(x[j]));
  }
  return TRR_OK;
}

extern "C" int trr_synth_query_terms(uint64_t seed, const uint64_t* cdf, uint32_t n_terms, uint64_t q0, uint64_t n,
                                     uint32_t* q_off, uint32_t* out_terms, uint64_t out_cap) {
  if (!cdf || !q_off) return trr_fail(TRR_ERR_INVALID_ARG, "trr_synth_query_terms: NULL argument");
  q_off[0] = 0;
  for (uint64_t i = 0; i < n; ++i) q_off[i + 1] = q_off[i] + trr_query_len(seed, q0 + i);
  if (!out_terms) return TRR_OK;
  if (out_cap < q_off[n]) return trr_fail(TRR_ERR_INVALID_ARG, "trr_synth_query_terms: output too small");
  for (uint64_t i = 0; i < n; ++i)
    for (uint32_t t = 0; t < q_off[i + 1] - q_off[i]; ++t)
      out_terms[q_off[i] + t] = trr_cdf_lookup(cdf, n_terms, trr_hash4(seed, TRR_STREAM_QTOK, q0 + i, t));
  return TRR_OK;
}

// distinct terms of one document with their frequencies (tokens sorted by term id)
static inline uint32_t doc_terms(uint64_t seed, const uint64_t* cdf, uint32_t n_terms, uint64_t doc, uint32_t* terms,
                                 uint32_t* tfs, uint32_t* len_out) {
  uint32_t buf[64];
  const uint32_t len = trr_doc_len(seed, doc);
  for (uint32_t t = 0; t < len; ++t) buf[t] = trr_cdf_lookup(cdf, n_terms, trr_hash4(seed, TRR_STREAM_DOCTOK, doc, t));
  std::sort(buf, buf + len);
  uint32_t m = 0;
  for (uint32_t a = 0; a < len;) {
    uint32_t e = a;
    while (e < len && buf[e] == buf[a]) ++e;
    terms[m] = buf[a];
    tfs[m] = e - a;
    ++m;
    a = e;
  }
  *len_out = len;
  return m;
}

extern "C" int trr_synth_bm25_count(uint64_t seed, const uint64_t* cdf, uint32_t n_terms, uint64_t doc_lo, uint64_t doc_hi,
                                    uint32_t* df_local, uint32_t* doc_len, uint64_t* total_len) {
  if (!cdf || !df_local || !doc_len || !total_len) return trr_fail(TRR_ERR_INVALID_ARG, "trr_synth_bm25_count: NULL argument");
  memset(df_local, 0, sizeof(uint32_t) * n_terms);
  uint64_t total = 0;
  const int64_t n = (int64_t)(doc_hi - doc_lo);
#pragma omp parallel for schedule(static) reduction(+ : total)
  for (int64_t i = 0; i < n; ++i) {
    uint32_t terms[64], tfs[64], len;
    const uint32_t m = doc_terms(seed, cdf, n_terms, doc_lo + (uint64_t)i, terms, tfs, &len);
    doc_len[i] = len;
    total += len;
    for (uint32_t j = 0; j < m; ++j) {
#pragma omp atomic
      df_local[terms[j]] += 1;
    }
  }
  *total_len = total;
  return TRR_OK;
}

extern "C" int trr_synth_bm25_fill(uint64_t seed, const uint64_t* cdf, uint32_t n_terms, uint64_t doc_lo, uint64_t doc_hi,
                                   const uint64_t* term_off, uint32_t* post_doc, uint32_t* post_tf) {
  if (!cdf || !term_off || !post_doc || !post_tf) return trr_fail(TRR_ERR_INVALID_ARG, "trr_synth_bm25_fill: NULL argument");
  const int64_t n = (int64_t)(doc_hi - doc_lo);
  int n_threads = 1;
#ifdef _OPENMP
  n_threads = omp_get_max_threads();
#endif
  // postings of a term must be in document order: thread t owns a contiguous block of documents, and its postings of
  // a term start after those of the threads before it.  counts[t][term] -> per-thread cursors.
  std::vector<std::vector<uint32_t>> counts((size_t)n_threads, std::vector<uint32_t>(n_terms, 0));
#pragma omp parallel num_threads(n_threads)
  {
    int t = 0;
#ifdef _OPENMP
    t = omp_get_thread_num();
#endif
    const int64_t lo = n * t / n_threads, hi = n * (t + 1) / n_threads;
    uint32_t terms[64], tfs[64], len;
    std::vector<uint32_t>& cnt = counts[(size_t)t];
    for (int64_t i = lo; i < hi; ++i) {
      const uint32_t m = doc_terms(seed, cdf, n_terms, doc_lo + (uint64_t)i, terms, tfs, &len);
      for (uint32_t j = 0; j < m; ++j) cnt[terms[j]] += 1;
    }
  }
  // exclusive scan over threads, per term (parallel over terms)
#pragma omp parallel for schedule(static)
  for (int64_t term = 0; term < (int64_t)n_terms; ++term) {
    uint32_t run = 0;
    for (int t = 0; t < n_threads; ++t) {
      const uint32_t c = counts[(size_t)t][(size_t)term];
      counts[(size_t)t][(size_t)term] = run;
      run += c;
    }
  }
#pragma omp parallel num_threads(n_threads)
  {
    int t = 0;
#ifdef _OPENMP
    t = omp_get_thread_num();
#endif
    const int64_t lo = n * t / n_threads, hi = n * (t + 1) / n_threads;
    uint32_t terms[64], tfs[64], len;
    std::vector<uint32_t>& cur = counts[(size_t)t];
    for (int64_t i = lo; i < hi; ++i) {
      const uint32_t m = doc_terms(seed, cdf, n_terms, doc_lo + (uint64_t)i, terms, tfs, &len);
      for (uint32_t j = 0; j < m; ++j) {
        const uint64_t p = term_off[terms[j]] + cur[terms[j]]++;
        post_doc[p] = (uint32_t)i;
        post_tf[p] = tfs[j];
      }
    }
  }
  return TRR_OK;
}
