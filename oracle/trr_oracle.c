/* trr_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the retrieval hot path of trueno-rag v0.1.8 (reference tree
 * /root/reference, Rust).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs may load this library; nothing under trueno_rag_b200/ links, imports or
 * calls it, and the product path has no CPU fallback.
 *
 * Parity status: PINNED by the reference's own in-tree known-answer tests (SURVEY.md §8c;
 * tests/test_oracle_kat.py replays each of them).  The reference cannot be compiled in this
 * image (no rustc/cargo), so there is no oracle/_ref; every arithmetic step on the path is
 * in-tree Rust std f32 arithmetic, restated here operation by operation:
 *   - f32 everywhere, sums are sequential folds starting from 0.0 in index order
 *     (Rust `iter().sum::<f32>()`), no FMA contraction: build with -ffp-contract=off;
 *   - `f32::ln` is glibc logf (same libm the Rust std would call on this platform);
 *   - the reference sorts with a stable sort over HashMap iteration order, which makes ties
 *     nondeterministic.  The canonical order used here and by the CUDA
 *     path is (score descending, insertion ordinal ascending), i.e. what a stable sort over
 *     an insertion-ordered Vec gives (reference crates/trueno-rag-cli/src/main.rs:480-492).
 *   - ChunkId (UUID) is replaced by the insertion ordinal (u32).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../trueno_rag_b200/csrc/synth_spec.h"

#define ORC_API __attribute__((visibility("default")))

enum { ORC_COSINE = 0, ORC_EUCLIDEAN = 1, ORC_DOT = 2 };
enum { ORC_RRF = 0, ORC_LINEAR = 1, ORC_CONVEX = 2, ORC_DBSF = 3, ORC_UNION = 4, ORC_INTERSECTION = 5 };

/* ------------------------------------------------------------------------------------------
 * Dense distance functions — reference src/index.rs:440-462
 * ---------------------------------------------------------------------------------------- */

/* src/index.rs:460-462  a.iter().zip(b).map(|(x,y)| x*y).sum() */
ORC_API float orc_dot(const float* a, const float* b, uint32_t d) {
  float s = 0.0f;
  for (uint32_t i = 0; i < d; ++i) s = s + a[i] * b[i];
  return s;
}

static float orc_sqnorm(const float* a, uint32_t d) {
  float s = 0.0f;
  for (uint32_t i = 0; i < d; ++i) s = s + a[i] * a[i];
  return s;
}

/* src/index.rs:440-450 */
ORC_API float orc_cosine(const float* a, const float* b, uint32_t d) {
  float dot = orc_dot(a, b, d);
  float norm_a = sqrtf(orc_sqnorm(a, d));
  float norm_b = sqrtf(orc_sqnorm(b, d));
  if (norm_a == 0.0f || norm_b == 0.0f) return 0.0f;
  return dot / (norm_a * norm_b);
}

/* src/index.rs:452-458   (x - y).powi(2) == (x-y)*(x-y) */
ORC_API float orc_euclidean(const float* a, const float* b, uint32_t d) {
  float s = 0.0f;
  for (uint32_t i = 0; i < d; ++i) {
    float t = a[i] - b[i];
    s = s + t * t;
  }
  return sqrtf(s);
}

/* src/index.rs:398-402 — the score used for ranking (higher is better) */
ORC_API float orc_dense_score(int metric, const float* q, const float* v, uint32_t d) {
  switch (metric) {
    case ORC_COSINE: return orc_cosine(q, v, d);
    case ORC_EUCLIDEAN: return -orc_euclidean(q, v, d);
    default: return orc_dot(q, v, d);
  }
}

/* ------------------------------------------------------------------------------------------
 * Canonical ordering
 * ---------------------------------------------------------------------------------------- */
typedef struct { float score; uint32_t ord; } orc_hit;

/* score desc (partial_cmp, b vs a), then ordinal asc.  Inputs are finite (documented). */
static int orc_hit_cmp(const void* pa, const void* pb) {
  const orc_hit* a = (const orc_hit*)pa;
  const orc_hit* b = (const orc_hit*)pb;
  if (a->score > b->score) return -1;
  if (a->score < b->score) return 1;
  if (a->ord < b->ord) return -1;
  if (a->ord > b->ord) return 1;
  return 0;
}

static int orc_hit_before(orc_hit a, orc_hit b) { return orc_hit_cmp(&a, &b) < 0; }

/* bounded selection equivalent to full sort + truncate(k): keeps `out` sorted canonically */
static void orc_topk_push(orc_hit* out, uint32_t* n, uint32_t k, orc_hit h) {
  if (k == 0) return;
  if (*n == k && !orc_hit_before(h, out[k - 1])) return;
  uint32_t pos = (*n < k) ? *n : k - 1;
  while (pos > 0 && orc_hit_before(h, out[pos - 1])) {
    out[pos] = out[pos - 1];
    --pos;
  }
  out[pos] = h;
  if (*n < k) ++*n;
}

/* ------------------------------------------------------------------------------------------
 * Dense search — reference src/index.rs:386-412 (score every stored vector, sort desc, truncate)
 *   rows: n x d row-major f32 (or bf16 bits widened to f32 when is_bf16);
 *   alive: optional n bytes, 0 = removed (VectorStore::remove, src/index.rs:421-424).
 *   literal != 0: score all, full sort, truncate (what the reference does);
 *   literal == 0: bounded selection, identical result.
 * ---------------------------------------------------------------------------------------- */
static void orc_load_row(const void* rows, int is_bf16, uint64_t i, uint32_t d, float* tmp, const float** out) {
  if (!is_bf16) {
    *out = (const float*)rows + i * d;
  } else {
    const uint16_t* r = (const uint16_t*)rows + i * d;
    for (uint32_t j = 0; j < d; ++j) tmp[j] = trr_bf16_bits_to_f32(r[j]);
    *out = tmp;
  }
}

ORC_API uint32_t orc_dense_search(int metric, const void* rows, int is_bf16, uint64_t n, uint32_t d,
                                  const uint8_t* alive, const float* q, uint32_t k, int literal,
                                  uint32_t* out_ord, float* out_score) {
  float* tmp = (float*)malloc(sizeof(float) * (d ? d : 1));
  uint32_t cnt = 0;
  if (literal) {
    orc_hit* all = (orc_hit*)malloc(sizeof(orc_hit) * (n ? n : 1));
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; ++i) {
      if (alive && !alive[i]) continue;
      const float* v;
      orc_load_row(rows, is_bf16, i, d, tmp, &v);
      all[m].score = orc_dense_score(metric, q, v, d);
      all[m].ord = (uint32_t)i;
      ++m;
    }
    qsort(all, m, sizeof(orc_hit), orc_hit_cmp);
    cnt = (uint32_t)(m < k ? m : k);
    for (uint32_t i = 0; i < cnt; ++i) { out_ord[i] = all[i].ord; out_score[i] = all[i].score; }
    free(all);
  } else {
    orc_hit* top = (orc_hit*)malloc(sizeof(orc_hit) * (k ? k : 1));
    for (uint64_t i = 0; i < n; ++i) {
      if (alive && !alive[i]) continue;
      const float* v;
      orc_load_row(rows, is_bf16, i, d, tmp, &v);
      orc_hit h = { orc_dense_score(metric, q, v, d), (uint32_t)i };
      orc_topk_push(top, &cnt, k, h);
    }
    for (uint32_t i = 0; i < cnt; ++i) { out_ord[i] = top[i].ord; out_score[i] = top[i].score; }
    free(top);
  }
  free(tmp);
  return cnt;
}

/* batch of B queries; `threads` > 1 parallelises over queries only (per-query arithmetic unchanged) */
ORC_API void orc_dense_search_batch(int metric, const void* rows, int is_bf16, uint64_t n, uint32_t d,
                                    const uint8_t* alive, const float* q, uint32_t B, uint32_t k, int literal,
                                    int threads, uint32_t* out_ord, float* out_score, uint32_t* out_n) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
  for (int64_t b = 0; b < (int64_t)B; ++b) {
    out_n[b] = orc_dense_search(metric, rows, is_bf16, n, d, alive, q + (uint64_t)b * d, k, literal,
                                out_ord + (uint64_t)b * k, out_score + (uint64_t)b * k);
  }
}

/* one query, documents split over `threads` workers (the "generous" all-cores CPU baseline,
 * BASELINE.md §4 B-par).  Per-document arithmetic is unchanged; partial top-k lists are merged
 * canonically, so the result is identical to orc_dense_search. */
ORC_API uint32_t orc_dense_search_par(int metric, const void* rows, int is_bf16, uint64_t n, uint32_t d,
                                      const float* q, uint32_t k, int threads, uint32_t* out_ord, float* out_score) {
  if (threads < 1) threads = 1;
  orc_hit* parts = (orc_hit*)malloc(sizeof(orc_hit) * (size_t)threads * (k ? k : 1));
  uint32_t* pn = (uint32_t*)calloc((size_t)threads, sizeof(uint32_t));
#pragma omp parallel num_threads(threads)
  {
    int t = 0;
#ifdef _OPENMP
    t = omp_get_thread_num();
#endif
    uint64_t lo = n * (uint64_t)t / (uint64_t)threads, hi = n * (uint64_t)(t + 1) / (uint64_t)threads;
    float* tmp = (float*)malloc(sizeof(float) * (d ? d : 1));
    orc_hit* top = parts + (size_t)t * k;
    uint32_t cnt = 0;
    for (uint64_t i = lo; i < hi; ++i) {
      const float* v;
      orc_load_row(rows, is_bf16, i, d, tmp, &v);
      orc_hit h = { orc_dense_score(metric, q, v, d), (uint32_t)i };
      orc_topk_push(top, &cnt, k, h);
    }
    pn[t] = cnt;
    free(tmp);
  }
  orc_hit* top = (orc_hit*)malloc(sizeof(orc_hit) * (k ? k : 1));
  uint32_t cnt = 0;
  for (int t = 0; t < threads; ++t)
    for (uint32_t i = 0; i < pn[t]; ++i) orc_topk_push(top, &cnt, k, parts[(size_t)t * k + i]);
  for (uint32_t i = 0; i < cnt; ++i) { out_ord[i] = top[i].ord; out_score[i] = top[i].score; }
  free(top); free(parts); free(pn);
  return cnt;
}

/* ------------------------------------------------------------------------------------------
 * BM25 — reference src/index.rs:127-243
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint32_t n_docs, n_terms;
  uint32_t n_stat;     /* N of the idf formula: n_docs, or the GLOBAL document count when this index is one shard of a
                          corpus scored with global statistics */
  uint64_t* term_off;  /* n_terms + 1 */
  uint32_t* post_doc;  /* postings sorted by doc within a term (insertion order, :193-196) */
  uint32_t* post_tf;
  uint32_t* doc_len;   /* token count after filtering (:178) */
  uint32_t* df;        /* :198-200 */
  float avgdl;         /* :157-164 (u32 sum) as f32 / doc_count as f32 */
  float k1, b;
  int borrowed;        /* arrays belong to the caller (orc_bm25_from_csr) */
} orc_bm25;

static int orc_u32_cmp(const void* a, const void* b) {
  uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

/* Build from tokenised documents (term ids), one `add` per document in ordinal order
 * (src/index.rs:176-204).  doc_off has n_docs+1 entries into tokens[]. */
ORC_API orc_bm25* orc_bm25_build(const uint64_t* doc_off, const uint32_t* tokens, uint32_t n_docs,
                                 uint32_t n_terms, float k1, float b) {
  orc_bm25* ix = (orc_bm25*)calloc(1, sizeof(orc_bm25));
  ix->n_docs = n_docs; ix->n_stat = n_docs; ix->n_terms = n_terms; ix->k1 = k1; ix->b = b;
  ix->term_off = (uint64_t*)calloc((size_t)n_terms + 1, sizeof(uint64_t));
  ix->doc_len = (uint32_t*)calloc(n_docs ? n_docs : 1, sizeof(uint32_t));
  ix->df = (uint32_t*)calloc(n_terms ? n_terms : 1, sizeof(uint32_t));
  uint32_t max_len = 0;
  for (uint32_t i = 0; i < n_docs; ++i) {
    uint64_t l = doc_off[i + 1] - doc_off[i];
    ix->doc_len[i] = (uint32_t)l;
    if (l > max_len) max_len = (uint32_t)l;
  }
  uint32_t* buf = (uint32_t*)malloc(sizeof(uint32_t) * (max_len ? max_len : 1));
  /* pass 1: df */
  for (uint32_t i = 0; i < n_docs; ++i) {
    uint32_t l = ix->doc_len[i];
    memcpy(buf, tokens + doc_off[i], sizeof(uint32_t) * l);
    qsort(buf, l, sizeof(uint32_t), orc_u32_cmp);
    for (uint32_t j = 0; j < l; ++j)
      if (j == 0 || buf[j] != buf[j - 1]) ix->df[buf[j]]++;
  }
  for (uint32_t t = 0; t < n_terms; ++t) ix->term_off[t + 1] = ix->term_off[t] + ix->df[t];
  uint64_t total = ix->term_off[n_terms];
  ix->post_doc = (uint32_t*)malloc(sizeof(uint32_t) * (total ? total : 1));
  ix->post_tf = (uint32_t*)malloc(sizeof(uint32_t) * (total ? total : 1));
  uint64_t* cur = (uint64_t*)malloc(sizeof(uint64_t) * (n_terms ? n_terms : 1));
  memcpy(cur, ix->term_off, sizeof(uint64_t) * n_terms);
  /* pass 2: postings, documents in ordinal order */
  for (uint32_t i = 0; i < n_docs; ++i) {
    uint32_t l = ix->doc_len[i];
    memcpy(buf, tokens + doc_off[i], sizeof(uint32_t) * l);
    qsort(buf, l, sizeof(uint32_t), orc_u32_cmp);
    uint32_t j = 0;
    while (j < l) {
      uint32_t e = j;
      while (e < l && buf[e] == buf[j]) ++e;
      uint64_t p = cur[buf[j]]++;
      ix->post_doc[p] = i;
      ix->post_tf[p] = e - j;
      j = e;
    }
  }
  /* :157-164 — u32 sum of all doc lengths, then as f32 / count as f32 */
  if (n_docs == 0) {
    ix->avgdl = 0.0f;
  } else {
    uint32_t sum = 0;
    for (uint32_t i = 0; i < n_docs; ++i) sum += ix->doc_len[i];
    ix->avgdl = (float)sum / (float)n_docs;
  }
  free(buf); free(cur);
  return ix;
}

/* Wraps caller-owned CSR arrays (the state orc_bm25_build would have produced) without copying them; used to check the
 * CUDA path at full corpus size, where N / df / avgdl are the GLOBAL statistics handed to the device index. */
ORC_API orc_bm25* orc_bm25_from_csr(uint32_t n_docs, uint32_t n_terms, uint64_t* term_off, uint32_t* post_doc,
                                    uint32_t* post_tf, uint32_t* doc_len, uint32_t* df, float avgdl, float k1, float b) {
  orc_bm25* ix = (orc_bm25*)calloc(1, sizeof(orc_bm25));
  ix->n_docs = n_docs; ix->n_stat = n_docs; ix->n_terms = n_terms; ix->k1 = k1; ix->b = b; ix->avgdl = avgdl;
  ix->term_off = term_off; ix->post_doc = post_doc; ix->post_tf = post_tf; ix->doc_len = doc_len; ix->df = df;
  ix->borrowed = 1;
  return ix;
}

ORC_API void orc_bm25_free(orc_bm25* ix) {
  if (!ix) return;
  if (ix->borrowed) { free(ix); return; }
  free(ix->term_off); free(ix->post_doc); free(ix->post_tf); free(ix->doc_len); free(ix->df); free(ix);
}

/* a shard of a larger corpus: N of the idf formula (df and avgdl are passed as global values to orc_bm25_from_csr) */
ORC_API void orc_bm25_set_stat_docs(orc_bm25* ix, uint32_t n_stat) { ix->n_stat = n_stat; }
ORC_API uint64_t orc_bm25_n_postings(const orc_bm25* ix) { return ix->term_off[ix->n_terms]; }
ORC_API float orc_bm25_avgdl(const orc_bm25* ix) { return ix->avgdl; }
ORC_API const uint64_t* orc_bm25_term_off(const orc_bm25* ix) { return ix->term_off; }
ORC_API const uint32_t* orc_bm25_post_doc(const orc_bm25* ix) { return ix->post_doc; }
ORC_API const uint32_t* orc_bm25_post_tf(const orc_bm25* ix) { return ix->post_tf; }
ORC_API const uint32_t* orc_bm25_doc_len(const orc_bm25* ix) { return ix->doc_len; }
ORC_API const uint32_t* orc_bm25_df(const orc_bm25* ix) { return ix->df; }

/* src/index.rs:147 */
ORC_API float orc_bm25_idf(uint32_t n_docs, uint32_t df_u) {
  float n = (float)n_docs, df = (float)df_u;
  return logf((n - df + 0.5f) / (df + 0.5f) + 1.0f);
}

/* src/index.rs:136-154, evaluation order exactly as written */
ORC_API float orc_bm25_score_term(uint32_t tf_u, uint32_t df_u, uint32_t n_docs, uint32_t doc_len_u,
                                  float avgdl, float k1, float b) {
  float tf = (float)tf_u;
  if (tf == 0.0f) return 0.0f;
  float doc_len = (float)doc_len_u;
  float idf = orc_bm25_idf(n_docs, df_u);
  float tf_norm = (tf * (k1 + 1.0f)) / (tf + k1 * (1.0f - b + b * doc_len / avgdl));
  return idf * tf_norm;
}

/* src/index.rs:127-133 — linear find in the term's posting list */
static uint32_t orc_bm25_term_frequency(const orc_bm25* ix, uint32_t term, uint32_t doc) {
  if (term >= ix->n_terms) return 0;
  for (uint64_t p = ix->term_off[term]; p < ix->term_off[term + 1]; ++p)
    if (ix->post_doc[p] == doc) return ix->post_tf[p];
  return 0;
}

/* LITERAL form of src/index.rs:212-243: candidate union, per candidate the sum over query terms in
 * query order (duplicates included) of score_term with its linear find; keep > 0.0; sort; truncate.
 * O(|cand| * sum df): only for small inputs.  Unknown terms have id >= n_terms. */
ORC_API uint32_t orc_bm25_search_literal(const orc_bm25* ix, const uint32_t* q_terms, uint32_t n_q, uint32_t k,
                                         uint32_t* out_ord, float* out_score) {
  if (n_q == 0) return 0;
  uint8_t* is_cand = (uint8_t*)calloc(ix->n_docs ? ix->n_docs : 1, 1);
  for (uint32_t t = 0; t < n_q; ++t) {
    uint32_t term = q_terms[t];
    if (term >= ix->n_terms) continue;
    for (uint64_t p = ix->term_off[term]; p < ix->term_off[term + 1]; ++p) is_cand[ix->post_doc[p]] = 1;
  }
  orc_hit* all = (orc_hit*)malloc(sizeof(orc_hit) * (ix->n_docs ? ix->n_docs : 1));
  uint32_t m = 0;
  for (uint32_t doc = 0; doc < ix->n_docs; ++doc) {
    if (!is_cand[doc]) continue;
    float score = 0.0f;
    for (uint32_t t = 0; t < n_q; ++t) {
      uint32_t term = q_terms[t];
      uint32_t tf = orc_bm25_term_frequency(ix, term, doc);
      uint32_t df = term < ix->n_terms ? ix->df[term] : 0;
      score = score + orc_bm25_score_term(tf, df, ix->n_stat, ix->doc_len[doc], ix->avgdl, ix->k1, ix->b);
    }
    if (score > 0.0f) { all[m].score = score; all[m].ord = doc; ++m; }
  }
  qsort(all, m, sizeof(orc_hit), orc_hit_cmp);
  uint32_t cnt = m < k ? m : k;
  for (uint32_t i = 0; i < cnt; ++i) { out_ord[i] = all[i].ord; out_score[i] = all[i].score; }
  free(all); free(is_cand);
  return cnt;
}

/* FAST-EQUIVALENT form: term-at-a-time accumulation in query-term order.
 * Adding 0.0 for a non-matching term is a no-op in f32, so the per-document sums are bit-identical
 * to the literal form (tests/test_oracle_kat.py proves it on random small indexes). */
ORC_API uint32_t orc_bm25_search(const orc_bm25* ix, const uint32_t* q_terms, uint32_t n_q, uint32_t k,
                                 uint32_t* out_ord, float* out_score) {
  if (n_q == 0) return 0;
  float* acc = (float*)calloc(ix->n_docs ? ix->n_docs : 1, sizeof(float));
  for (uint32_t t = 0; t < n_q; ++t) {
    uint32_t term = q_terms[t];
    if (term >= ix->n_terms) continue;
    uint32_t df = ix->df[term];
    for (uint64_t p = ix->term_off[term]; p < ix->term_off[term + 1]; ++p) {
      uint32_t doc = ix->post_doc[p];
      acc[doc] = acc[doc] + orc_bm25_score_term(ix->post_tf[p], df, ix->n_stat, ix->doc_len[doc], ix->avgdl, ix->k1, ix->b);
    }
  }
  orc_hit* top = (orc_hit*)malloc(sizeof(orc_hit) * (k ? k : 1));
  uint32_t cnt = 0;
  for (uint32_t doc = 0; doc < ix->n_docs; ++doc) {
    if (acc[doc] > 0.0f) {
      orc_hit h = { acc[doc], doc };
      orc_topk_push(top, &cnt, k, h);
    }
  }
  for (uint32_t i = 0; i < cnt; ++i) { out_ord[i] = top[i].ord; out_score[i] = top[i].score; }
  free(top); free(acc);
  return cnt;
}

ORC_API void orc_bm25_search_batch(const orc_bm25* ix, const uint32_t* q_terms, const uint32_t* q_off, uint32_t B,
                                   uint32_t k, int threads, uint32_t* out_ord, float* out_score, uint32_t* out_n) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
  for (int64_t b = 0; b < (int64_t)B; ++b) {
    out_n[b] = orc_bm25_search(ix, q_terms + q_off[b], q_off[b + 1] - q_off[b], k,
                               out_ord + (uint64_t)b * k, out_score + (uint64_t)b * k);
  }
}

/* ------------------------------------------------------------------------------------------
 * Fusion — reference src/fusion.rs:42-231.  The reference accumulates in a HashMap keyed by id;
 * here a first-appearance-ordered table plays that role (per-id accumulation order is identical:
 * dense list in order, then sparse list in order).
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint32_t id; float score; uint32_t rank; } orc_slot;

static int orc_find(const orc_slot* tab, uint32_t n, uint32_t id) {
  for (uint32_t i = 0; i < n; ++i) if (tab[i].id == id) return (int)i;
  return -1;
}

static float* orc_entry(orc_slot* tab, uint32_t* n, uint32_t id) { /* entry(id).or_insert(0.0) */
  int i = orc_find(tab, *n, id);
  if (i < 0) { i = (int)(*n)++; tab[i].id = id; tab[i].score = 0.0f; tab[i].rank = 0; }
  return &tab[i].score;
}

/* src/fusion.rs:183-202 */
static void orc_min_max_normalize(const float* s, uint32_t n, float* out) {
  if (n == 0) return;
  float mn = INFINITY, mx = -INFINITY;
  for (uint32_t i = 0; i < n; ++i) mn = fminf(mn, s[i]);
  for (uint32_t i = 0; i < n; ++i) mx = fmaxf(mx, s[i]);
  float range = mx - mn;
  if (fabsf(range) < 1.1920929e-07f) { /* f32::EPSILON */
    for (uint32_t i = 0; i < n; ++i) out[i] = 1.0f;
    return;
  }
  for (uint32_t i = 0; i < n; ++i) out[i] = (s[i] - mn) / range;
}

/* src/fusion.rs:205-224 */
static void orc_z_score_normalize(const float* s, uint32_t n, float* out) {
  if (n == 0) return;
  float nf = (float)n;
  float sum = 0.0f;
  for (uint32_t i = 0; i < n; ++i) sum = sum + s[i];
  float mean = sum / nf;
  float vs = 0.0f;
  for (uint32_t i = 0; i < n; ++i) { float t = s[i] - mean; vs = vs + t * t; }
  float variance = vs / nf;
  float std_dev = sqrtf(variance);
  if (fabsf(std_dev) < 1.1920929e-07f) {
    for (uint32_t i = 0; i < n; ++i) out[i] = 0.0f;
    return;
  }
  for (uint32_t i = 0; i < n; ++i) out[i] = (s[i] - mean) / std_dev;
}

ORC_API void orc_min_max(const float* s, uint32_t n, float* out) { orc_min_max_normalize(s, n, out); }
ORC_API void orc_z_score(const float* s, uint32_t n, float* out) { orc_z_score_normalize(s, n, out); }

static int orc_slot_rank_cmp(const void* pa, const void* pb) {
  const orc_slot* a = (const orc_slot*)pa; const orc_slot* b = (const orc_slot*)pb;
  return a->rank < b->rank ? -1 : (a->rank > b->rank ? 1 : 0);
}

static int orc_slot_score_cmp(const void* pa, const void* pb) {
  const orc_slot* a = (const orc_slot*)pa; const orc_slot* b = (const orc_slot*)pb;
  orc_hit ha = { a->score, a->id }, hb = { b->score, b->id };
  return orc_hit_cmp(&ha, &hb);
}

/* returns the fused length (<= nd + ns); out buffers must hold nd + ns entries */
ORC_API uint32_t orc_fuse(int strategy, float param, const uint32_t* d_id, const float* d_sc, uint32_t nd,
                          const uint32_t* s_id, const float* s_sc, uint32_t ns, uint32_t* out_id, float* out_sc) {
  uint32_t cap = nd + ns;
  orc_slot* tab = (orc_slot*)malloc(sizeof(orc_slot) * (cap ? cap : 1));
  float* dn = (float*)malloc(sizeof(float) * (nd ? nd : 1));
  float* sn = (float*)malloc(sizeof(float) * (ns ? ns : 1));
  uint32_t n = 0;
  int by_rank = 0;
  switch (strategy) {
    case ORC_RRF: { /* :68-84 */
      float k = param;
      for (uint32_t r = 0; r < nd; ++r) { float* e = orc_entry(tab, &n, d_id[r]); *e = *e + 1.0f / (k + (float)r + 1.0f); }
      for (uint32_t r = 0; r < ns; ++r) { float* e = orc_entry(tab, &n, s_id[r]); *e = *e + 1.0f / (k + (float)r + 1.0f); }
      break;
    }
    case ORC_LINEAR:
    case ORC_CONVEX: { /* :87-119 */
      float dense_weight = param;
      float sparse_weight = 1.0f - dense_weight;
      orc_min_max_normalize(d_sc, nd, dn);
      orc_min_max_normalize(s_sc, ns, sn);
      for (uint32_t r = 0; r < nd; ++r) { float* e = orc_entry(tab, &n, d_id[r]); *e = *e + dense_weight * dn[r]; }
      for (uint32_t r = 0; r < ns; ++r) { float* e = orc_entry(tab, &n, s_id[r]); *e = *e + sparse_weight * sn[r]; }
      break;
    }
    case ORC_DBSF: { /* :122-138 */
      orc_z_score_normalize(d_sc, nd, dn);
      orc_z_score_normalize(s_sc, ns, sn);
      for (uint32_t r = 0; r < nd; ++r) { float* e = orc_entry(tab, &n, d_id[r]); *e = *e + dn[r]; }
      for (uint32_t r = 0; r < ns; ++r) { float* e = orc_entry(tab, &n, s_id[r]); *e = *e + sn[r]; }
      break;
    }
    case ORC_UNION: { /* :141-160 — insert overwrites for dense; or_insert for sparse; sort by rank */
      for (uint32_t r = 0; r < nd; ++r) {
        int i = orc_find(tab, n, d_id[r]);
        if (i < 0) { i = (int)n++; tab[i].id = d_id[r]; }
        tab[i].score = d_sc[r]; tab[i].rank = r;
      }
      for (uint32_t r = 0; r < ns; ++r) {
        int i = orc_find(tab, n, s_id[r]);
        if (i < 0) { i = (int)n++; tab[i].id = s_id[r]; tab[i].score = s_sc[r]; tab[i].rank = nd + r; }
      }
      by_rank = 1;
      break;
    }
    default: { /* ORC_INTERSECTION :163-180 — maps built by collect(): last occurrence wins */
      for (uint32_t r = 0; r < nd; ++r) {
        if (orc_find(tab, n, d_id[r]) >= 0) continue; /* id already emitted */
        float dsc = 0.0f, ssc = 0.0f; int in_s = 0;
        for (uint32_t j = 0; j < nd; ++j) if (d_id[j] == d_id[r]) dsc = d_sc[j];
        for (uint32_t j = 0; j < ns; ++j) if (s_id[j] == d_id[r]) { ssc = s_sc[j]; in_s = 1; }
        if (in_s) { tab[n].id = d_id[r]; tab[n].score = (dsc + ssc) / 2.0f; tab[n].rank = 0; ++n; }
      }
      break;
    }
  }
  qsort(tab, n, sizeof(orc_slot), by_rank ? orc_slot_rank_cmp : orc_slot_score_cmp); /* :227-231 */
  for (uint32_t i = 0; i < n; ++i) { out_id[i] = tab[i].id; out_sc[i] = tab[i].score; }
  free(tab); free(dn); free(sn);
  return n;
}

/* ------------------------------------------------------------------------------------------
 * Hybrid retrieve — reference src/retrieve.rs:175-220: fuse the two top-C lists, take(k), attach
 * dense/sparse scores when the id was in that source's list (maps built by collect(): last wins).
 * Absent scores are NaN.  `alive` (optional, indexed by ordinal) models `self.dense.get(id)` failing
 * for ids missing from the dense store (:205) — such ids are skipped but still consume take(k).
 * ---------------------------------------------------------------------------------------- */
ORC_API uint32_t orc_hybrid_assemble(int strategy, float param, const uint32_t* d_id, const float* d_sc, uint32_t nd,
                                     const uint32_t* s_id, const float* s_sc, uint32_t ns, uint32_t k,
                                     const uint8_t* alive, uint32_t* out_id, float* out_fused, float* out_dense,
                                     float* out_sparse) {
  uint32_t cap = nd + ns;
  uint32_t* fid = (uint32_t*)malloc(sizeof(uint32_t) * (cap ? cap : 1));
  float* fsc = (float*)malloc(sizeof(float) * (cap ? cap : 1));
  uint32_t nf = orc_fuse(strategy, param, d_id, d_sc, nd, s_id, s_sc, ns, fid, fsc);
  uint32_t take = nf < k ? nf : k, m = 0;
  for (uint32_t i = 0; i < take; ++i) {
    uint32_t id = fid[i];
    if (alive && !alive[id]) continue;
    float ds = NAN, ss = NAN;
    for (uint32_t j = 0; j < nd; ++j) if (d_id[j] == id) ds = d_sc[j];
    for (uint32_t j = 0; j < ns; ++j) if (s_id[j] == id) ss = s_sc[j];
    out_id[m] = id; out_fused[m] = fsc[i]; out_dense[m] = ds; out_sparse[m] = ss; ++m;
  }
  free(fid); free(fsc);
  return m;
}

/* ------------------------------------------------------------------------------------------
 * Synthetic inputs (shared spec with the device generators, synth_spec.h) — not reference code.
 * ---------------------------------------------------------------------------------------- */

/* one embedding row: uniform [-1,1), L2-normalised with the sequential f32 sum, optional bf16 rounding.
 * planted (query streams): row = corpus_row + 0.1 * noise, renormalised. */
static void orc_synth_base_row(uint64_t seed, uint32_t stream, uint64_t row, uint32_t d, float* out) {
  for (uint32_t j = 0; j < d; ++j) out[j] = trr_uniform_pm1(trr_hash4(seed, stream, row, j));
}

static void orc_normalize_row(float* x, uint32_t d) {
  float s = 0.0f;
  for (uint32_t j = 0; j < d; ++j) s = s + x[j] * x[j];
  float nrm = sqrtf(s);
  if (nrm > 0.0f) for (uint32_t j = 0; j < d; ++j) x[j] = x[j] / nrm;
}

ORC_API void orc_synth_corpus_rows(uint64_t seed, uint64_t row0, uint64_t n, uint32_t d, int to_bf16, int dups,
                                   float* out_f32, uint16_t* out_bf16) {
#pragma omp parallel
  {
    float* tmp = (float*)malloc(sizeof(float) * d);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
      uint64_t row = row0 + (uint64_t)i;
      uint64_t src = trr_dup_source(seed, row, dups);
      orc_synth_base_row(seed, TRR_STREAM_CORPUS, src, d, tmp);
      orc_normalize_row(tmp, d);
      if (to_bf16) {
        for (uint32_t j = 0; j < d; ++j) {
          uint16_t b = trr_f32_to_bf16_bits(tmp[j]);
          if (out_bf16) out_bf16[(uint64_t)i * d + j] = b;
          if (out_f32) out_f32[(uint64_t)i * d + j] = trr_bf16_bits_to_f32(b);
        }
      } else {
        memcpy(out_f32 + (uint64_t)i * d, tmp, sizeof(float) * d);
      }
    }
    free(tmp);
  }
}

ORC_API void orc_synth_queries(uint64_t seed, uint64_t q0, uint64_t n, uint32_t d, uint64_t n_corpus, int corpus_bf16,
                               int dups, float* out) {
  float* base = (float*)malloc(sizeof(float) * d);
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t q = q0 + i, row;
    float* x = out + i * d;
    if (trr_query_planted(seed, q, n_corpus, &row)) {
      orc_synth_corpus_rows(seed, row, 1, d, corpus_bf16, dups, base, NULL);
      for (uint32_t j = 0; j < d; ++j) {
        float nz = trr_uniform_pm1(trr_hash4(seed, TRR_STREAM_QNOISE, q, j));
        x[j] = base[j] + 0.1f * nz;
      }
    } else {
      orc_synth_base_row(seed, TRR_STREAM_QUERY, q, d, x);
    }
    orc_normalize_row(x, d);
  }
  free(base);
}

/* tokens of documents [doc0, doc0+n): returns total token count; doc_off (n+1) is relative to out_tokens */
ORC_API uint64_t orc_synth_doc_tokens(uint64_t seed, const uint64_t* cdf, uint32_t n_terms, uint64_t doc0, uint64_t n,
                                      uint64_t* doc_off, uint32_t* out_tokens) {
  doc_off[0] = 0;
  for (uint64_t i = 0; i < n; ++i) doc_off[i + 1] = doc_off[i] + trr_doc_len(seed, doc0 + i);
  if (out_tokens) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
      uint32_t l = (uint32_t)(doc_off[i + 1] - doc_off[i]);
      for (uint32_t t = 0; t < l; ++t)
        out_tokens[doc_off[i] + t] = trr_cdf_lookup(cdf, n_terms, trr_hash4(seed, TRR_STREAM_DOCTOK, doc0 + (uint64_t)i, t));
    }
  }
  return doc_off[n];
}

ORC_API uint64_t orc_synth_query_terms(uint64_t seed, const uint64_t* cdf, uint32_t n_terms, uint64_t q0, uint64_t n,
                                       uint32_t* q_off, uint32_t* out_terms) {
  q_off[0] = 0;
  for (uint64_t i = 0; i < n; ++i) q_off[i + 1] = q_off[i] + trr_query_len(seed, q0 + i);
  if (out_terms) {
    for (uint64_t i = 0; i < n; ++i) {
      uint32_t l = q_off[i + 1] - q_off[i];
      for (uint32_t t = 0; t < l; ++t)
        out_terms[q_off[i] + t] = trr_cdf_lookup(cdf, n_terms, trr_hash4(seed, TRR_STREAM_QTOK, q0 + i, t));
    }
  }
  return q_off[n];
}
