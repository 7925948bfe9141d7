// bm25.cuh — argument blocks of the BM25 kernels (bm25.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int TRR_BM25_CONSUMER_WARPS = 16;
constexpr int TRR_BM25_THREADS = (TRR_BM25_CONSUMER_WARPS + 1) * 32;  // 16 consumer warps + 1 producer warp
constexpr uint32_t TRR_BM25_MAX_QUERY_TERMS = 512;
constexpr uint32_t TRR_BM25_MIN_RANGE_SHIFT = 11, TRR_BM25_MAX_RANGE_SHIFT = 15;

struct Bm25BuildArgs {
  uint64_t n_postings;
  uint32_t n_terms, n_docs;
  const uint64_t* term_off;  // n_terms + 1
  const uint32_t* post_doc;  // nullable: document ids are then read from post[].x
  const uint32_t* post_tf;
  const uint32_t* doc_len;
  const float* idf;
  float avgdl, k1, b;
  uint32_t range_shift, n_ranges, skip_ld;
  uint2* post;     // out: {doc, impact bits}
  uint32_t* skip;  // out: [n_terms][skip_ld]
  uint32_t* term_min;  // out: [n_terms] smallest impact of the term (order-preserving u32 image), pre-set to 0xFFFFFFFF
  uint32_t* term_max;  // out: [n_terms] largest impact of the term (order-preserving u32 image), pre-set to 0
  uint32_t* flags;     // out: [0] bit 0 = some live impact is not a finite value > 0 (disables the threshold bootstrap and the
                       //      integer fast pass), pre-set to 0; the host ORs bit 1 in when dead postings exist (no bootstrap)
};

struct Bm25MergeArgs {  // trr_bm25_append: old CSR + CSR of the appended documents -> new CSR (doc ids and tf only)
  uint64_t n_postings_new;
  uint32_t n_terms_new, n_terms_old, n_docs_old;
  const uint64_t* new_off;    // n_terms_new + 1
  const uint64_t* old_off;    // n_terms_old + 1
  const uint64_t* delta_off;  // n_terms_new + 1
  const uint2* old_post;
  const uint32_t* old_tf;
  const uint32_t* delta_doc;  // ids relative to the first appended document
  const uint32_t* delta_tf;
  uint2* new_post;
  uint32_t* new_tf;
};

struct Bm25SearchArgs {
  const uint2* post;
  const uint32_t* skip;
  uint32_t skip_ld, n_terms, n_docs, n_ranges, range_shift, doc_base;
  const uint32_t* q_terms;
  const uint32_t* q_off;
  uint32_t B, k;
  uint32_t stage_cap;  // postings per stage buffer (even)
  uint32_t cand_cap;   // candidate buffer capacity: power of two > k
  uint32_t n_chunks;   // work items per query (contiguous chunks of document ranges)
  const uint32_t* term_min;  // [n_terms] smallest impact per term (orderable image)
  const uint32_t* term_max;  // [n_terms] largest impact per term (orderable image)
  const uint32_t* flags;     // [0] != 0: no threshold bootstrap (see Bm25BuildArgs::flags)
  uint64_t* thr0;      // [B] initial threshold key per query for the exact kernel (written by the plan kernel; 0 = none)
  uint32_t* order;     // [B] queries by decreasing posting volume (written by the plan kernel); the exact kernel run as the
                       //     fallback of the fast pass gets the list of flagged queries here
  const uint32_t* n_sel_ptr;  // nullable: DEVICE count of entries of `order` to process (else B)
  uint32_t* queue;     // [1] dynamic work queue (reset by the plan kernel)
  uint64_t* partial;   // n_chunks > 1: [position in `order`][n_chunks][k] keys, merged by topk_merge_kernel
  uint64_t* out_keys;  // n_chunks == 1: nullable [B][k]
  uint32_t* out_ord;   // nullable [B][k]
  float* out_score;    // nullable
  uint32_t* out_n;     // nullable
  uint32_t* dbg;       // nullable host-mapped word: site of a barrier timeout
  // integer fast pass (bm25_fast_kernel): k = kf candidates per query by fast score
  uint32_t kf;         // candidates kept per query by the fast pass (> the k of the search)
  uint32_t rmul_shift; // the fast pass walks 2^rmul_shift ranges of the skip table per accumulator pass
  uint32_t n_acc;      // accumulators in the fast pass's ring (2: accumulate range i + 1 while range i is scanned)
  const float* qscale;     // [B] scale of the query's fixed-point scores for the cell width of this launch
  const uint32_t* thr0f;   // [B] initial fixed-point threshold for the cell width of this launch (0 = none)
  float* qscale16;         // plan kernel outputs: [B] scale / bootstrap threshold for 16-bit cells ...
  uint32_t* thr0f16;
  float* qscale32;         // ... and for 32-bit cells
  uint32_t* thr0f32;
  uint64_t* fast_keys; // [position in `order`][n_chunks][kf] keys (fixed-point score << 32 | ~ordinal), descending, 0 = empty
  uint32_t triage;     // -DTRR_TRIAGE builds only (timing experiments that break the results); 0 otherwise
  uint32_t* triage_out; // -DTRR_TRIAGE builds only: per-warp counters
};

// exact re-scoring of the fast pass's candidates + proof (bm25_rescore_kernel)
struct Bm25RescoreArgs {
  const uint2* post;
  const uint32_t* skip;
  uint32_t skip_ld, n_terms, range_shift, doc_base;
  const uint32_t* q_terms;
  const uint32_t* q_off;
  uint32_t B, k, kf, cap2;     // cap2 = power of two >= kf
  const uint32_t* sel;         // CTA i handles query sel[i] ...
  const uint32_t* sel_n;       // ... for i < *sel_n (nullable: B)
  const uint64_t* fast_keys;   // [position i][kf]
  const float* qscale;         // [B] scale of the fast pass that produced the keys
  uint32_t* out_ord;           // [B][k]
  float* out_score;
  uint32_t* out_n;
  uint32_t* flagged;           // [B] queries whose proof failed
  uint32_t* n_flagged;         // device counter (zeroed by the plan kernel)
  uint32_t margin_f;           // fixed-point units added to the bound (0; tests use it to force the fallbacks)
};

cudaError_t trr_launch_bm25_build(const Bm25BuildArgs& a, cudaStream_t st);
cudaError_t trr_launch_bm25_merge(const Bm25MergeArgs& a, cudaStream_t st);
cudaError_t trr_launch_bm25_check(const uint32_t* skip, uint32_t skip_ld, uint32_t n_terms, uint32_t n_ranges,
                                  const uint64_t* term_off, const uint2* post, uint64_t n_postings, uint32_t n_docs,
                                  uint32_t* bad, cudaStream_t st);
cudaError_t trr_launch_bm25_kill(const uint2* post, uint32_t* tf, uint64_t n_postings, const uint32_t* dead_bits,
                                 uint32_t* n_killed, cudaStream_t st);
size_t trr_bm25_search_smem(uint32_t range_shift, uint32_t stage_cap, uint32_t cand_cap);
// plan_keys: scratch of max(pow2ceil(B), 1) u64
cudaError_t trr_launch_bm25_plan(const Bm25SearchArgs& a, uint64_t* plan_keys, cudaStream_t st);
cudaError_t trr_launch_bm25_search(const Bm25SearchArgs& a, unsigned grid, cudaStream_t st);
// integer fast pass + exact re-scoring (the default BM25 search path)
size_t trr_bm25_fast_smem(int bits, uint32_t range_shift, uint32_t n_acc, uint32_t stage_cap, uint32_t cand_cap);
cudaError_t trr_launch_bm25_fast(const Bm25SearchArgs& a, int bits, unsigned grid, cudaStream_t st);
cudaError_t trr_launch_bm25_rescore(const Bm25RescoreArgs& a, cudaStream_t st);
#ifndef TRR_BM25_FAST_STAGES_N
#define TRR_BM25_FAST_STAGES_N 2  /* (two 64 KB stages: nearly every range is one pass; 5.81 vs 5.91 ms with three 48 KB stages at cfg4) */
#endif
constexpr uint32_t TRR_BM25_FAST_STAGES = TRR_BM25_FAST_STAGES_N;
// 16-bit level over 32K-document ranges: two 64 KB accumulators in a ring (default), or one 128 KB accumulator over two
// ranges of the skip table (TRR_BM25_FAST_RMUL16 = 1)
#ifndef TRR_BM25_FAST_RMUL16
#define TRR_BM25_FAST_RMUL16 0
#endif
