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
  uint32_t* flags;     // out: [0] = 1 when some impact is not > 0 (disables the threshold bootstrap), pre-set to 0
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
  const uint32_t* flags;     // [0] != 0: impacts are not all positive, no bootstrap
  uint64_t* thr0;      // [B] initial threshold key per query (written by the plan kernel; 0 = none)
  uint32_t* order;     // [B] queries by decreasing posting volume (written by the plan kernel)
  uint32_t* queue;     // [2] dynamic work queue (reset by the plan kernel)
  uint64_t* partial;   // n_chunks > 1: [B][n_chunks][k] keys, merged by topk_merge_kernel
  uint64_t* out_keys;  // n_chunks == 1: nullable [B][k]
  uint32_t* out_ord;   // nullable [B][k]
  float* out_score;    // nullable
  uint32_t* out_n;     // nullable
  uint32_t* dbg;       // nullable host-mapped word: site of a barrier timeout
  uint32_t debug_mode; // perf triage (TRR_BM25_DEBUG): 8 = CTA 0 records where its cycles go
  // bm25_search_warp_kernel (TRR_BM25_V2=1) only: fine skip table of the frequent terms, 2048-document sub-ranges
  const uint32_t* fine_row;  // [n_terms] row of the term in `fine`, 0xFFFFFFFF = not a frequent term
  const uint32_t* fine;      // [n_fine][fine_ld]: first posting of the term with doc >= j * 2048
  uint32_t fine_ld, n_sub;
};

constexpr uint32_t TRR_BM25_SUB_SHIFT = 11;   // documents per warp accumulator of the V2 kernel: 2048
constexpr uint32_t TRR_BM25_V2_WARPS = 8;
constexpr uint32_t TRR_BM25_FINE_MIN_DF = 2048;  // terms at least this frequent get a row in the fine skip table

cudaError_t trr_launch_bm25_build(const Bm25BuildArgs& a, cudaStream_t st);
cudaError_t trr_launch_bm25_merge(const Bm25MergeArgs& a, cudaStream_t st);
cudaError_t trr_launch_bm25_kill(const uint2* post, uint32_t* tf, uint64_t n_postings, const uint32_t* dead_bits,
                                 uint32_t* n_killed, cudaStream_t st);
size_t trr_bm25_search_smem(uint32_t range_shift, uint32_t stage_cap, uint32_t cand_cap);
// plan_keys: scratch of max(pow2ceil(B), 1) u64
cudaError_t trr_launch_bm25_plan(const Bm25SearchArgs& a, uint64_t* plan_keys, cudaStream_t st);
cudaError_t trr_launch_bm25_search(const Bm25SearchArgs& a, unsigned grid, cudaStream_t st);
// V2 (opt-in): warps of a CTA work on different 2048-document sub-ranges of one query without per-range barriers
size_t trr_bm25_search_warp_smem(uint32_t cand_cap);
cudaError_t trr_launch_bm25_search_warp(const Bm25SearchArgs& a, unsigned grid, int variant, cudaStream_t st);
// fine skip table: fine[row][j] for the n_fine terms listed in fine_terms (term ids), j = 0..n_sub
cudaError_t trr_launch_bm25_fine(const uint2* post, const uint64_t* term_off, const uint32_t* fine_terms, uint32_t n_fine,
                                 uint32_t* fine, uint32_t fine_ld, uint32_t n_sub, cudaStream_t st);
