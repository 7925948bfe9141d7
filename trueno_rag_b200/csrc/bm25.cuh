// bm25.cuh — argument blocks of the BM25 kernels (bm25.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int TRR_BM25_THREADS = 512;

struct Bm25BuildArgs {
  uint64_t n_postings;
  uint32_t n_terms, n_docs;
  const uint64_t* term_off;  // n_terms + 1
  const uint32_t* post_doc;
  const uint32_t* post_tf;
  const uint32_t* doc_len;
  const float* idf;
  float avgdl, k1, b;
  uint32_t range_shift, n_ranges, skip_ld;
  uint2* post;     // out: {doc, impact bits}
  uint32_t* skip;  // out: [n_terms][skip_ld]
};

struct Bm25SearchArgs {
  const uint2* post;
  const uint32_t* skip;
  uint32_t skip_ld, n_terms, n_docs, n_ranges, range_shift, doc_base;
  const uint32_t* q_terms;
  const uint32_t* q_off;
  uint32_t B, k;
  uint32_t stage_cap;  // postings staged per batch
  uint32_t cand_cap;   // power of two >= k + TRR_BM25_THREADS
  uint32_t* counter;   // two dynamic work queues (fast kernel, general kernel), zeroed by the caller
  const uint32_t* fast_list;  // queries served by the fast kernel (<= BM25_FAST_TMAX terms)
  uint32_t n_fast;
  uint32_t* slow_list; // queries for the general kernel: host-listed ones first, the fast kernel appends
  uint32_t* n_slow;    // device count of slow_list
  uint64_t* out_keys;  // nullable [B][k]
  uint32_t* out_ord;   // nullable [B][k]
  float* out_score;    // nullable
  uint32_t* out_n;     // nullable
};

cudaError_t trr_launch_bm25_build(const Bm25BuildArgs& a, cudaStream_t st);
constexpr uint32_t TRR_BM25_FAST_TMAX = 128;
size_t trr_bm25_general_smem(const Bm25SearchArgs& a);
size_t trr_bm25_fast_smem(const Bm25SearchArgs& a);
cudaError_t trr_launch_bm25_search(const Bm25SearchArgs& a, unsigned grid_fast, unsigned grid_slow, cudaStream_t st);
