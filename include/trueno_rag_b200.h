/* trueno_rag_b200.h — C ABI of the B200-native retrieval hot path for trueno-rag.
 *
 * This is the drop-in boundary.  The reference (trueno-rag v0.1.8, Rust) has no
 * FFI of its own: `VectorStore`, `BM25Index`, `FusionStrategy` and `HybridRetriever` are concrete
 * Rust types.  A Rust `-sys` crate binds the entry points below and the bodies of those types
 * become thin shims over them (INTEGRATION.md shows the binding).  Every entry point cites the
 * reference code it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - every function returns a trr_status (0 = OK); trr_last_error() gives a thread-local message;
 *   - plain pointers and sizes only; the caller allocates every output buffer;
 *   - handles are opaque and own device memory; one trr_ctx == one GPU (one process per GPU);
 *   - documents are addressed by insertion ordinal (u32); the ChunkId<->ordinal map stays in the
 *     host language.  A shard holds a contiguous ordinal range [base, base+n);
 *   - result order is canonical: score descending, ordinal ascending;
 *   - entry points taking HOST buffers copy in/out and synchronise before returning;
 *     `_device` variants take DEVICE pointers, enqueue on the context stream and do not sync;
 *   - there is no CPU fallback: without a CUDA device trr_ctx_create fails with TRR_ERR_NO_DEVICE.
 *   - search entry points are safe to call concurrently on one handle from several host threads
 *     (internally serialised per context); mutation (append/remove/build) requires exclusivity,
 *     which the Rust borrow rules (&mut self) already guarantee.
 */
#ifndef TRUENO_RAG_B200_H
#define TRUENO_RAG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define TRR_API
#else
#define TRR_API __attribute__((visibility("default")))
#endif

typedef enum {
  TRR_OK = 0,
  TRR_ERR_INVALID_ARG = 1,   /* -> Error::InvalidConfig (src/error.rs) */
  TRR_ERR_DIM_MISMATCH = 2,  /* -> Error::DimensionMismatch{expected,actual} (src/index.rs:366-369,388-391) */
  TRR_ERR_CUDA = 3,          /* -> Error::VectorStore(String) (src/error.rs:38-39) */
  TRR_ERR_OOM = 4,
  TRR_ERR_NOT_FROZEN = 5,
  TRR_ERR_UNSUPPORTED = 6,
  TRR_ERR_NO_DEVICE = 7
} trr_status;

/* DistanceMetric, src/index.rs:310-319 */
typedef enum { TRR_METRIC_COSINE = 0, TRR_METRIC_EUCLIDEAN = 1, TRR_METRIC_DOT = 2 } trr_metric;
/* storage type of the embedding slab; the arithmetic is always f32 over the stored values */
typedef enum { TRR_DTYPE_F32 = 0, TRR_DTYPE_BF16 = 1 } trr_dtype;
/* FusionStrategy, src/fusion.rs:9-31.  `param` = RRF k | Linear dense_weight | Convex alpha | unused */
typedef enum {
  TRR_FUSE_RRF = 0, TRR_FUSE_LINEAR = 1, TRR_FUSE_CONVEX = 2, TRR_FUSE_DBSF = 3, TRR_FUSE_UNION = 4,
  TRR_FUSE_INTERSECTION = 5
} trr_fusion;
/* which dense kernel serves a search: AUTO picks SCAN for small batches and GEMM for large ones */
typedef enum { TRR_DENSE_AUTO = 0, TRR_DENSE_SCAN = 1, TRR_DENSE_GEMM = 2 } trr_dense_mode;

typedef struct trr_ctx trr_ctx;
typedef struct trr_dense trr_dense;
typedef struct trr_bm25 trr_bm25;

/* per-handle counters of the last search (diagnostics; all device times in milliseconds) */
typedef struct {
  uint32_t mode_used;        /* trr_dense_mode actually run */
  uint32_t n_queries;
  uint32_t n_guard_fallbacks;/* queries whose candidate proof failed: GEMM path -> re-run through SCAN; BM25 -> re-run with 32-bit cells */
  uint32_t n_kernel_launches;/* kernels of this library launched by the call */
  float ms_total;            /* device time of the call (CUDA events on the context stream) */
  float ms_main_kernel;      /* device time of the dominant kernel (scan / gemm / bm25) */
  float max_fast_exact_gap;  /* GEMM path: max |fast score - exact score| over rescored candidates */
  float eps_bound;           /* GEMM path: the a-priori bound used by the candidate proof */
  uint32_t n_exact_fallbacks;/* BM25: queries whose 32-bit proof failed too and were re-run by the exact kernel */
  uint32_t rescore_width;    /* GEMM path: candidates re-scored exactly per query in the first pass (adapts to the data) */
} trr_stats;

TRR_API const char* trr_last_error(void);
TRR_API int trr_version(void);
/* number of CUDA devices visible (0 when there is none) */
TRR_API int trr_device_count(void);

/* ---- context --------------------------------------------------------------------------------- */
/* replaces: nothing (the reference has no device); created lazily by VectorStore::new in the shim */
TRR_API int trr_ctx_create(int device, trr_ctx** out);
TRR_API int trr_ctx_destroy(trr_ctx* ctx);
TRR_API int trr_ctx_sync(trr_ctx* ctx);
/* the CUDA stream (cudaStream_t) every `_device` entry point enqueues on */
TRR_API int trr_ctx_stream(trr_ctx* ctx, void** out_stream);
TRR_API int trr_ctx_sm_count(trr_ctx* ctx, int* out);
/* makes the context enqueue on a caller-owned stream (e.g. the host framework's current stream) from now on */
TRR_API int trr_ctx_set_stream(trr_ctx* ctx, void* stream);
/* number of kernels of this library launched on the context since it was created (diagnostics / bench) */
TRR_API int trr_ctx_launch_count(trr_ctx* ctx, uint64_t* out);

/* ---- dense store: VectorStore, src/index.rs:322-437 ------------------------------------------ */
/* VectorStore::new / with_dimension (src/index.rs:335-350) */
TRR_API int trr_dense_create(trr_ctx* ctx, uint32_t dim, int metric, int dtype, uint64_t capacity_hint,
                             trr_dense** out);
TRR_API int trr_dense_destroy(trr_dense* h);
/* global ordinal of local row 0 (document sharding); default 0 */
TRR_API int trr_dense_set_base(trr_dense* h, uint32_t base_ordinal);
/* VectorStore::insert / insert_batch (src/index.rs:359-383): rows are n x dim f32 on the host; ordinals
 * are assigned in call order.  A bf16 store rounds to nearest-even. */
TRR_API int trr_dense_append(trr_dense* h, const float* rows, uint64_t n);
/* same, rows given as raw bf16 bits (bf16 stores only) */
TRR_API int trr_dense_append_bf16(trr_dense* h, const uint16_t* rows, uint64_t n);
/* same, rows already on this device in the store's dtype */
TRR_API int trr_dense_append_device(trr_dense* h, const void* d_rows, uint64_t n);
/* VectorStore::remove (src/index.rs:421-424): tombstones a local ordinal */
TRR_API int trr_dense_remove(trr_dense* h, uint32_t ordinal);
/* builds the derived arrays (norms in reference summation order, GEMM operands); implicit on first search */
TRR_API int trr_dense_freeze(trr_dense* h);
/* VectorStore::len (src/index.rs:428-430): live rows */
TRR_API int trr_dense_len(trr_dense* h, uint64_t* out_n);
TRR_API int trr_dense_set_mode(trr_dense* h, int mode);
/* VectorStore::search (src/index.rs:386-412) for B queries at once (the reference has no batch API;
 * B = 1 is the reference call).  q: B x dim f32.  Outputs: out_ord/out_score are B x k (row b holds
 * out_n[b] = min(k, live rows) valid entries, canonical order).  Scores are the reference's f32 values. */
TRR_API int trr_dense_search(trr_dense* h, const float* q, uint32_t B, uint32_t k, uint32_t* out_ord,
                             float* out_score, uint32_t* out_n);
TRR_API int trr_dense_search_device(trr_dense* h, const float* d_q, uint32_t B, uint32_t k, uint32_t* d_ord,
                                    float* d_score, uint32_t* d_n);
TRR_API int trr_dense_last_stats(trr_dense* h, trr_stats* out);
/* test/diagnostic: copies the stored norms (sqrt of the sequential f32 sum of squares) to the host */
TRR_API int trr_dense_copy_norms(trr_dense* h, float* out_norms, uint64_t n);
/* test/diagnostic: copies n stored rows (LOCAL ordinals, in the store's dtype, dim elements each) to the host */
TRR_API int trr_dense_copy_rows(trr_dense* h, const uint32_t* ordinals, uint64_t n, void* out_rows);

/* ---- sparse index: BM25Index, src/index.rs:30-280 -------------------------------------------- */
/* Builds the device index from a host CSR over term ids (the tokenizer src/index.rs:111-124 and the
 * String->id dictionary stay in the host language).  Replaces the state built by BM25Index::add
 * (src/index.rs:176-204): term_off[n_terms+1] into post_doc/post_tf; postings of a term sorted by
 * LOCAL doc id (0..n_docs-1); doc_len = token count after filtering; avgdl, k1, b as in the struct;
 * idf[t] = ln((N - df + 0.5)/(df + 0.5) + 1) computed by the host with the platform logf (src/index.rs:147;
 * N and df are GLOBAL statistics when the corpus is sharded).  doc_base = global ordinal of local doc 0. */
TRR_API int trr_bm25_build(trr_ctx* ctx, uint32_t n_docs, uint32_t n_terms, const uint64_t* term_off,
                           const uint32_t* post_doc, const uint32_t* post_tf, const uint32_t* doc_len, float avgdl,
                           float k1, float b, const float* idf, uint32_t doc_base, trr_bm25** out);
/* BM25Index::add after the index was built (src/index.rs:176-204): appends n_new_docs documents given as a CSR over
 * term ids with doc ids RELATIVE to the first new document (0..n_new_docs-1); n_terms_new >= the current vocabulary
 * size (new terms get the next ids).  avgdl / idf are the NEW global statistics; all impacts are re-weighted on the
 * device (they depend on N, df and avgdl), so the result is bit-identical to trr_bm25_build over all documents. */
TRR_API int trr_bm25_append(trr_bm25* h, uint32_t n_new_docs, uint32_t n_terms_new, const uint64_t* delta_term_off,
                            const uint32_t* delta_post_doc, const uint32_t* delta_post_tf, const uint32_t* delta_doc_len,
                            float avgdl, float k1, float b, const float* idf);
/* BM25Index::remove (src/index.rs:245-275) without a rebuild: the removed documents' postings stay in place with weight
 * +0.0 (such a document scores 0.0 and is dropped by `score > 0.0`, src/index.rs:236); all other postings are re-weighted
 * with the NEW global statistics (N, df -> idf, avgdl).  Results equal a rebuild without those documents bit for bit.
 * ordinals are LOCAL doc ids; out_dead_postings (nullable) = postings of removed documents still held by the index. */
TRR_API int trr_bm25_remove(trr_bm25* h, const uint32_t* ordinals, uint32_t n, float avgdl, float k1, float b,
                            const float* idf, uint64_t* out_dead_postings);
TRR_API int trr_bm25_destroy(trr_bm25* h);
TRR_API int trr_bm25_n_postings(trr_bm25* h, uint64_t* out);
/* BM25Index::search (src/index.rs:212-243) for B tokenised queries: q_terms holds the term ids of all
 * queries back to back in query order, duplicates kept, unknown terms as 0xFFFFFFFF;
 * q_off[B+1] delimits them.  Outputs as trr_dense_search; only documents with score > 0.0 are returned. */
TRR_API int trr_bm25_search(trr_bm25* h, const uint32_t* q_terms, const uint32_t* q_off, uint32_t B, uint32_t k,
                            uint32_t* out_ord, float* out_score, uint32_t* out_n);
TRR_API int trr_bm25_search_device(trr_bm25* h, const uint32_t* d_q_terms, const uint32_t* d_q_off,
                                   const uint32_t* h_q_off, uint32_t B, uint32_t k, uint32_t* d_ord,
                                   float* d_score, uint32_t* d_n);
TRR_API int trr_bm25_last_stats(trr_bm25* h, trr_stats* out);
/* test/diagnostic: per-posting BM25 impact idf*tf_norm (src/index.rs:136-154) as stored on the device */
TRR_API int trr_bm25_copy_impacts(trr_bm25* h, float* out, uint64_t n);

/* ---- fusion: FusionStrategy::fuse, src/fusion.rs:42-231 -------------------------------------- */
/* Fuses, per query, a dense and a sparse result list.  Lists are B x C (row stride C, valid prefix
 * d_n[b] / s_n[b]).  Output rows are B x k_out where k_out >= 1: the first out_n[b] = min(k_out, fused
 * length) entries of the fused ranking.  out_dense/out_sparse carry the source scores of each fused id or
 * NaN when the id was not in that source's list (RetrievalResult, src/retrieve.rs:13-24, 204-213); either
 * may be NULL. */
TRR_API int trr_fuse(trr_ctx* ctx, int strategy, float param, const uint32_t* d_ord, const float* d_score,
                     const uint32_t* d_n, const uint32_t* s_ord, const float* s_score, const uint32_t* s_n,
                     uint32_t B, uint32_t C, uint32_t k_out, uint32_t* out_ord, float* out_fused, float* out_dense,
                     float* out_sparse, uint32_t* out_n);

/* ---- hybrid: HybridRetriever::retrieve, src/retrieve.rs:175-220 ------------------------------ */
/* One call = dense top-C (if use_dense) + sparse top-C (if use_sparse) + fuse + take(k) for B queries.
 * C = HybridRetrieverConfig::candidates_per_source (src/retrieve.rs:80-100, default 50). */
TRR_API int trr_hybrid_search(trr_dense* dense, trr_bm25* bm25, const float* q, const uint32_t* q_terms,
                              const uint32_t* q_off, uint32_t B, uint32_t C, int strategy, float param, uint32_t k,
                              int use_dense, int use_sparse, uint32_t* out_ord, float* out_fused, float* out_dense,
                              float* out_sparse, uint32_t* out_n);

/* ---- document-sharded execution (one process per GPU) ----------------------------- */
/* Size in bytes of one rank's exchange record for B queries and C candidates per source:
 * layout { u32 ord[2][B][C]; f32 score[2][B][C]; u32 n[2][B]; } with source 0 = dense, 1 = sparse. */
TRR_API size_t trr_exchange_bytes(uint32_t B, uint32_t C);
/* Shard-local stage: writes this shard's top-C per source (global ordinals, canonical order) into the
 * DEVICE exchange record.  q / q_terms / q_off are HOST buffers (copied inside). */
TRR_API int trr_hybrid_local(trr_dense* dense, trr_bm25* bm25, const float* q, const uint32_t* q_terms,
                             const uint32_t* q_off, uint32_t B, uint32_t C, int use_dense, int use_sparse,
                             void* d_exchange);
/* Same with DEVICE inputs (h_q_off is the host copy of q_off, used for validation only); enqueues, no sync. */
TRR_API int trr_hybrid_local_device(trr_dense* dense, trr_bm25* bm25, const float* d_q, const uint32_t* d_q_terms,
                                    const uint32_t* d_q_off, const uint32_t* h_q_off, uint32_t B, uint32_t C,
                                    int use_dense, int use_sparse, void* d_exchange);
/* Merge stage: d_gathered holds G exchange records back to back (the all-gather output, DEVICE).  Merges the
 * G sorted lists per source into the global top-C, then fuses and takes k exactly as trr_hybrid_search.
 * Outputs are HOST buffers. */
TRR_API int trr_hybrid_merge(trr_ctx* ctx, const void* d_gathered, uint32_t G, uint32_t B, uint32_t C, int strategy,
                             float param, uint32_t k, uint32_t* out_ord, float* out_fused, float* out_dense,
                             float* out_sparse, uint32_t* out_n);

/* Same with DEVICE outputs; enqueues, no sync. */
TRR_API int trr_hybrid_merge_device(trr_ctx* ctx, const void* d_gathered, uint32_t G, uint32_t B, uint32_t C, int strategy,
                                    float param, uint32_t k, uint32_t* d_out_ord, float* d_out_fused, float* d_out_dense,
                                    float* d_out_sparse, uint32_t* d_out_n);

/* ---- sharded search: one process per GPU, the exchange inside the call ------------------- */
/* HybridRetriever::retrieve (reference src/retrieve.rs:175-220) over a corpus sharded by document across the GPUs of one
 * node, one process (rank) per GPU: ONE call = shard-local dense + BM25 top-C -> cross-GPU exchange -> merge of the G
 * sorted lists per source + fusion + top-k on every rank.  The group owns the communicator (NCCL, resolved with dlopen
 * at run time) and the exchange buffers; the host language only carries the 128-byte rendezvous id from rank 0 to the
 * other ranks.  Two exchanges: NCCL all-gather, or peer memory - every rank stores its record straight into the gather
 * buffers of all peers over NVLink (CUDA IPC mappings set up once) and publishes a flag, so a step has no collective
 * launch; the peer exchange falls back to NCCL when IPC is not available. */
typedef struct trr_group trr_group;
enum trr_exchange { TRR_EXCHANGE_NCCL = 0, TRR_EXCHANGE_PEER = 1 };
TRR_API int trr_group_unique_id(void* out_id128);                 /* rank 0: ncclGetUniqueId */
/* collective over the ranks.  max_record_bytes: the largest trr_exchange_bytes(B, C) the peer exchange has to serve
 * (larger batches use NCCL); ignored for TRR_EXCHANGE_NCCL.  world == 1 needs no id and no NCCL. */
TRR_API int trr_group_create(trr_ctx* ctx, const void* id128, int rank, int world, int exchange, size_t max_record_bytes,
                             trr_group** out);
TRR_API int trr_group_destroy(trr_group* g);
TRR_API int trr_group_info(trr_group* g, int* out_rank, int* out_world, int* out_exchange /* trr_exchange in use */);
/* waits for everything enqueued by the *_device call below (both streams) */
TRR_API int trr_group_sync(trr_group* g);
/* element-wise sum (op_max = 0) or max (1) over the ranks of n u64 values in HOST memory: the global BM25 statistics
 * (df, total document length; reference src/index.rs:157-164) every shard's impacts have to be computed from */
TRR_API int trr_group_allreduce_u64(trr_group* g, uint64_t* inout, size_t n, int op_max);
/* HOST buffers, blocking; arguments as trr_hybrid_search; every rank passes the same queries and receives the same result */
TRR_API int trr_hybrid_search_sharded(trr_group* g, trr_dense* dense, trr_bm25* bm25, const float* q, const uint32_t* q_terms,
                                      const uint32_t* q_off, uint32_t B, uint32_t C, int strategy, float param, uint32_t k,
                                      int use_dense, int use_sparse, uint32_t* out_ord, float* out_fused, float* out_dense,
                                      float* out_sparse, uint32_t* out_n);
/* The same call without the final wait: the host buffers must stay valid (and should be page-locked) until
 * trr_group_sync; consecutive calls overlap their input / result copies with each other's kernels. */
TRR_API int trr_hybrid_search_sharded_async(trr_group* g, trr_dense* dense, trr_bm25* bm25, const float* q,
                                            const uint32_t* q_terms, const uint32_t* q_off, uint32_t B, uint32_t C, int strategy,
                                            float param, uint32_t k, int use_dense, int use_sparse, uint32_t* out_ord,
                                            float* out_fused, float* out_dense, float* out_sparse, uint32_t* out_n);
/* DEVICE buffers; enqueues and returns.  The exchange + merge of a call run on the group's second stream and overlap the
 * shard-local kernels of the next call; outputs are valid after trr_group_sync (h_q_off: host copy of q_off). */
TRR_API int trr_hybrid_search_sharded_device(trr_group* g, trr_dense* dense, trr_bm25* bm25, const float* d_q,
                                             const uint32_t* d_q_terms, const uint32_t* d_q_off, const uint32_t* h_q_off,
                                             uint32_t B, uint32_t C, int strategy, float param, uint32_t k, int use_dense,
                                             int use_sparse, uint32_t* d_out_ord, float* d_out_fused, float* d_out_dense,
                                             float* d_out_sparse, uint32_t* d_out_n);

/* ---- persistence -> device load --------------------------------------------- */
/* The reference can serialise only BM25Index (bincode + LZ4/ZSTD, src/compressed.rs:71-108); VectorStore is not
 * serialisable (src/compressed.rs:9-10) and the CLI dumps embeddings as JSON (crates/trueno-rag-cli/src/main.rs:135-146).
 * These entry points add flat little-endian snapshot files of the DEVICE structures, so that a 10M-document index is back
 * in HBM at file-read speed instead of being re-inserted / re-built: header, then the arrays exactly as they live on the
 * device (embedding slab + tombstones; postings with impacts + skip table + per-term minimum impacts).  The ChunkId <->
 * ordinal map, chunk texts and the term dictionary stay with the host language, as everywhere else in this ABI. */
/* configuration of a store (e.g. one restored from a snapshot): dimension, trr_metric, trr_dtype, base ordinal */
TRR_API int trr_dense_info(trr_dense* h, uint32_t* out_dim, int* out_metric, int* out_dtype, uint32_t* out_base);
TRR_API int trr_dense_save(trr_dense* h, const char* path);
TRR_API int trr_dense_load(trr_ctx* ctx, const char* path, trr_dense** out);
TRR_API int trr_bm25_save(trr_bm25* h, const char* path);
TRR_API int trr_bm25_load(trr_ctx* ctx, const char* path, trr_bm25** out);

/* ---- synthetic inputs for tests and benches (not reference behaviour) ------------- */
/* appends n rows generated on the device by the counter-based recipe of csrc/synth_spec.h */
TRR_API int trr_dense_append_synth(trr_dense* h, uint64_t seed, uint64_t first_row, uint64_t n, int dups);
/* host-side generators of the synthetic hybrid workload (csrc/synth_spec.h): queries, query terms and the BM25
 * postings of the documents [doc_lo, doc_hi) as a CSR with LOCAL doc ids.  Two passes so that the caller can
 * all-reduce df / total length across shards before computing term_off and idf:
 *   count: df_local[n_terms] (documents of this shard containing the term), doc_len[doc_hi-doc_lo], total length;
 *   fill : post_doc / post_tf given term_off[n_terms+1] = exclusive prefix sum of df_local. */
TRR_API int trr_synth_queries(uint64_t seed, uint64_t q0, uint64_t n, uint32_t dim, uint64_t n_corpus, int corpus_bf16,
                              int dups, int round_to_bf16, float* out);
TRR_API int trr_synth_query_terms(uint64_t seed, const uint64_t* cdf, uint32_t n_terms, uint64_t q0, uint64_t n,
                                  uint32_t* q_off, uint32_t* out_terms, uint64_t out_cap);
TRR_API int trr_synth_bm25_count(uint64_t seed, const uint64_t* cdf, uint32_t n_terms, uint64_t doc_lo, uint64_t doc_hi,
                                 uint32_t* df_local, uint32_t* doc_len, uint64_t* total_len);
TRR_API int trr_synth_bm25_fill(uint64_t seed, const uint64_t* cdf, uint32_t n_terms, uint64_t doc_lo, uint64_t doc_hi,
                                const uint64_t* term_off, uint32_t* post_doc, uint32_t* post_tf);
/* L2 flush helper for benches: writes `bytes` of device memory on the context stream */
TRR_API int trr_ctx_flush_l2(trr_ctx* ctx, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* TRUENO_RAG_B200_H */
