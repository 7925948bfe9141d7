/* trueno_rag_host.h — flat C view of the C++ host mirror (include/trueno_rag.hpp) so that the pytest suite can
 * drive `VectorStore`, `BM25Index`, `FusionStrategy::fuse` and `HybridRetriever` through ctypes exactly the way the
 * reference's own Rust tests drive them.  A Rust integration would NOT use this layer (its host logic is Rust, see
 * INTEGRATION.md); it binds trueno_rag_b200.h directly.
 * Status codes: 0 ok, 1 InvalidConfig, 2 DimensionMismatch (expected/actual via trrh_last_expected/actual),
 * 3 VectorStore (device error), 6 Unsupported, 7 SerializationError. */
#ifndef TRUENO_RAG_HOST_H
#define TRUENO_RAG_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#define TRRH_API __attribute__((visibility("default")))
typedef struct trrh_store trrh_store;
typedef struct trrh_bm25 trrh_bm25;
typedef struct trrh_retriever trrh_retriever;
typedef struct { uint64_t hi, lo; } trrh_id;

TRRH_API const char* trrh_last_error(void);
TRRH_API uint64_t trrh_last_expected(void);
TRRH_API uint64_t trrh_last_actual(void);

/* VectorStore (reference src/index.rs:322-437) */
TRRH_API int trrh_store_new(uint32_t dim, int metric, int dtype, trrh_store** out);
TRRH_API void trrh_store_free(trrh_store* s);
TRRH_API int trrh_store_insert(trrh_store* s, trrh_id id, const char* content, const float* emb, uint32_t emb_len, int has_emb);
TRRH_API int trrh_store_search(trrh_store* s, const float* q, uint32_t q_len, uint32_t k, trrh_id* out_ids, float* out_scores, uint32_t* out_n);
TRRH_API int trrh_store_get(trrh_store* s, trrh_id id, const char** out_content);
TRRH_API int trrh_store_remove(trrh_store* s, trrh_id id);
TRRH_API uint64_t trrh_store_len(trrh_store* s);
TRRH_API int trrh_store_set_mode(trrh_store* s, int mode);
TRRH_API int trrh_store_clone(trrh_store* s, trrh_store** out);

/* BM25Index (reference src/index.rs:30-280) */
TRRH_API int trrh_bm25_new(float k1, float b, trrh_bm25** out);
TRRH_API void trrh_bm25_free(trrh_bm25* s);
TRRH_API int trrh_bm25_tokenize(trrh_bm25* s, const char* text, char* out, uint32_t cap, uint32_t* out_len);
TRRH_API int trrh_bm25_add(trrh_bm25* s, trrh_id id, const char* content);
TRRH_API int trrh_bm25_search(trrh_bm25* s, const char* query, uint32_t k, trrh_id* out_ids, float* out_scores, uint32_t* out_n);
TRRH_API int trrh_bm25_remove(trrh_bm25* s, trrh_id id);
TRRH_API uint64_t trrh_bm25_len(trrh_bm25* s);
TRRH_API float trrh_bm25_avgdl(trrh_bm25* s);
TRRH_API float trrh_bm25_k1(trrh_bm25* s);
TRRH_API float trrh_bm25_b(trrh_bm25* s);
TRRH_API int trrh_bm25_contains_term(trrh_bm25* s, const char* term);
/* Persistence in the reference's format (src/compressed.rs:13-108; bincode 1.3 of the struct at src/index.rs:30-51).
 * compression: -1 = plain bincode, 0 = LZ4 (lz4_flex size-prepended block), 1 = ZSTD (standard frames).
 * Output buffers are owned by the library until trrh_bytes_free.  Status 7 = SerializationError. */
TRRH_API int trrh_compress(int compression, const uint8_t* data, uint64_t n, uint8_t** out, uint64_t* out_n);
TRRH_API int trrh_decompress(int compression, const uint8_t* data, uint64_t n, uint8_t** out, uint64_t* out_n);
TRRH_API void trrh_bytes_free(uint8_t* p);
TRRH_API int trrh_bm25_to_bytes(trrh_bm25* s, int compression, uint8_t** out, uint64_t* out_n);
TRRH_API int trrh_bm25_from_bytes(const uint8_t* data, uint64_t n, int compression, trrh_bm25** out);

/* The CLI's index.json (reference crates/trueno-rag-cli/src/main.rs:133-154): parse (:437-439) and the brute-force query
 * scan (:479-492) on the device.  Strings are owned by the handle; NULL = None. */
typedef struct trrh_cli_index trrh_cli_index;
TRRH_API int trrh_cli_index_from_json(const char* text, uint64_t n, trrh_cli_index** out);
/* building and writing one (:407-424): an empty index, chunks with their embeddings, serde_json::to_string_pretty */
TRRH_API int trrh_cli_index_new(uint64_t dimension, const char* embedder_type, const char* model_name, trrh_cli_index** out);
TRRH_API int trrh_cli_index_push(trrh_cli_index* h, const char* content, const char* title, const char* source,
                                 const float* embedding, uint64_t len);
TRRH_API int trrh_cli_index_to_json(trrh_cli_index* h, uint8_t** out, uint64_t* out_n);
TRRH_API void trrh_cli_index_free(trrh_cli_index* h);
TRRH_API uint64_t trrh_cli_index_len(trrh_cli_index* h);
TRRH_API uint64_t trrh_cli_index_n_embeddings(trrh_cli_index* h);
TRRH_API uint64_t trrh_cli_index_dimension(trrh_cli_index* h);
TRRH_API const char* trrh_cli_index_embedder_type(trrh_cli_index* h);
TRRH_API const char* trrh_cli_index_model_name(trrh_cli_index* h);
TRRH_API int trrh_cli_index_chunk(trrh_cli_index* h, uint64_t i, const char** content, uint64_t* content_len,
                                  const char** title, const char** source);
TRRH_API int trrh_cli_index_embedding(trrh_cli_index* h, uint64_t i, const float** data, uint64_t* len);
TRRH_API int trrh_cli_index_query(trrh_cli_index* h, const float* q, uint64_t q_len, uint64_t top_k, uint64_t* out_idx,
                                  float* out_score, uint64_t* out_n);

/* FusionStrategy::fuse (reference src/fusion.rs:42-63); out buffers hold nd + ns entries */
TRRH_API int trrh_fuse(int kind, float param, const trrh_id* d_ids, const float* d_sc, uint32_t nd, const trrh_id* s_ids,
                       const float* s_sc, uint32_t ns, trrh_id* out_ids, float* out_sc, uint32_t* out_n);

/* HybridRetriever (reference src/retrieve.rs:103-263).  Takes ownership of the store and the index.  The Embedder is
 * out of scope: the query embedding is passed with each call. */
TRRH_API int trrh_retriever_new(trrh_store* store, trrh_bm25* bm25, uint32_t candidates_per_source, int fusion_kind,
                                float fusion_param, int use_dense, int use_sparse, trrh_retriever** out);
TRRH_API void trrh_retriever_free(trrh_retriever* r);
TRRH_API int trrh_retriever_index(trrh_retriever* r, trrh_id id, const char* content, const float* emb, uint32_t emb_len, int has_emb);
/* which: 0 = retrieve (hybrid), 1 = retrieve_dense, 2 = retrieve_sparse.  Absent scores are NaN. */
TRRH_API int trrh_retriever_retrieve(trrh_retriever* r, int which, const char* query, const float* q_emb, uint32_t q_len,
                                     uint32_t k, trrh_id* out_ids, float* out_fused, float* out_dense, float* out_sparse,
                                     uint32_t* out_n);
TRRH_API uint64_t trrh_retriever_len(trrh_retriever* r);
#ifdef __cplusplus
}
#endif
#endif
