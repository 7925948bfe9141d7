// trueno_rag.hpp — C++ host-side mirror of the reference's retrieval API, sitting on the C ABI of
// trueno_rag_b200.h.  The reference is Rust (no rustc in the build image), so the host layer that the
// Rust shim would contain is written here in C++ with the same type names, method names, argument meaning
// and error behaviour, so that the parity tests read like the reference's own tests:
//
//   reference (Rust)                         this header
//   ---------------------------------------  ----------------------------------------------
//   ChunkId, Chunk            src/chunk.rs:9-99          trueno_rag::ChunkId, Chunk
//   Error                     src/error.rs:9-64          trueno_rag::Error (kind + fields)
//   VectorStoreConfig, DistanceMetric, VectorStore  src/index.rs:283-437   same names
//   SparseIndex, BM25Index    src/index.rs:8-280         BM25Index (tokenize/add/search/remove/len)
//   FusionStrategy::fuse      src/fusion.rs:9-63         FusionStrategy::fuse
//   RetrievalResult, HybridRetrieverConfig, HybridRetriever  src/retrieve.rs:13-263   same names
//
// What stays on the host (as it would in the Rust shim): ChunkId <-> ordinal maps, the Chunk cache behind
// VectorStore::get, the tokenizer and the String -> term-id dictionary, CSR construction, idf (platform logf).
// Everything numeric on the query path runs in the CUDA kernels.  There is no CPU fallback.
#pragma once

#include <cstdint>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "trueno_rag_b200.h"

namespace trueno_rag {

struct ChunkId {
  uint64_t hi = 0, lo = 0;  // the 128 bits of the reference's UUID
  bool operator==(const ChunkId& o) const { return hi == o.hi && lo == o.lo; }
  bool operator!=(const ChunkId& o) const { return !(*this == o); }
  static ChunkId random();  // ChunkId::new (UUID v4)
};
struct ChunkIdHash {
  size_t operator()(const ChunkId& c) const { return std::hash<uint64_t>()(c.hi * 0x9E3779B97F4A7C15ull ^ c.lo); }
};

struct Chunk {  // src/chunk.rs:46-61 (metadata omitted: never read on the retrieval path)
  ChunkId id;
  ChunkId document_id;
  std::string content;
  size_t start_offset = 0, end_offset = 0;
  std::optional<std::vector<float>> embedding;
  Chunk() = default;
  Chunk(std::string text, size_t start, size_t end) : id(ChunkId::random()), content(std::move(text)), start_offset(start), end_offset(end) {}
  void set_embedding(std::vector<float> e) { embedding = std::move(e); }
};

class Error : public std::runtime_error {  // src/error.rs:9-64 (variants reachable from the path)
 public:
  enum class Kind { InvalidConfig, DimensionMismatch, VectorStore, Unsupported, Serialization };
  Kind kind;
  size_t expected = 0, actual = 0;  // DimensionMismatch fields
  Error(Kind k, const std::string& msg, size_t exp = 0, size_t act = 0) : std::runtime_error(msg), kind(k), expected(exp), actual(act) {}
};

enum class DistanceMetric { Cosine = 0, Euclidean = 1, DotProduct = 2 };  // src/index.rs:310-319

struct VectorStoreConfig {  // src/index.rs:283-307
  size_t dimension = 384;
  DistanceMetric metric = DistanceMetric::Cosine;
  size_t hnsw_m = 16, hnsw_ef_construction = 100, hnsw_ef_search = 50;  // carried, never read (as in the reference)
  // additive, B200-specific
  int storage_dtype = TRR_DTYPE_F32;
};

using Scored = std::pair<ChunkId, float>;

namespace detail { struct DeviceDense; struct DeviceBm25; }

class VectorStore {  // src/index.rs:322-437
 public:
  explicit VectorStore(VectorStoreConfig config);
  static VectorStore with_dimension(size_t dimension);
  ~VectorStore();
  VectorStore(VectorStore&&) noexcept;
  VectorStore& operator=(VectorStore&&) noexcept;
  VectorStore(const VectorStore&) = delete;
  VectorStore clone() const;  // #[derive(Clone)]

  const VectorStoreConfig& config() const { return config_; }
  void insert(Chunk chunk);                      // :359-375
  void insert_batch(std::vector<Chunk> chunks);  // :378-383
  std::vector<Scored> search(const std::vector<float>& query_vector, size_t k) const;  // :386-412
  // additive batch API (the reference has none): B queries, row-major
  std::vector<std::vector<Scored>> search_batch(const std::vector<float>& queries, size_t B, size_t k) const;
  const Chunk* get(const ChunkId& id) const;     // :416-418
  std::optional<Chunk> remove(const ChunkId& id);// :421-424
  size_t len() const { return ord_of_.size(); }  // :428-430
  bool is_empty() const { return ord_of_.empty(); }

  // plumbing used by HybridRetriever
  trr_dense* device_handle() const;  // flushes pending inserts
  const ChunkId& id_of(uint32_t ordinal) const { return id_of_[ordinal]; }
  bool ordinal_of(const ChunkId& id, uint32_t* out) const;
  uint32_t next_ordinal() const { return (uint32_t)id_of_.size(); }
  void set_mode(int trr_dense_mode_value);

 private:
  void flush() const;
  VectorStoreConfig config_;
  std::unordered_map<ChunkId, uint32_t, ChunkIdHash> ord_of_;
  std::vector<ChunkId> id_of_;
  std::unordered_map<ChunkId, Chunk, ChunkIdHash> chunks_;
  mutable std::vector<float> pending_;  // rows not yet on the device
  mutable std::shared_ptr<detail::DeviceDense> dev_;
};

// src/compressed.rs:13-66 — compression of serialised indexes, both formats restated here (no liblz4 / libzstd in the
// build).  LZ4 = lz4_flex 0.11 `compress_prepend_size` / `decompress_size_prepended` (u32 little-endian length + one LZ4
// block).  ZSTD = standard frames (RFC 8878): `decompress` is a complete frame decoder (Huffman literals, FSE sequences,
// repeat offsets, checksums; no dictionaries), so anything the reference wrote loads; `compress` writes frames the
// reference reads (LZ77 matches, FSE-coded sequences with the predefined distributions, raw literals - smaller than the
// LZ4 output, larger than libzstd's own).  Empty input <-> empty output.
enum class Compression { Lz4 = 0, Zstd = 1 };
const char* compression_as_str(Compression c);                                  // :24-30
std::vector<uint8_t> compress(Compression c, const uint8_t* data, size_t n);    // :36-47
std::vector<uint8_t> decompress(Compression c, const uint8_t* data, size_t n);  // :53-66

class BM25Index {  // src/index.rs:30-280 (SparseIndex impl included)
 public:
  BM25Index();
  static BM25Index with_params(float k1, float b);
  BM25Index with_stopwords(std::unordered_set<std::string> stopwords) &&;
  ~BM25Index();
  BM25Index(BM25Index&&) noexcept;
  BM25Index& operator=(BM25Index&&) noexcept;
  BM25Index(const BM25Index&) = delete;

  std::vector<std::string> tokenize(const std::string& text) const;  // :111-124
  void add(const Chunk& chunk);                                      // :176-204
  void add_batch(const std::vector<Chunk>& chunks);                  // :206-210
  std::vector<Scored> search(const std::string& query, size_t k) const;  // :212-243
  void remove(const ChunkId& id);                                    // :245-275
  size_t len() const { return doc_count_; }
  bool is_empty() const { return doc_count_ == 0; }
  float k1() const { return k1_; }
  float b() const { return b_; }
  float avg_doc_length() const;
  bool contains_term(const std::string& term) const { return dict_.count(term) != 0; }

  // Persistence in the REFERENCE's format: `#[derive(Serialize, Deserialize)]` of the struct at
  // src/index.rs:30-51 through bincode 1.3 default options (little-endian, fixed-width integers, u64 lengths; ChunkId =
  // uuid bytes with a u64 length of 16), optionally compressed (src/compressed.rs:71-108).  A file written by the
  // reference loads here and vice versa.  The reference's maps carry no insertion order, so a loaded index numbers its
  // chunks by ascending ChunkId and its terms by ascending term string (ties in the reference are hash-order anyway).
  std::vector<uint8_t> to_bytes() const;                                                        // bincode::serialize
  static BM25Index from_bytes(const uint8_t* data, size_t n);                                   // bincode::deserialize
  std::vector<uint8_t> to_compressed_bytes(Compression c) const;                                // :92-94
  static BM25Index from_compressed_bytes(const uint8_t* data, size_t n, Compression c);         // :101-103

  // plumbing used by HybridRetriever
  trr_bm25* device_handle() const;  // rebuilds the device index if the host state changed
  std::vector<uint32_t> term_ids(const std::vector<std::string>& tokens) const;
  const ChunkId& id_of(uint32_t ordinal) const { return id_of_[ordinal]; }
  uint32_t next_ordinal() const { return (uint32_t)id_of_.size(); }

 private:
  void freeze() const;
  std::unordered_map<std::string, uint32_t> dict_;                          // term -> term id
  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> postings_;         // term id -> [(ordinal, tf)]
  std::vector<uint32_t> df_;                                                // doc_freqs
  std::vector<uint32_t> doc_len_;                                           // by ordinal (0 once removed)
  std::vector<uint8_t> live_;
  std::unordered_map<ChunkId, uint32_t, ChunkIdHash> ord_of_;
  std::vector<ChunkId> id_of_;
  uint32_t doc_count_ = 0;
  float k1_ = 1.2f, b_ = 0.75f;
  bool lowercase_ = true;
  std::unordered_set<std::string> stopwords_;
  mutable bool dirty_ = true;
  mutable float avg_doc_length_ = 0.0f;
  std::optional<float> avg_loaded_;  // avg_doc_length of a loaded file that disagrees with its doc_lengths; dropped by add/remove, which recompute (:157-164)
  mutable std::shared_ptr<detail::DeviceBm25> dev_;
  // incremental device updates: documents [0, frozen_docs_) and the first frozen_len_[t] postings of term t are on the
  // device; adds since then go through trr_bm25_append, removes through trr_bm25_remove (rebuild once a quarter of the postings is dead)
  mutable uint32_t frozen_docs_ = 0;
  mutable std::vector<uint32_t> frozen_len_;
  mutable std::vector<uint32_t> pending_removed_;  // ordinals removed since the last freeze (all < frozen_docs_)
  mutable uint64_t frozen_postings_ = 0;
  mutable bool needs_rebuild_ = true;
};

// ------------------------------------------------------------------------------------------------
// The CLI's persisted index (crates/trueno-rag-cli/src/main.rs:133-154): `index.json` written by `trueno-rag index`
// (serde_json::to_string_pretty, :423) and scanned by `trueno-rag query` (:437-439, 479-492).  SURVEY 8(f) ranks 2 and 4:
// the file is parsed straight into an embedding slab and the query's brute-force cosine loop runs through K1.
// ------------------------------------------------------------------------------------------------
struct PersistedChunk {  // :148-153
  std::string content;
  std::optional<std::string> title, source;
};
class PersistedIndex {  // :133-146
 public:
  std::vector<PersistedChunk> chunks;
  std::vector<std::vector<float>> embeddings;
  size_t dimension = 0;
  std::string embedder_type;               // #[serde(default)]
  std::optional<std::string> model_name;   // #[serde(default)]

  // serde_json::from_str::<PersistedIndex>: unknown keys ignored, `chunks` / `embeddings` / `dimension` required, numbers
  // parsed as f64 and narrowed to f32 like serde_json does; malformed input -> Error::Serialization
  static PersistedIndex from_json(const char* text, size_t n);
  // serde_json::to_string_pretty(&persisted) (:423): two-space indentation, struct field order, every f32 printed with the
  // fewest digits that parse back to the same value (what serde_json's ryu does; the exponent style may differ, the value
  // never), non-finite numbers as `null` (which, as in the reference, cannot be read back), non-ASCII text unescaped
  std::string to_json() const;
  // run_query's scoring (:479-492): cosine of `query` against every stored embedding (an embedding of another length
  // scores 0.0, :529-531), stable sort by score descending (ties keep index order), first top_k.  The query embedding comes
  // from the caller (the embedders are out of scope).  Pairs are (index into `chunks`, score).  top_k <= 1024 (the C ABI's
  // per-query limit; larger values raise Error::Unsupported - the CLI's default is 5).
  std::vector<std::pair<size_t, float>> query(const std::vector<float>& query_embedding, size_t top_k) const;

  PersistedIndex();
  ~PersistedIndex();
  PersistedIndex(PersistedIndex&&) noexcept;
  PersistedIndex& operator=(PersistedIndex&&) noexcept;

 private:
  struct DeviceRows {  // the embeddings of one length on the device, with their indexes into `chunks`
    std::shared_ptr<detail::DeviceDense> dev;
    std::vector<uint32_t> index_of;
    size_t n_other = 0;  // embeddings of any other length (score 0.0)
  };
  mutable std::unordered_map<size_t, DeviceRows> by_len_;
};

struct FusionStrategy {  // src/fusion.rs:9-63
  enum class Kind { RRF = 0, Linear = 1, Convex = 2, DBSF = 3, Union = 4, Intersection = 5 };
  Kind kind = Kind::RRF;
  float param = 60.0f;  // RRF k | Linear dense_weight | Convex alpha
  static FusionStrategy RRF(float k) { return {Kind::RRF, k}; }
  static FusionStrategy Linear(float dense_weight) { return {Kind::Linear, dense_weight}; }
  static FusionStrategy Convex(float alpha) { return {Kind::Convex, alpha}; }
  static FusionStrategy DBSF() { return {Kind::DBSF, 0.0f}; }
  static FusionStrategy Union() { return {Kind::Union, 0.0f}; }
  static FusionStrategy Intersection() { return {Kind::Intersection, 0.0f}; }
  // ids are compared by value; ties in the fused score are ordered by first appearance in
  // (dense list, then sparse list) — the canonical stand-in for the reference's hash order
  std::vector<Scored> fuse(const std::vector<Scored>& dense_results, const std::vector<Scored>& sparse_results) const;
};

struct RetrievalResult {  // src/retrieve.rs:13-76
  Chunk chunk;
  std::optional<float> dense_score, sparse_score, fused_score, rerank_score;
  float best_score() const {
    if (rerank_score) return *rerank_score;
    if (fused_score) return *fused_score;
    if (dense_score) return *dense_score;
    if (sparse_score) return *sparse_score;
    return 0.0f;
  }
};

struct HybridRetrieverConfig {  // src/retrieve.rs:80-100
  size_t candidates_per_source = 50;
  FusionStrategy fusion;
  bool use_dense = true, use_sparse = true;
};

using Embedder = std::function<std::vector<float>(const std::string&)>;  // Embedder::embed_query, src/embed.rs:69-71

class HybridRetriever {  // src/retrieve.rs:103-263
 public:
  HybridRetriever(VectorStore dense, BM25Index sparse, Embedder embedder);
  HybridRetriever with_config(HybridRetrieverConfig config) &&;
  const VectorStore& dense_store() const { return dense_; }
  VectorStore& dense_store_mut() { aligned_ = false; return dense_; }
  const BM25Index& sparse_index() const { return sparse_; }
  BM25Index& sparse_index_mut() { aligned_ = false; return sparse_; }
  void index(Chunk chunk);                       // :156-164
  void index_batch(std::vector<Chunk> chunks);   // :167-172
  std::vector<RetrievalResult> retrieve(const std::string& query, size_t k) const;         // :175-220
  std::vector<RetrievalResult> retrieve_dense(const std::string& query, size_t k) const;   // :223-235
  std::vector<RetrievalResult> retrieve_sparse(const std::string& query, size_t k) const;  // :238-250
  size_t len() const { return dense_.len(); }
  bool is_empty() const { return dense_.is_empty(); }

 private:
  VectorStore dense_;
  BM25Index sparse_;
  Embedder embedder_;
  HybridRetrieverConfig config_;
  bool aligned_;  // dense and sparse ordinals coincide -> one fused device call
};

// process-wide device context (one GPU per process; device = $TRR_DEVICE or $LOCAL_RANK or 0)
trr_ctx* default_context();

}  // namespace trueno_rag
