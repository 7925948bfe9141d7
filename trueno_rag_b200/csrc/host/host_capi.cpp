// host_capi.cpp — flat C wrappers over the C++ host mirror (see include/trueno_rag_host.h).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../../include/trueno_rag.hpp"
#include "../../../include/trueno_rag_host.h"

using namespace trueno_rag;

static thread_local std::string g_err;
static thread_local uint64_t g_expected = 0, g_actual = 0;

template <typename F>
static int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const Error& e) {
    g_err = e.what(); g_expected = e.expected; g_actual = e.actual;
    switch (e.kind) {
      case Error::Kind::InvalidConfig: return 1;
      case Error::Kind::DimensionMismatch: return 2;
      case Error::Kind::Unsupported: return 6;
      case Error::Kind::Serialization: return 7;
      default: return 3;
    }
  } catch (const std::exception& e) {
    g_err = e.what();
    return 3;
  }
}

struct trrh_store { VectorStore v; };
struct trrh_bm25 { BM25Index v; };
struct trrh_retriever {
  std::vector<float> cur_q;
  std::unique_ptr<HybridRetriever> r;
};

static ChunkId cid(trrh_id i) { ChunkId c; c.hi = i.hi; c.lo = i.lo; return c; }
static trrh_id tid(const ChunkId& c) { return trrh_id{c.hi, c.lo}; }
static Chunk make_chunk(trrh_id id, const char* content, const float* emb, uint32_t emb_len, int has_emb) {
  Chunk c(content ? content : "", 0, content ? strlen(content) : 0);
  c.id = cid(id);
  if (has_emb) c.set_embedding(std::vector<float>(emb, emb + emb_len));
  return c;
}

extern "C" {
const char* trrh_last_error(void) { return g_err.c_str(); }
uint64_t trrh_last_expected(void) { return g_expected; }
uint64_t trrh_last_actual(void) { return g_actual; }

int trrh_store_new(uint32_t dim, int metric, int dtype, trrh_store** out) {
  return guarded([&] {
    VectorStoreConfig c;
    c.dimension = dim; c.metric = (DistanceMetric)metric; c.storage_dtype = dtype;
    *out = new trrh_store{VectorStore(c)};
  });
}
void trrh_store_free(trrh_store* s) { delete s; }
int trrh_store_insert(trrh_store* s, trrh_id id, const char* content, const float* emb, uint32_t emb_len, int has_emb) {
  return guarded([&] { s->v.insert(make_chunk(id, content, emb, emb_len, has_emb)); });
}
int trrh_store_search(trrh_store* s, const float* q, uint32_t q_len, uint32_t k, trrh_id* out_ids, float* out_scores,
                      uint32_t* out_n) {
  return guarded([&] {
    auto r = s->v.search(std::vector<float>(q, q + q_len), k);
    *out_n = (uint32_t)r.size();
    for (size_t i = 0; i < r.size(); ++i) { out_ids[i] = tid(r[i].first); out_scores[i] = r[i].second; }
  });
}
int trrh_store_get(trrh_store* s, trrh_id id, const char** out_content) {
  const Chunk* c = s->v.get(cid(id));
  if (!c) return 0;
  if (out_content) *out_content = c->content.c_str();
  return 1;
}
int trrh_store_remove(trrh_store* s, trrh_id id) {
  int found = 0;
  int st = guarded([&] { found = s->v.remove(cid(id)).has_value() ? 1 : 0; });
  return st == 0 ? found : -st;
}
uint64_t trrh_store_len(trrh_store* s) { return s->v.len(); }
int trrh_store_set_mode(trrh_store* s, int mode) { return guarded([&] { s->v.set_mode(mode); }); }
int trrh_store_clone(trrh_store* s, trrh_store** out) { return guarded([&] { *out = new trrh_store{s->v.clone()}; }); }

int trrh_bm25_new(float k1, float b, trrh_bm25** out) {
  return guarded([&] { *out = new trrh_bm25{BM25Index::with_params(k1, b)}; });
}
void trrh_bm25_free(trrh_bm25* s) { delete s; }
int trrh_bm25_tokenize(trrh_bm25* s, const char* text, char* out, uint32_t cap, uint32_t* out_len) {
  return guarded([&] {
    std::string joined;
    for (const auto& t : s->v.tokenize(text)) { if (!joined.empty()) joined.push_back('\n'); joined += t; }
    *out_len = (uint32_t)joined.size();
    if (cap) { const size_t n = std::min<size_t>(cap - 1, joined.size()); memcpy(out, joined.data(), n); out[n] = 0; }
  });
}
int trrh_bm25_add(trrh_bm25* s, trrh_id id, const char* content) {
  return guarded([&] { s->v.add(make_chunk(id, content, nullptr, 0, 0)); });
}
int trrh_bm25_search(trrh_bm25* s, const char* query, uint32_t k, trrh_id* out_ids, float* out_scores, uint32_t* out_n) {
  return guarded([&] {
    auto r = s->v.search(query, k);
    *out_n = (uint32_t)r.size();
    for (size_t i = 0; i < r.size(); ++i) { out_ids[i] = tid(r[i].first); out_scores[i] = r[i].second; }
  });
}
int trrh_bm25_remove(trrh_bm25* s, trrh_id id) { return guarded([&] { s->v.remove(cid(id)); }); }
uint64_t trrh_bm25_len(trrh_bm25* s) { return s->v.len(); }
float trrh_bm25_avgdl(trrh_bm25* s) { float v = 0; guarded([&] { v = s->v.avg_doc_length(); }); return v; }
float trrh_bm25_k1(trrh_bm25* s) { return s->v.k1(); }
float trrh_bm25_b(trrh_bm25* s) { return s->v.b(); }
int trrh_bm25_contains_term(trrh_bm25* s, const char* term) { return s->v.contains_term(term) ? 1 : 0; }

static void bytes_out(const std::vector<uint8_t>& v, uint8_t** out, uint64_t* out_n) {
  uint8_t* p = static_cast<uint8_t*>(malloc(v.size() ? v.size() : 1));
  if (!p) throw Error(Error::Kind::VectorStore, "out of host memory");
  if (!v.empty()) memcpy(p, v.data(), v.size());
  *out = p;
  *out_n = v.size();
}
int trrh_compress(int compression, const uint8_t* data, uint64_t n, uint8_t** out, uint64_t* out_n) {
  return guarded([&] { bytes_out(compress((Compression)compression, data, n), out, out_n); });
}
int trrh_decompress(int compression, const uint8_t* data, uint64_t n, uint8_t** out, uint64_t* out_n) {
  return guarded([&] { bytes_out(decompress((Compression)compression, data, n), out, out_n); });
}
void trrh_bytes_free(uint8_t* p) { free(p); }
int trrh_bm25_to_bytes(trrh_bm25* s, int compression, uint8_t** out, uint64_t* out_n) {
  return guarded([&] {
    bytes_out(compression < 0 ? s->v.to_bytes() : s->v.to_compressed_bytes((Compression)compression), out, out_n);
  });
}
int trrh_bm25_from_bytes(const uint8_t* data, uint64_t n, int compression, trrh_bm25** out) {
  return guarded([&] {
    *out = new trrh_bm25{compression < 0 ? BM25Index::from_bytes(data, n)
                                         : BM25Index::from_compressed_bytes(data, n, (Compression)compression)};
  });
}
struct trrh_cli_index { PersistedIndex v; };
int trrh_cli_index_from_json(const char* text, uint64_t n, trrh_cli_index** out) {
  return guarded([&] { *out = new trrh_cli_index{PersistedIndex::from_json(text, n)}; });
}
void trrh_cli_index_free(trrh_cli_index* h) { delete h; }
int trrh_cli_index_new(uint64_t dimension, const char* embedder_type, const char* model_name, trrh_cli_index** out) {
  return guarded([&] {
    PersistedIndex p;
    p.dimension = dimension;
    p.embedder_type = embedder_type ? embedder_type : "";
    if (model_name) p.model_name = std::string(model_name);
    *out = new trrh_cli_index{std::move(p)};
  });
}
int trrh_cli_index_push(trrh_cli_index* h, const char* content, const char* title, const char* source,
                        const float* embedding, uint64_t len) {
  return guarded([&] {
    PersistedChunk c;
    c.content = content ? content : "";
    if (title) c.title = std::string(title);
    if (source) c.source = std::string(source);
    h->v.chunks.push_back(std::move(c));
    h->v.embeddings.emplace_back(embedding, embedding + len);
  });
}
int trrh_cli_index_to_json(trrh_cli_index* h, uint8_t** out, uint64_t* out_n) {
  return guarded([&] {
    const std::string j = h->v.to_json();
    bytes_out(std::vector<uint8_t>(j.begin(), j.end()), out, out_n);
  });
}
uint64_t trrh_cli_index_len(trrh_cli_index* h) { return h->v.chunks.size(); }
uint64_t trrh_cli_index_n_embeddings(trrh_cli_index* h) { return h->v.embeddings.size(); }
uint64_t trrh_cli_index_dimension(trrh_cli_index* h) { return h->v.dimension; }
const char* trrh_cli_index_embedder_type(trrh_cli_index* h) { return h->v.embedder_type.c_str(); }
const char* trrh_cli_index_model_name(trrh_cli_index* h) { return h->v.model_name ? h->v.model_name->c_str() : nullptr; }
int trrh_cli_index_chunk(trrh_cli_index* h, uint64_t i, const char** content, uint64_t* content_len, const char** title,
                         const char** source) {
  return guarded([&] {
    if (i >= h->v.chunks.size()) throw Error(Error::Kind::InvalidConfig, "chunk index out of range");
    const PersistedChunk& c = h->v.chunks[i];
    *content = c.content.data();
    *content_len = c.content.size();
    *title = c.title ? c.title->c_str() : nullptr;
    *source = c.source ? c.source->c_str() : nullptr;
  });
}
int trrh_cli_index_embedding(trrh_cli_index* h, uint64_t i, const float** data, uint64_t* len) {
  return guarded([&] {
    if (i >= h->v.embeddings.size()) throw Error(Error::Kind::InvalidConfig, "embedding index out of range");
    *data = h->v.embeddings[i].data();
    *len = h->v.embeddings[i].size();
  });
}
int trrh_cli_index_query(trrh_cli_index* h, const float* q, uint64_t q_len, uint64_t top_k, uint64_t* out_idx,
                         float* out_score, uint64_t* out_n) {
  return guarded([&] {
    const auto r = h->v.query(std::vector<float>(q, q + q_len), top_k);
    *out_n = r.size();
    for (size_t i = 0; i < r.size(); ++i) { out_idx[i] = r[i].first; out_score[i] = r[i].second; }
  });
}
int trrh_fuse(int kind, float param, const trrh_id* d_ids, const float* d_sc, uint32_t nd, const trrh_id* s_ids,
              const float* s_sc, uint32_t ns, trrh_id* out_ids, float* out_sc, uint32_t* out_n) {
  return guarded([&] {
    std::vector<Scored> d, sp;
    for (uint32_t i = 0; i < nd; ++i) d.emplace_back(cid(d_ids[i]), d_sc[i]);
    for (uint32_t i = 0; i < ns; ++i) sp.emplace_back(cid(s_ids[i]), s_sc[i]);
    FusionStrategy f{(FusionStrategy::Kind)kind, param};
    auto r = f.fuse(d, sp);
    *out_n = (uint32_t)r.size();
    for (size_t i = 0; i < r.size(); ++i) { out_ids[i] = tid(r[i].first); out_sc[i] = r[i].second; }
  });
}

int trrh_retriever_new(trrh_store* store, trrh_bm25* bm25, uint32_t C, int kind, float param, int use_dense,
                       int use_sparse, trrh_retriever** out) {
  return guarded([&] {
    auto* w = new trrh_retriever();
    HybridRetrieverConfig cfg;
    cfg.candidates_per_source = C; cfg.fusion = FusionStrategy{(FusionStrategy::Kind)kind, param};
    cfg.use_dense = use_dense != 0; cfg.use_sparse = use_sparse != 0;
    w->r.reset(new HybridRetriever(
        HybridRetriever(std::move(store->v), std::move(bm25->v), [w](const std::string&) { return w->cur_q; })
            .with_config(cfg)));
    delete store;
    delete bm25;
    *out = w;
  });
}
void trrh_retriever_free(trrh_retriever* r) { delete r; }
int trrh_retriever_index(trrh_retriever* r, trrh_id id, const char* content, const float* emb, uint32_t emb_len, int has_emb) {
  return guarded([&] { r->r->index(make_chunk(id, content, emb, emb_len, has_emb)); });
}
int trrh_retriever_retrieve(trrh_retriever* r, int which, const char* query, const float* q_emb, uint32_t q_len, uint32_t k,
                            trrh_id* out_ids, float* out_fused, float* out_dense, float* out_sparse, uint32_t* out_n) {
  return guarded([&] {
    r->cur_q.assign(q_emb, q_emb + q_len);
    std::vector<RetrievalResult> res = which == 0 ? r->r->retrieve(query, k)
                                     : which == 1 ? r->r->retrieve_dense(query, k)
                                                  : r->r->retrieve_sparse(query, k);
    *out_n = (uint32_t)res.size();
    for (size_t i = 0; i < res.size(); ++i) {
      out_ids[i] = tid(res[i].chunk.id);
      out_fused[i] = res[i].fused_score ? *res[i].fused_score : NAN;
      out_dense[i] = res[i].dense_score ? *res[i].dense_score : NAN;
      out_sparse[i] = res[i].sparse_score ? *res[i].sparse_score : NAN;
    }
  });
}
uint64_t trrh_retriever_len(trrh_retriever* r) { return r->r->len(); }
}
