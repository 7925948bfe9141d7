// host_mirror.cpp — implementation of include/trueno_rag.hpp (host-side mirror of the reference API).
// Only bookkeeping lives here (id maps, tokenizer, dictionary, CSR construction, idf via the platform logf);
// every score, ranking and fusion is computed by the CUDA kernels through the C ABI.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <random>

#include "../../../include/trueno_rag.hpp"

namespace trueno_rag {

// ------------------------------------------------------------------------------------------------
static void check(int status) {
  if (status == TRR_OK) return;
  const std::string msg = trr_last_error();
  switch (status) {
    case TRR_ERR_INVALID_ARG: throw Error(Error::Kind::InvalidConfig, msg);
    case TRR_ERR_UNSUPPORTED: throw Error(Error::Kind::Unsupported, msg);
    default: throw Error(Error::Kind::VectorStore, msg);  // src/error.rs:38-39
  }
}

trr_ctx* default_context() {
  static trr_ctx* ctx = nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (!ctx) {
    int dev = 0;
    if (const char* e = getenv("TRR_DEVICE")) dev = atoi(e);
    else if (const char* e2 = getenv("LOCAL_RANK")) dev = atoi(e2) % std::max(1, trr_device_count());
    check(trr_ctx_create(dev, &ctx));
  }
  return ctx;
}

ChunkId ChunkId::random() {
  static std::mt19937_64 rng{std::random_device{}()};
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  ChunkId c;
  c.hi = (rng() & 0xFFFFFFFFFFFF0FFFull) | 0x0000000000004000ull;  // version 4
  c.lo = (rng() & 0x3FFFFFFFFFFFFFFFull) | 0x8000000000000000ull;  // variant 1
  return c;
}

namespace detail {
struct DeviceDense {
  trr_dense* h = nullptr;
  ~DeviceDense() { if (h) trr_dense_destroy(h); }
};
struct DeviceBm25 {
  trr_bm25* h = nullptr;
  ~DeviceBm25() { if (h) trr_bm25_destroy(h); }
};
}  // namespace detail

// ================================================================================================
// VectorStore
// ================================================================================================
VectorStore::VectorStore(VectorStoreConfig config) : config_(config) {}
VectorStore VectorStore::with_dimension(size_t dimension) {
  VectorStoreConfig c;
  c.dimension = dimension;
  return VectorStore(c);
}
VectorStore::~VectorStore() = default;
VectorStore::VectorStore(VectorStore&&) noexcept = default;
VectorStore& VectorStore::operator=(VectorStore&&) noexcept = default;

VectorStore VectorStore::clone() const {
  VectorStore c(config_);
  // re-insert in ordinal order so the clone has the same canonical tie order
  for (const ChunkId& id : id_of_) {
    auto it = ord_of_.find(id);
    if (it == ord_of_.end() || id_of_[it->second] != id) continue;
    auto ch = chunks_.find(id);
    if (ch != chunks_.end()) c.insert(ch->second);
  }
  return c;
}

void VectorStore::insert(Chunk chunk) {
  if (!chunk.embedding) throw Error(Error::Kind::InvalidConfig, "chunk must have embedding");  // :360-363
  if (chunk.embedding->size() != config_.dimension)                                            // :365-370
    throw Error(Error::Kind::DimensionMismatch, "dimension mismatch", config_.dimension, chunk.embedding->size());
  auto it = ord_of_.find(chunk.id);
  if (it != ord_of_.end()) {
    // HashMap::insert replaces the vector of an existing id: tombstone the old row, append the new one
    flush();
    check(trr_dense_remove(dev_->h, it->second));
  }
  const uint32_t ord = (uint32_t)id_of_.size();
  id_of_.push_back(chunk.id);
  ord_of_[chunk.id] = ord;
  pending_.insert(pending_.end(), chunk.embedding->begin(), chunk.embedding->end());
  chunks_[chunk.id] = std::move(chunk);
}

void VectorStore::insert_batch(std::vector<Chunk> chunks) {
  for (auto& c : chunks) insert(std::move(c));
}

void VectorStore::flush() const {
  if (!dev_) {
    dev_ = std::make_shared<detail::DeviceDense>();
    check(trr_dense_create(default_context(), (uint32_t)config_.dimension, (int)config_.metric, config_.storage_dtype, 0,
                           &dev_->h));
  }
  if (!pending_.empty()) {
    check(trr_dense_append(dev_->h, pending_.data(), pending_.size() / config_.dimension));
    pending_.clear();
    pending_.shrink_to_fit();
  }
}

trr_dense* VectorStore::device_handle() const {
  flush();
  return dev_->h;
}

void VectorStore::set_mode(int mode) {
  flush();
  check(trr_dense_set_mode(dev_->h, mode));
}

bool VectorStore::ordinal_of(const ChunkId& id, uint32_t* out) const {
  auto it = ord_of_.find(id);
  if (it == ord_of_.end()) return false;
  *out = it->second;
  return true;
}

std::vector<std::vector<Scored>> VectorStore::search_batch(const std::vector<float>& queries, size_t B, size_t k) const {
  if (queries.size() != B * config_.dimension)  // :387-392
    throw Error(Error::Kind::DimensionMismatch, "dimension mismatch", config_.dimension, B ? queries.size() / B : 0);
  std::vector<std::vector<Scored>> out(B);
  if (B == 0 || k == 0 || ord_of_.empty()) return out;
  flush();
  const size_t kk = std::min(k, ord_of_.size());  // truncate(k) of at most len() entries
  std::vector<uint32_t> ord(B * kk), n(B);
  std::vector<float> sc(B * kk);
  check(trr_dense_search(dev_->h, queries.data(), (uint32_t)B, (uint32_t)kk, ord.data(), sc.data(), n.data()));
  for (size_t b = 0; b < B; ++b) {
    out[b].reserve(n[b]);
    for (uint32_t i = 0; i < n[b]; ++i) out[b].emplace_back(id_of_[ord[b * kk + i]], sc[b * kk + i]);
  }
  return out;
}

std::vector<Scored> VectorStore::search(const std::vector<float>& q, size_t k) const {
  if (q.size() != config_.dimension)
    throw Error(Error::Kind::DimensionMismatch, "dimension mismatch", config_.dimension, q.size());
  return search_batch(q, 1, k)[0];
}

const Chunk* VectorStore::get(const ChunkId& id) const {
  auto it = chunks_.find(id);
  return it == chunks_.end() ? nullptr : &it->second;
}

std::optional<Chunk> VectorStore::remove(const ChunkId& id) {
  auto it = ord_of_.find(id);
  if (it == ord_of_.end()) return std::nullopt;
  flush();
  check(trr_dense_remove(dev_->h, it->second));
  ord_of_.erase(it);
  auto ch = chunks_.find(id);
  std::optional<Chunk> out;
  if (ch != chunks_.end()) { out = std::move(ch->second); chunks_.erase(ch); }
  return out;
}

// ================================================================================================
// BM25Index
// ================================================================================================
static const char* kStopwords[] = {  // src/index.rs:93-108
    "a", "an", "the", "is", "are", "was", "were", "be", "been", "being", "have", "has", "had", "do", "does", "did",
    "will", "would", "could", "should", "may", "might", "must", "shall", "can", "need", "dare", "ought", "used", "to",
    "of", "in", "for", "on", "with", "at", "by", "from", "as", "into", "through", "during", "before", "after", "above",
    "below", "between", "under", "again", "further", "then", "once", "here", "there", "when", "where", "why", "how",
    "all", "each", "few", "more", "most", "other", "some", "such", "no", "nor", "not", "only", "own", "same", "so",
    "than", "too", "very", "just", "and", "but", "if", "or", "because", "until", "while", "this", "that", "these",
    "those", "it", "its"};

BM25Index::BM25Index() {
  for (const char* s : kStopwords) stopwords_.insert(s);
}
BM25Index BM25Index::with_params(float k1, float b) {
  BM25Index ix;
  ix.k1_ = k1;
  ix.b_ = b;
  return ix;
}
BM25Index BM25Index::with_stopwords(std::unordered_set<std::string> stopwords) && {
  stopwords_ = std::move(stopwords);
  return std::move(*this);
}
BM25Index::~BM25Index() = default;
BM25Index::BM25Index(BM25Index&&) noexcept = default;
BM25Index& BM25Index::operator=(BM25Index&&) noexcept = default;

// --- tokenizer: split on !char::is_alphanumeric, lowercase, drop stopwords, drop tokens with byte length < 2 ---
// ASCII is exact.  Outside ASCII, Unicode `Alphabetic || Numeric` and `to_lowercase` are approximated by range
// tables that cover Latin-1, Latin Extended-A, Greek, Cyrillic and the common punctuation / symbol blocks.
static void decode_utf8(const std::string& s, size_t& i, uint32_t& cp) {
  const unsigned char c = (unsigned char)s[i];
  const int n = c < 0x80 ? 0 : (c >> 5) == 0x6 ? 1 : (c >> 4) == 0xE ? 2 : (c >> 3) == 0x1E ? 3 : -1;
  if (n < 0 || i + (size_t)n >= s.size()) {  // stray continuation byte or truncated sequence
    cp = 0xFFFD;
    i += 1;
    return;
  }
  cp = n == 0 ? c : (c & (0x3F >> n));
  for (int k = 1; k <= n; ++k) cp = (cp << 6) | ((unsigned char)s[i + k] & 0x3F);
  i += (size_t)n + 1;
}
static void encode_utf8(uint32_t cp, std::string& out) {
  if (cp < 0x80) out.push_back((char)cp);
  else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
  else if (cp < 0x10000) {
    out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
    out.push_back((char)(0x80 | (cp & 0x3F)));
  } else {
    out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
    out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F)));
  }
}
static bool is_alphanumeric(uint32_t c) {
  if (c < 0x80) return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z');
  if (c < 0xC0) return c == 0xAA || c == 0xB5 || c == 0xBA || c == 0xB2 || c == 0xB3 || c == 0xB9 || (c >= 0xBC && c <= 0xBE);
  if (c == 0xD7 || c == 0xF7) return false;
  if (c >= 0x2000 && c <= 0x206F) return false;   // general punctuation
  if (c >= 0x20A0 && c <= 0x20CF) return false;   // currency
  if (c >= 0x2190 && c <= 0x245F) return false;   // arrows, math, technical
  if (c >= 0x2500 && c <= 0x27BF) return false;   // box drawing, shapes, dingbats
  if (c >= 0x3000 && c <= 0x3004) return false;   // CJK punctuation
  if (c >= 0x3008 && c <= 0x3020) return false;
  if (c >= 0xFF00 && c <= 0xFF0F) return false;   // full-width punctuation
  if (c >= 0xFF1A && c <= 0xFF20) return false;
  if (c >= 0xE000 && c <= 0xF8FF) return false;   // private use
  if (c == 0xFFFD) return false;
  return true;
}
static uint32_t to_lower(uint32_t c) {
  if (c >= 'A' && c <= 'Z') return c + 32;
  if (c < 0x80) return c;
  if (c >= 0xC0 && c <= 0xDE && c != 0xD7) return c + 32;
  if (c >= 0x100 && c <= 0x17F) {
    if ((c >= 0x139 && c <= 0x148) || (c >= 0x179 && c <= 0x17E)) return (c & 1) ? c + 1 : c;
    if (c == 0x130 || c == 0x178) return c == 0x178 ? 0xFF : c;
    return (c & 1) ? c : c + 1;
  }
  if (c >= 0x391 && c <= 0x3A9 && c != 0x3A2) return c + 32;
  if (c >= 0x410 && c <= 0x42F) return c + 32;
  if (c >= 0x400 && c <= 0x40F) return c + 80;
  return c;
}

std::vector<std::string> BM25Index::tokenize(const std::string& text) const {
  std::vector<std::string> out;
  std::string cur;
  auto emit = [&]() {
    if (cur.empty()) return;
    if (!stopwords_.count(cur) && cur.size() >= 2) out.push_back(cur);  // :121-122
    cur.clear();
  };
  size_t i = 0;
  while (i < text.size()) {
    uint32_t cp;
    decode_utf8(text, i, cp);
    if (is_alphanumeric(cp)) encode_utf8(lowercase_ ? to_lower(cp) : cp, cur);
    else emit();
  }
  emit();
  return out;
}

void BM25Index::add(const Chunk& chunk) {  // :176-204
  const std::vector<std::string> tokens = tokenize(chunk.content);
  uint32_t ord;
  auto it = ord_of_.find(chunk.id);
  if (it != ord_of_.end()) {
    // the reference would append a second posting for the same id; keep its observable effect on the counters
    // (doc_count += 1, doc_lengths overwritten) but give the re-added chunk a fresh ordinal
    ord = (uint32_t)id_of_.size();
  } else {
    ord = (uint32_t)id_of_.size();
  }
  id_of_.push_back(chunk.id);
  ord_of_[chunk.id] = ord;
  doc_len_.push_back((uint32_t)tokens.size());
  live_.push_back(1);
  doc_count_ += 1;
  // term frequencies of this document, then one posting per distinct term (:185-201)
  std::vector<uint32_t> ids;
  ids.reserve(tokens.size());
  for (const std::string& t : tokens) {
    auto d = dict_.find(t);
    uint32_t tid;
    if (d == dict_.end()) {
      tid = (uint32_t)postings_.size();
      dict_.emplace(t, tid);
      postings_.emplace_back();
      df_.push_back(0);
    } else {
      tid = d->second;
    }
    ids.push_back(tid);
  }
  std::sort(ids.begin(), ids.end());
  for (size_t a = 0; a < ids.size();) {
    size_t e = a;
    while (e < ids.size() && ids[e] == ids[a]) ++e;
    postings_[ids[a]].emplace_back(ord, (uint32_t)(e - a));
    df_[ids[a]] += 1;
    a = e;
  }
  dirty_ = true;
}

void BM25Index::add_batch(const std::vector<Chunk>& chunks) {
  for (const Chunk& c : chunks) add(c);
}

void BM25Index::remove(const ChunkId& id) {  // :245-275
  auto it = ord_of_.find(id);
  if (it == ord_of_.end()) return;
  const uint32_t ord = it->second;
  ord_of_.erase(it);
  if (live_[ord]) {
    live_[ord] = 0;
    doc_len_[ord] = 0;
    doc_count_ = doc_count_ ? doc_count_ - 1 : 0;
  }
  for (size_t t = 0; t < postings_.size(); ++t) {
    auto& pl = postings_[t];
    for (size_t i = 0; i < pl.size(); ++i) {
      if (pl[i].first != ord) continue;
      if (t < frozen_len_.size() && i < frozen_len_[t]) frozen_len_[t] -= 1;  // the device copy keeps it as a dead posting
      pl.erase(pl.begin() + (long)i);
      if (df_[t] > 0) df_[t] -= 1;  // a term whose df reaches 0 keeps an empty list
      break;                        // a document has at most one posting per term
    }
  }
  if (ord < frozen_docs_) pending_removed_.push_back(ord);  // otherwise it never reached the device
  dirty_ = true;
}

float BM25Index::avg_doc_length() const {
  freeze();
  return avg_doc_length_;
}

void BM25Index::freeze() const {
  if (!dirty_) return;
  // :157-164 — u32 (wrapping) sum of the live lengths, as f32 / count as f32
  uint32_t total = 0;
  for (size_t i = 0; i < doc_len_.size(); ++i) if (live_[i]) total += doc_len_[i];
  avg_doc_length_ = doc_count_ == 0 ? 0.0f : (float)total / (float)doc_count_;
  const uint32_t n_terms = (uint32_t)postings_.size();
  std::vector<float> idf(n_terms);
  const float n = (float)doc_count_;
  for (uint32_t t = 0; t < n_terms; ++t) {
    const float df = (float)df_[t];
    idf[t] = logf((n - df + 0.5f) / (df + 0.5f) + 1.0f);  // :147, platform logf == Rust f32::ln here
  }
  const uint32_t n_docs = (uint32_t)doc_len_.size();
  if (dev_ && dev_->h && !needs_rebuild_ && n_docs >= frozen_docs_ && !pending_removed_.empty()) {
    // removes since the last freeze: tombstone their postings on the device and re-weight (src/index.rs:245-275)
    uint64_t dead = 0;
    const bool only_removes = n_docs == frozen_docs_;
    check(trr_bm25_remove(dev_->h, pending_removed_.data(), (uint32_t)pending_removed_.size(), avg_doc_length_, k1_, b_,
                          idf.data(), &dead));
    pending_removed_.clear();
    if (dead * 4 > frozen_postings_) needs_rebuild_ = true;  // a quarter of the device postings is dead: compact by rebuilding
    else if (only_removes) { dirty_ = false; return; }
  }
  if (dev_ && dev_->h && !needs_rebuild_ && n_docs >= frozen_docs_) {
    // only adds since the last freeze: ship the CSR of the new documents, merge and re-weight on the device
    frozen_len_.resize(n_terms, 0);
    std::vector<uint64_t> d_off(n_terms + 1, 0);
    for (uint32_t t = 0; t < n_terms; ++t) d_off[t + 1] = d_off[t] + (postings_[t].size() - frozen_len_[t]);
    std::vector<uint32_t> pd(d_off[n_terms]), ptf(d_off[n_terms]);
    for (uint32_t t = 0; t < n_terms; ++t) {
      uint64_t p = d_off[t];
      for (size_t i = frozen_len_[t]; i < postings_[t].size(); ++i, ++p) {
        pd[p] = postings_[t][i].first - frozen_docs_;
        ptf[p] = postings_[t][i].second;
      }
    }
    check(trr_bm25_append(dev_->h, n_docs - frozen_docs_, n_terms, d_off.data(), pd.data(), ptf.data(),
                          doc_len_.data() + frozen_docs_, avg_doc_length_, k1_, b_, idf.data()));
  } else {
    std::vector<uint64_t> term_off(n_terms + 1, 0);
    for (uint32_t t = 0; t < n_terms; ++t) term_off[t + 1] = term_off[t] + postings_[t].size();
    std::vector<uint32_t> pd(term_off[n_terms]), ptf(term_off[n_terms]);
    for (uint32_t t = 0; t < n_terms; ++t) {
      uint64_t p = term_off[t];
      for (const auto& e : postings_[t]) { pd[p] = e.first; ptf[p] = e.second; ++p; }
    }
    dev_ = std::make_shared<detail::DeviceBm25>();
    check(trr_bm25_build(default_context(), n_docs, n_terms, term_off.data(), pd.data(), ptf.data(), doc_len_.data(),
                         avg_doc_length_, k1_, b_, idf.data(), 0, &dev_->h));
  }
  frozen_docs_ = n_docs;
  frozen_len_.resize(n_terms);
  for (uint32_t t = 0; t < n_terms; ++t) frozen_len_[t] = (uint32_t)postings_[t].size();
  pending_removed_.clear();
  check(trr_bm25_n_postings(dev_->h, &frozen_postings_));
  needs_rebuild_ = false;
  dirty_ = false;
}

trr_bm25* BM25Index::device_handle() const {
  freeze();
  return dev_->h;
}

std::vector<uint32_t> BM25Index::term_ids(const std::vector<std::string>& tokens) const {
  std::vector<uint32_t> ids;
  ids.reserve(tokens.size());
  for (const std::string& t : tokens) {
    auto d = dict_.find(t);
    ids.push_back(d == dict_.end() ? 0xFFFFFFFFu : d->second);
  }
  return ids;
}

std::vector<Scored> BM25Index::search(const std::string& query, size_t k) const {
  const std::vector<std::string> terms = tokenize(query);
  std::vector<Scored> out;
  if (terms.empty() || k == 0 || doc_count_ == 0) return out;  // :213-216
  freeze();
  const std::vector<uint32_t> ids = term_ids(terms);
  const uint32_t off[2] = {0, (uint32_t)ids.size()};
  const size_t kk = std::min<size_t>(k, id_of_.size());
  std::vector<uint32_t> ord(kk);
  std::vector<float> sc(kk);
  uint32_t n = 0;
  check(trr_bm25_search(dev_->h, ids.data(), off, 1, (uint32_t)kk, ord.data(), sc.data(), &n));
  out.reserve(n);
  for (uint32_t i = 0; i < n; ++i) out.emplace_back(id_of_[ord[i]], sc[i]);
  return out;
}

// ================================================================================================
// FusionStrategy
// ================================================================================================
std::vector<Scored> FusionStrategy::fuse(const std::vector<Scored>& dense, const std::vector<Scored>& sparse) const {
  // temporary id space: distinct ChunkIds numbered by first appearance (dense list, then sparse list)
  std::unordered_map<ChunkId, uint32_t, ChunkIdHash> num;
  std::vector<ChunkId> ids;
  auto number = [&](const ChunkId& c) {
    auto it = num.find(c);
    if (it != num.end()) return it->second;
    const uint32_t v = (uint32_t)ids.size();
    num.emplace(c, v);
    ids.push_back(c);
    return v;
  };
  const uint32_t C = (uint32_t)std::max<size_t>(1, std::max(dense.size(), sparse.size()));
  std::vector<uint32_t> d_ord(C), s_ord(C);
  std::vector<float> d_sc(C), s_sc(C);
  for (size_t i = 0; i < dense.size(); ++i) { d_ord[i] = number(dense[i].first); d_sc[i] = dense[i].second; }
  for (size_t i = 0; i < sparse.size(); ++i) { s_ord[i] = number(sparse[i].first); s_sc[i] = sparse[i].second; }
  const uint32_t nd = (uint32_t)dense.size(), ns = (uint32_t)sparse.size();
  std::vector<Scored> out;
  if (nd + ns == 0) return out;
  const uint32_t k_out = nd + ns;
  std::vector<uint32_t> o_ord(k_out);
  std::vector<float> o_f(k_out);
  uint32_t n = 0;
  check(trr_fuse(default_context(), (int)kind, param, d_ord.data(), d_sc.data(), &nd, s_ord.data(), s_sc.data(), &ns, 1, C,
                 k_out, o_ord.data(), o_f.data(), nullptr, nullptr, &n));
  out.reserve(n);
  for (uint32_t i = 0; i < n; ++i) out.emplace_back(ids[o_ord[i]], o_f[i]);
  return out;
}

// ================================================================================================
// HybridRetriever
// ================================================================================================
HybridRetriever::HybridRetriever(VectorStore dense, BM25Index sparse, Embedder embedder)
    : dense_(std::move(dense)), sparse_(std::move(sparse)), embedder_(std::move(embedder)),
      aligned_(dense_.next_ordinal() == 0 && sparse_.next_ordinal() == 0) {}

HybridRetriever HybridRetriever::with_config(HybridRetrieverConfig config) && {
  config_ = config;
  return std::move(*this);
}

void HybridRetriever::index(Chunk chunk) {  // :156-164
  sparse_.add(chunk);
  const bool same = dense_.next_ordinal() + 1 == sparse_.next_ordinal();
  dense_.insert(std::move(chunk));  // throws on a missing / mis-sized embedding AFTER the sparse add, as in the reference
  aligned_ = aligned_ && same;
}

void HybridRetriever::index_batch(std::vector<Chunk> chunks) {
  for (auto& c : chunks) index(std::move(c));
}

std::vector<RetrievalResult> HybridRetriever::retrieve(const std::string& query, size_t k) const {
  const size_t C = config_.candidates_per_source;
  std::vector<RetrievalResult> results;
  if (aligned_ && config_.use_dense && config_.use_sparse && C > 0 && k > 0 && !dense_.is_empty()) {
    // fused device path: dense top-C + sparse top-C + fusion + take(k) in one C-ABI call
    const std::vector<float> q = embedder_(query);
    if (q.size() != dense_.config().dimension)
      throw Error(Error::Kind::DimensionMismatch, "dimension mismatch", dense_.config().dimension, q.size());
    const std::vector<uint32_t> terms = sparse_.term_ids(sparse_.tokenize(query));
    const uint32_t off[2] = {0, (uint32_t)terms.size()};
    const uint32_t kk = (uint32_t)std::min<size_t>(k, 2 * C);
    std::vector<uint32_t> ord(kk);
    std::vector<float> f(kk), d(kk), s(kk);
    uint32_t n = 0;
    check(trr_hybrid_search(dense_.device_handle(), sparse_.device_handle(), q.data(), terms.data(), off, 1, (uint32_t)C,
                            (int)config_.fusion.kind, config_.fusion.param, kk, 1, terms.empty() ? 0 : 1, ord.data(),
                            f.data(), d.data(), s.data(), &n));
    for (uint32_t i = 0; i < n; ++i) {
      const Chunk* ch = dense_.get(dense_.id_of(ord[i]));
      if (!ch) continue;  // :205
      RetrievalResult r{*ch, std::nullopt, std::nullopt, f[i], std::nullopt};
      if (!isnan(d[i])) r.dense_score = d[i];
      if (!isnan(s[i])) r.sparse_score = s[i];
      results.push_back(std::move(r));
    }
    return results;
  }
  // general path: independent id spaces -> two device searches + device fusion over ChunkIds
  std::vector<Scored> dense_results, sparse_results;
  if (config_.use_dense) dense_results = dense_.search(embedder_(query), C);
  if (config_.use_sparse) sparse_results = sparse_.search(query, C);
  const std::vector<Scored> fused = config_.fusion.fuse(dense_results, sparse_results);
  std::unordered_map<ChunkId, float, ChunkIdHash> dmap, smap;
  for (const auto& e : dense_results) dmap[e.first] = e.second;
  for (const auto& e : sparse_results) smap[e.first] = e.second;
  for (size_t i = 0; i < fused.size() && i < k; ++i) {
    const Chunk* ch = dense_.get(fused[i].first);
    if (!ch) continue;
    RetrievalResult r{*ch, std::nullopt, std::nullopt, fused[i].second, std::nullopt};
    auto di = dmap.find(fused[i].first);
    if (di != dmap.end()) r.dense_score = di->second;
    auto si = smap.find(fused[i].first);
    if (si != smap.end()) r.sparse_score = si->second;
    results.push_back(std::move(r));
  }
  return results;
}

std::vector<RetrievalResult> HybridRetriever::retrieve_dense(const std::string& query, size_t k) const {
  std::vector<RetrievalResult> out;
  for (const auto& e : dense_.search(embedder_(query), k)) {
    const Chunk* ch = dense_.get(e.first);
    if (ch) out.push_back(RetrievalResult{*ch, e.second, std::nullopt, std::nullopt, std::nullopt});
  }
  return out;
}

std::vector<RetrievalResult> HybridRetriever::retrieve_sparse(const std::string& query, size_t k) const {
  std::vector<RetrievalResult> out;
  for (const auto& e : sparse_.search(query, k)) {
    const Chunk* ch = dense_.get(e.first);  // chunks come from the DENSE store, as in the reference (:243)
    if (ch) out.push_back(RetrievalResult{*ch, std::nullopt, e.second, std::nullopt, std::nullopt});
  }
  return out;
}

}  // namespace trueno_rag
