// host_mirror.cpp — implementation of include/trueno_rag.hpp (host-side mirror of the reference API).
// Only bookkeeping lives here (id maps, tokenizer, dictionary, CSR construction, idf via the platform logf);
// every score, ranking and fusion is computed by the CUDA kernels through the C ABI.
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
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
// Unicode semantics come from generated UCD tables (unicode_tables.inc, tools/gen_unicode_tables.py): Alphabetic || N* for
// the split, the full lowercase mapping plus the Final_Sigma rule for the case fold.
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
#include "unicode_tables.inc"

template <size_t N>
static bool in_ranges(const uint32_t (&r)[N][2], uint32_t c) {
  size_t lo = 0, hi = N;  // first range whose end is >= c
  while (lo < hi) {
    const size_t mid = (lo + hi) >> 1;
    if (r[mid][1] < c) lo = mid + 1; else hi = mid;
  }
  return lo < N && r[lo][0] <= c;
}
// char::is_alphanumeric = Alphabetic || Nd || Nl || No
static bool is_alphanumeric(uint32_t c) {
  if (c < 0x80) return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'Z') || (c >= 'a' && c <= 'z');
  return in_ranges(TRR_UC_ALNUM, c);
}
static const TrrLowerEntry* lower_entry(uint32_t c) {
  size_t lo = 0, hi = sizeof(TRR_UC_LOWER) / sizeof(TRR_UC_LOWER[0]);
  while (lo < hi) {
    const size_t mid = (lo + hi) >> 1;
    if (TRR_UC_LOWER[mid].cp < c) lo = mid + 1; else hi = mid;
  }
  return lo < sizeof(TRR_UC_LOWER) / sizeof(TRR_UC_LOWER[0]) && TRR_UC_LOWER[lo].cp == c ? &TRR_UC_LOWER[lo] : nullptr;
}
// str::to_lowercase of one token (code points): full mapping, and U+03A3 becomes the final sigma when it is preceded by
// a cased letter and not followed by one, case-ignorable characters skipped (Final_Sigma, as library/alloc/src/str.rs)
static void lowercase_token(const std::vector<uint32_t>& tok, std::string& out) {
  auto ignorable_then_cased = [&](long from, long step) {
    for (long k = from; k >= 0 && k < (long)tok.size(); k += step) {
      if (in_ranges(TRR_UC_CASE_IGNORABLE, tok[(size_t)k])) continue;
      return in_ranges(TRR_UC_CASED, tok[(size_t)k]);
    }
    return false;
  };
  for (size_t i = 0; i < tok.size(); ++i) {
    const uint32_t c = tok[i];
    if (c < 0x80) { out.push_back((char)((c >= 'A' && c <= 'Z') ? c + 32 : c)); continue; }
    if (c == 0x3A3) {
      const bool final_sigma = ignorable_then_cased((long)i - 1, -1) && !ignorable_then_cased((long)i + 1, 1);
      encode_utf8(final_sigma ? 0x3C2 : 0x3C3, out);
      continue;
    }
    if (const TrrLowerEntry* e = lower_entry(c)) {
      for (uint32_t k = 0; k < e->n; ++k) encode_utf8(e->to[k], out);
    } else {
      encode_utf8(c, out);
    }
  }
}

std::vector<std::string> BM25Index::tokenize(const std::string& text) const {
  std::vector<std::string> out;
  std::vector<uint32_t> cur;
  std::string tok;
  auto emit = [&]() {
    if (cur.empty()) return;
    tok.clear();
    if (lowercase_) lowercase_token(cur, tok);
    else for (uint32_t c : cur) encode_utf8(c, tok);
    if (!stopwords_.count(tok) && tok.size() >= 2) out.push_back(tok);  // :121-122 (byte length)
    cur.clear();
  };
  size_t i = 0;
  while (i < text.size()) {
    uint32_t cp;
    decode_utf8(text, i, cp);
    if (is_alphanumeric(cp)) cur.push_back(cp);
    else emit();
  }
  emit();
  return out;
}

void BM25Index::add(const Chunk& chunk) {  // :176-204
  const std::vector<std::string> tokens = tokenize(chunk.content);
  // (a chunk id that is added twice gets a second ordinal: the reference appends a second posting for the same id and
  // counts the document twice, :193-203)
  const uint32_t ord = (uint32_t)id_of_.size();
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
  avg_loaded_.reset();  // update_avg_doc_length (:203)
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
  avg_loaded_.reset();  // update_avg_doc_length (:274)
  dirty_ = true;
}

static float avg_of(const std::vector<uint32_t>& doc_len, const std::vector<uint8_t>& live, uint32_t doc_count) {
  // :157-164 — u32 (wrapping) sum of the live lengths, as f32 / count as f32
  uint32_t total = 0;
  for (size_t i = 0; i < doc_len.size(); ++i) if (live[i]) total += doc_len[i];
  return doc_count == 0 ? 0.0f : (float)total / (float)doc_count;
}

float BM25Index::avg_doc_length() const {  // host only: no device work
  return avg_loaded_ ? *avg_loaded_ : avg_of(doc_len_, live_, doc_count_);
}

void BM25Index::freeze() const {
  if (!dirty_) return;
  avg_doc_length_ = avg_doc_length();
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
  // a token the index has never seen scores 0.0 for every document (src/index.rs:137-140) and adding +0.0 leaves an f32 sum
  // unchanged, so unknown tokens are dropped here instead of spending query-term slots of the device kernel on them
  for (const std::string& t : tokens) {
    auto d = dict_.find(t);
    if (d != dict_.end()) ids.push_back(d->second);
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
// Persistence in the reference's format: bincode 1.3 of BM25Index (src/index.rs:30-51), LZ4 (src/compressed.rs)
// ================================================================================================
const char* compression_as_str(Compression c) { return c == Compression::Lz4 ? "lz4" : "zstd"; }

// zstd_codec.cpp
std::vector<uint8_t> zstd_decompress(const uint8_t* src, size_t n);
std::vector<uint8_t> zstd_store(const uint8_t* src, size_t n);
std::vector<uint8_t> zstd_compress(const uint8_t* src, size_t n);

namespace {
[[noreturn]] void ser_fail(const std::string& m) { throw Error(Error::Kind::Serialization, m); }

inline uint32_t load32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

// One LZ4 block (the format lz4_flex writes): sequences of [token][literal length bytes][literals][offset u16]
// [match length bytes]; the block ends with a literals-only sequence, the last match starts at least 12 bytes and ends
// at least 5 bytes before the end of the input.
std::vector<uint8_t> lz4_compress_block(const uint8_t* src, size_t n) {
  std::vector<uint8_t> out;
  out.reserve(n + n / 255 + 16);
  auto put_len = [&](size_t v) {  // the part of a length beyond the 15 held by the token nibble
    while (v >= 255) { out.push_back(255); v -= 255; }
    out.push_back((uint8_t)v);
  };
  auto emit = [&](size_t lit_start, size_t lit_len, size_t match_len, size_t offset) {
    const size_t ml = match_len ? match_len - 4 : 0;
    out.push_back((uint8_t)((std::min<size_t>(lit_len, 15) << 4) | (match_len ? std::min<size_t>(ml, 15) : 0)));
    if (lit_len >= 15) put_len(lit_len - 15);
    out.insert(out.end(), src + lit_start, src + lit_start + lit_len);
    if (match_len) {
      out.push_back((uint8_t)(offset & 0xFF));
      out.push_back((uint8_t)(offset >> 8));
      if (ml >= 15) put_len(ml - 15);
    }
  };
  size_t anchor = 0;
  if (n > 12) {
    std::vector<int64_t> table((size_t)1 << 16, -1);
    const size_t match_start_limit = n - 12, match_end_limit = n - 5;
    size_t i = 0;
    while (i <= match_start_limit) {
      const uint32_t v = load32(src + i);
      const uint32_t h = (v * 2654435761u) >> 16;
      const int64_t cand = table[h];
      table[h] = (int64_t)i;
      if (cand >= 0 && i - (size_t)cand <= 65535 && load32(src + cand) == v) {
        size_t ml = 4;
        while (i + ml < match_end_limit && src[(size_t)cand + ml] == src[i + ml]) ++ml;
        emit(anchor, i - anchor, ml, i - (size_t)cand);
        i += ml;
        anchor = i;
      } else {
        ++i;
      }
    }
  }
  emit(anchor, n - anchor, 0, 0);
  return out;
}

std::vector<uint8_t> lz4_decompress_block(const uint8_t* src, size_t n, size_t out_size) {
  std::vector<uint8_t> out;
  out.reserve(out_size);
  size_t i = 0;
  auto get_len = [&](size_t base) {
    size_t v = base;
    if (base == 15) {
      uint8_t b;
      do {
        if (i >= n) ser_fail("LZ4 decompression failed: truncated length");
        b = src[i++];
        v += b;
      } while (b == 255);
    }
    return v;
  };
  while (i < n) {
    const uint8_t token = src[i++];
    const size_t lit = get_len(token >> 4);
    if (lit > n - i || out.size() + lit > out_size) ser_fail("LZ4 decompression failed: literals out of bounds");
    out.insert(out.end(), src + i, src + i + lit);
    i += lit;
    if (i == n) break;  // the last sequence has no match
    if (n - i < 2) ser_fail("LZ4 decompression failed: truncated offset");
    const size_t offset = (size_t)src[i] | ((size_t)src[i + 1] << 8);
    i += 2;
    const size_t ml = get_len(token & 15) + 4;
    if (offset == 0 || offset > out.size()) ser_fail("LZ4 decompression failed: offset out of bounds");
    if (out.size() + ml > out_size) ser_fail("LZ4 decompression failed: output too large");
    size_t from = out.size() - offset;
    for (size_t k = 0; k < ml; ++k) out.push_back(out[from + k]);  // may overlap its own output (run-length style)
  }
  if (out.size() != out_size) ser_fail("LZ4 decompression failed: size mismatch");
  return out;
}

struct BinWriter {
  std::vector<uint8_t> b;
  void raw(const void* p, size_t n) { const uint8_t* q = static_cast<const uint8_t*>(p); b.insert(b.end(), q, q + n); }
  void u8(uint8_t v) { b.push_back(v); }
  void u32(uint32_t v) { raw(&v, 4); }
  void u64(uint64_t v) { raw(&v, 8); }
  void f32(float v) { raw(&v, 4); }
  void str(const std::string& s) { u64(s.size()); raw(s.data(), s.size()); }
  void id(const ChunkId& c) {  // uuid::Uuid in a binary format: serialize_bytes(as_bytes()) = u64 length + 16 big-endian bytes
    u64(16);
    for (int k = 7; k >= 0; --k) u8((uint8_t)(c.hi >> (8 * k)));
    for (int k = 7; k >= 0; --k) u8((uint8_t)(c.lo >> (8 * k)));
  }
};

struct BinReader {
  const uint8_t* p;
  size_t n, i = 0;
  void need(size_t k) const { if (k > n - i) ser_fail("Bincode deserialization failed: unexpected end of input"); }
  uint8_t u8() { need(1); return p[i++]; }
  uint32_t u32() { need(4); uint32_t v; memcpy(&v, p + i, 4); i += 4; return v; }
  uint64_t u64() { need(8); uint64_t v; memcpy(&v, p + i, 8); i += 8; return v; }
  float f32() { need(4); float v; memcpy(&v, p + i, 4); i += 4; return v; }
  uint64_t len(size_t min_elem_bytes) {  // a length prefix that the remaining input can actually hold
    const uint64_t v = u64();
    if (min_elem_bytes && v > (n - i) / min_elem_bytes) ser_fail("Bincode deserialization failed: length exceeds input");
    return v;
  }
  std::string str() {
    const uint64_t l = len(1);
    std::string s(reinterpret_cast<const char*>(p + i), (size_t)l);
    i += (size_t)l;
    return s;
  }
  ChunkId id() {
    if (u64() != 16) ser_fail("Bincode deserialization failed: a ChunkId is 16 bytes");
    need(16);
    ChunkId c;
    for (int k = 0; k < 8; ++k) c.hi = (c.hi << 8) | p[i++];
    for (int k = 0; k < 8; ++k) c.lo = (c.lo << 8) | p[i++];
    return c;
  }
};
}  // namespace

std::vector<uint8_t> compress(Compression c, const uint8_t* data, size_t n) {
  if (n == 0) return {};  // :37-39
  if (c == Compression::Zstd) return zstd_compress(data, n);
  if (n > 0xFFFFFFFFull) ser_fail("LZ4 compression failed: input larger than 4 GiB");
  std::vector<uint8_t> block = lz4_compress_block(data, n);
  std::vector<uint8_t> out(4);
  const uint32_t sz = (uint32_t)n;
  memcpy(out.data(), &sz, 4);  // compress_prepend_size: u32 little-endian
  out.insert(out.end(), block.begin(), block.end());
  return out;
}

std::vector<uint8_t> decompress(Compression c, const uint8_t* data, size_t n) {
  if (n == 0) return {};  // :54-56
  if (c == Compression::Zstd) return zstd_decompress(data, n);
  if (n < 4) ser_fail("LZ4 decompression failed: missing size prefix");
  return lz4_decompress_block(data + 4, n - 4, load32(data));
}

std::vector<uint8_t> BM25Index::to_bytes() const {
  BinWriter w;
  std::vector<const std::string*> term_of(postings_.size(), nullptr);
  for (const auto& kv : dict_) term_of[kv.second] = &kv.first;
  uint64_t n_terms = 0;  // the reference drops a term when its df reaches 0 (:262-271); here it keeps an empty list
  for (size_t t = 0; t < postings_.size(); ++t) n_terms += !postings_[t].empty();
  // inverted_index: HashMap<String, Vec<(ChunkId, u32)>>
  w.u64(n_terms);
  for (size_t t = 0; t < postings_.size(); ++t) {
    if (postings_[t].empty()) continue;
    w.str(*term_of[t]);
    w.u64(postings_[t].size());
    for (const auto& e : postings_[t]) { w.id(id_of_[e.first]); w.u32(e.second); }
  }
  // doc_freqs: HashMap<String, u32>
  w.u64(n_terms);
  for (size_t t = 0; t < postings_.size(); ++t) {
    if (postings_[t].empty()) continue;
    w.str(*term_of[t]);
    w.u32(df_[t]);
  }
  // doc_lengths: HashMap<ChunkId, u32>
  uint64_t n_live = 0;
  for (size_t i = 0; i < live_.size(); ++i) n_live += ord_of_.count(id_of_[i]) && ord_of_.at(id_of_[i]) == i && live_[i];
  w.u64(n_live);
  for (size_t i = 0; i < live_.size(); ++i)
    if (live_[i] && ord_of_.count(id_of_[i]) && ord_of_.at(id_of_[i]) == i) { w.id(id_of_[i]); w.u32(doc_len_[i]); }
  w.f32(avg_doc_length());
  w.u32(doc_count_);
  w.f32(k1_);
  w.f32(b_);
  w.u8(lowercase_ ? 1 : 0);
  w.u64(stopwords_.size());
  for (const std::string& s : stopwords_) w.str(s);
  return std::move(w.b);
}

BM25Index BM25Index::from_bytes(const uint8_t* data, size_t n) {
  BinReader r{data, n};
  struct TermIn { std::string term; std::vector<std::pair<ChunkId, uint32_t>> postings; };
  std::vector<TermIn> terms((size_t)r.len(16));
  for (TermIn& t : terms) {
    t.term = r.str();
    t.postings.resize((size_t)r.len(28));
    for (auto& e : t.postings) { e.first = r.id(); e.second = r.u32(); }
  }
  std::unordered_map<std::string, uint32_t> doc_freqs;
  for (uint64_t k = r.len(12); k > 0; --k) { std::string t = r.str(); doc_freqs[std::move(t)] = r.u32(); }
  std::unordered_map<ChunkId, uint32_t, ChunkIdHash> doc_lengths;
  for (uint64_t k = r.len(28); k > 0; --k) { const ChunkId c = r.id(); doc_lengths[c] = r.u32(); }
  BM25Index ix;
  const float avg = r.f32();
  ix.doc_count_ = r.u32();
  ix.k1_ = r.f32();
  ix.b_ = r.f32();
  const uint8_t lc = r.u8();
  if (lc > 1) ser_fail("Bincode deserialization failed: invalid bool");
  ix.lowercase_ = lc != 0;
  ix.stopwords_.clear();
  for (uint64_t k = r.len(8); k > 0; --k) ix.stopwords_.insert(r.str());
  if (r.i != r.n) ser_fail("Bincode deserialization failed: trailing bytes");

  // chunks: every id of doc_lengths plus any id that only occurs in a posting (length 0, :144 unwrap_or(0)), numbered by
  // ascending ChunkId
  std::vector<ChunkId> ids;
  ids.reserve(doc_lengths.size());
  for (const auto& kv : doc_lengths) ids.push_back(kv.first);
  {
    std::unordered_set<ChunkId, ChunkIdHash> extra;
    for (const TermIn& t : terms)
      for (const auto& e : t.postings)
        if (!doc_lengths.count(e.first) && extra.insert(e.first).second) ids.push_back(e.first);
  }
  std::sort(ids.begin(), ids.end(), [](const ChunkId& a, const ChunkId& b) { return a.hi != b.hi ? a.hi < b.hi : a.lo < b.lo; });
  ix.id_of_ = ids;
  ix.doc_len_.resize(ids.size());
  ix.live_.assign(ids.size(), 1);
  for (size_t i = 0; i < ids.size(); ++i) {
    ix.ord_of_[ids[i]] = (uint32_t)i;
    auto it = doc_lengths.find(ids[i]);
    ix.doc_len_[i] = it == doc_lengths.end() ? 0u : it->second;
  }
  // terms by ascending string; postings by ascending ordinal, first occurrence of a chunk wins (:128-132 `find`)
  std::sort(terms.begin(), terms.end(), [](const TermIn& a, const TermIn& b) { return a.term < b.term; });
  for (size_t t = 0; t < terms.size(); ++t) {
    if (t && terms[t].term == terms[t - 1].term) ser_fail("Bincode deserialization failed: duplicate term");
    ix.dict_.emplace(terms[t].term, (uint32_t)t);
    std::vector<std::pair<uint32_t, uint32_t>> pl;
    pl.reserve(terms[t].postings.size());
    for (const auto& e : terms[t].postings) pl.emplace_back(ix.ord_of_.at(e.first), e.second);
    std::stable_sort(pl.begin(), pl.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    pl.erase(std::unique(pl.begin(), pl.end(), [](const auto& a, const auto& b) { return a.first == b.first; }), pl.end());
    ix.postings_.push_back(std::move(pl));
    auto df = doc_freqs.find(terms[t].term);
    ix.df_.push_back(df == doc_freqs.end() ? 0u : df->second);  // :140 unwrap_or(0)
  }
  // the reference scores with the stored average until the next add/remove recomputes it
  const float recomputed = avg_of(ix.doc_len_, ix.live_, ix.doc_count_);
  if (memcmp(&recomputed, &avg, 4) != 0) ix.avg_loaded_ = avg;
  ix.dirty_ = true;
  return ix;
}

std::vector<uint8_t> BM25Index::to_compressed_bytes(Compression c) const {
  const std::vector<uint8_t> b = to_bytes();
  return compress(c, b.data(), b.size());
}

BM25Index BM25Index::from_compressed_bytes(const uint8_t* data, size_t n, Compression c) {
  const std::vector<uint8_t> b = decompress(c, data, n);
  return from_bytes(b.data(), b.size());
}

// ================================================================================================
// PersistedIndex: the CLI's index.json (crates/trueno-rag-cli/src/main.rs:133-154, 437-439, 479-492)
// ================================================================================================
namespace {
// a small recursive-descent JSON reader: exactly what serde_json::from_str needs for PersistedIndex
struct Json {
  const char* p;
  const char* end;
  [[noreturn]] void fail(const char* what) const { ser_fail(std::string("JSON deserialization failed: ") + what); }
  void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
  bool peek(char c) { ws(); return p < end && *p == c; }
  void expect(char c) { ws(); if (p >= end || *p != c) fail("unexpected character"); ++p; }
  bool literal(const char* lit) {
    ws();
    const size_t l = strlen(lit);
    if ((size_t)(end - p) >= l && memcmp(p, lit, l) == 0) { p += l; return true; }
    return false;
  }
  static int hex(char c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1; }
  uint32_t hex4() {
    if (end - p < 4) fail("truncated \\u escape");
    uint32_t v = 0;
    for (int k = 0; k < 4; ++k) { const int h = hex(p[k]); if (h < 0) fail("bad \\u escape"); v = v * 16 + (uint32_t)h; }
    p += 4;
    return v;
  }
  std::string string() {
    expect('"');
    std::string out;
    while (true) {
      if (p >= end) fail("unterminated string");
      const unsigned char c = (unsigned char)*p++;
      if (c == '"') break;
      if (c < 0x20) fail("control character in string");
      if (c != '\\') { out.push_back((char)c); continue; }
      if (p >= end) fail("unterminated escape");
      const char e = *p++;
      switch (e) {
        case '"': out.push_back('"'); break;
        case '\\': out.push_back('\\'); break;
        case '/': out.push_back('/'); break;
        case 'b': out.push_back('\b'); break;
        case 'f': out.push_back('\f'); break;
        case 'n': out.push_back('\n'); break;
        case 'r': out.push_back('\r'); break;
        case 't': out.push_back('\t'); break;
        case 'u': {
          uint32_t cp = hex4();
          if (cp >= 0xD800 && cp <= 0xDBFF) {  // surrogate pair
            if (end - p < 2 || p[0] != '\\' || p[1] != 'u') fail("lone surrogate");
            p += 2;
            const uint32_t lo = hex4();
            if (lo < 0xDC00 || lo > 0xDFFF) fail("lone surrogate");
            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          } else if (cp >= 0xDC00 && cp <= 0xDFFF) {
            fail("lone surrogate");
          }
          encode_utf8(cp, out);
          break;
        }
        default: fail("bad escape");
      }
    }
    return out;
  }
  // the JSON number grammar; returns the token
  std::string number_token() {
    ws();
    const char* s = p;
    if (p < end && *p == '-') ++p;
    if (p >= end || *p < '0' || *p > '9') fail("expected a number");
    if (*p == '0') ++p; else while (p < end && *p >= '0' && *p <= '9') ++p;
    if (p < end && *p == '.') {
      ++p;
      if (p >= end || *p < '0' || *p > '9') fail("digit expected after '.'");
      while (p < end && *p >= '0' && *p <= '9') ++p;
    }
    if (p < end && (*p == 'e' || *p == 'E')) {
      ++p;
      if (p < end && (*p == '+' || *p == '-')) ++p;
      if (p >= end || *p < '0' || *p > '9') fail("digit expected in exponent");
      while (p < end && *p >= '0' && *p <= '9') ++p;
    }
    return std::string(s, p);
  }
  float f32() {  // serde_json: f64, then `as f32`
    const std::string t = number_token();
    const double d = strtod(t.c_str(), nullptr);
    const float f = (float)d;
    if (std::isinf(f) && !std::isinf(d)) return f;  // `as f32` saturates to infinity like the cast
    return f;
  }
  size_t usize() {
    const std::string t = number_token();
    if (t.empty() || t[0] == '-' || t.find_first_of(".eE") != std::string::npos) fail("expected an unsigned integer");
    errno = 0;
    const unsigned long long v = strtoull(t.c_str(), nullptr, 10);
    if (errno) fail("integer out of range");
    return (size_t)v;
  }
  std::optional<std::string> opt_string() {
    if (literal("null")) return std::nullopt;
    return string();
  }
  void skip_value() {  // a value of a key the struct does not have
    ws();
    if (p >= end) fail("unexpected end of input");
    if (*p == '"') { string(); return; }
    if (*p == '{') {
      ++p;
      if (peek('}')) { ++p; return; }
      while (true) { string(); expect(':'); skip_value(); if (peek(',')) { ++p; continue; } expect('}'); return; }
    }
    if (*p == '[') {
      ++p;
      if (peek(']')) { ++p; return; }
      while (true) { skip_value(); if (peek(',')) { ++p; continue; } expect(']'); return; }
    }
    if (literal("true") || literal("false") || literal("null")) return;
    number_token();
  }
  // calls body(key) for every member of an object
  template <typename F>
  void object(F&& body) {
    expect('{');
    if (peek('}')) { ++p; return; }
    while (true) {
      const std::string key = string();
      expect(':');
      body(key);
      if (peek(',')) { ++p; continue; }
      expect('}');
      return;
    }
  }
  template <typename F>
  void array(F&& body) {
    expect('[');
    if (peek(']')) { ++p; return; }
    while (true) {
      body();
      if (peek(',')) { ++p; continue; }
      expect(']');
      return;
    }
  }
};
}  // namespace

PersistedIndex::PersistedIndex() = default;
PersistedIndex::~PersistedIndex() = default;
PersistedIndex::PersistedIndex(PersistedIndex&&) noexcept = default;
PersistedIndex& PersistedIndex::operator=(PersistedIndex&&) noexcept = default;

PersistedIndex PersistedIndex::from_json(const char* text, size_t n) {
  Json j{text, text + n};
  PersistedIndex out;
  bool has_chunks = false, has_emb = false, has_dim = false, has_type = false, has_model = false;
  auto once = [&](bool& seen, const char* name) {
    if (seen) j.fail((std::string("duplicate field `") + name + "`").c_str());
    seen = true;
  };
  j.object([&](const std::string& key) {
    if (key == "chunks") {
      once(has_chunks, "chunks");
      j.array([&] {
        PersistedChunk c;
        bool has_content = false;
        j.object([&](const std::string& k2) {
          if (k2 == "content") { c.content = j.string(); has_content = true; }
          else if (k2 == "title") c.title = j.opt_string();
          else if (k2 == "source") c.source = j.opt_string();
          else j.skip_value();
        });
        if (!has_content) j.fail("missing field `content`");
        out.chunks.push_back(std::move(c));
      });
    } else if (key == "embeddings") {
      once(has_emb, "embeddings");
      j.array([&] {
        std::vector<float> row;
        if (!out.embeddings.empty()) row.reserve(out.embeddings.back().size());
        j.array([&] { row.push_back(j.f32()); });
        out.embeddings.push_back(std::move(row));
      });
    } else if (key == "dimension") {
      once(has_dim, "dimension");
      out.dimension = j.usize();
    } else if (key == "embedder_type") {
      once(has_type, "embedder_type");
      out.embedder_type = j.string();
    } else if (key == "model_name") {
      once(has_model, "model_name");
      out.model_name = j.opt_string();
    } else {
      j.skip_value();
    }
  });
  j.ws();
  if (j.p != j.end) j.fail("trailing characters");
  if (!has_chunks) j.fail("missing field `chunks`");
  if (!has_emb) j.fail("missing field `embeddings`");
  if (!has_dim) j.fail("missing field `dimension`");
  return out;
}

namespace {
void json_string(const std::string& v, std::string& o) {
  o.push_back('"');
  for (const char ch : v) {
    const unsigned char c = (unsigned char)ch;
    switch (c) {
      case '"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\b': o += "\\b"; break;
      case '\f': o += "\\f"; break;
      case '\n': o += "\\n"; break;
      case '\r': o += "\\r"; break;
      case '\t': o += "\\t"; break;
      default:
        if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
        else o.push_back(ch);
    }
  }
  o.push_back('"');
}
void json_f32(float v, std::string& o) {
  if (!std::isfinite(v)) { o += "null"; return; }
  char b[40];
  for (int prec = 1; prec <= 9; ++prec) {  // fewest significant digits that round-trip (9 always do for f32)
    snprintf(b, sizeof b, "%.*g", prec, (double)v);
    if ((float)strtod(b, nullptr) == v && (v != 0.0f || std::signbit((float)strtod(b, nullptr)) == std::signbit(v))) break;
  }
  std::string t(b);
  if (t.find_first_of(".eEn") == std::string::npos) t += ".0";           // keep it a float token, like serde_json
  const size_t e = t.find("e+");
  if (e != std::string::npos) t.erase(e + 1, 1);                          // 1e+21 -> 1e21
  for (size_t k = t.find('e'); k != std::string::npos;) {                 // e-07 -> e-7
    size_t d = k + 1 + (t[k + 1] == '-' ? 1 : 0);
    while (d + 1 < t.size() && t[d] == '0') t.erase(d, 1);
    break;
  }
  o += t;
}
void json_opt(const std::optional<std::string>& v, std::string& o) {
  if (v) json_string(*v, o); else o += "null";
}
}  // namespace

std::string PersistedIndex::to_json() const {
  std::string o;
  o.reserve(64 + embeddings.size() * (dimension * 14 + 32));
  o += "{\n  \"chunks\": [";
  for (size_t i = 0; i < chunks.size(); ++i) {
    o += i ? ",\n    {\n" : "\n    {\n";
    o += "      \"content\": "; json_string(chunks[i].content, o);
    o += ",\n      \"title\": "; json_opt(chunks[i].title, o);
    o += ",\n      \"source\": "; json_opt(chunks[i].source, o);
    o += "\n    }";
  }
  o += chunks.empty() ? "],\n" : "\n  ],\n";
  o += "  \"embeddings\": [";
  for (size_t i = 0; i < embeddings.size(); ++i) {
    o += i ? ",\n    [" : "\n    [";
    for (size_t j = 0; j < embeddings[i].size(); ++j) {
      o += j ? ",\n      " : "\n      ";
      json_f32(embeddings[i][j], o);
    }
    o += embeddings[i].empty() ? "]" : "\n    ]";
  }
  o += embeddings.empty() ? "],\n" : "\n  ],\n";
  o += "  \"dimension\": " + std::to_string(dimension) + ",\n";
  o += "  \"embedder_type\": "; json_string(embedder_type, o);
  o += ",\n  \"model_name\": "; json_opt(model_name, o);
  o += "\n}";
  return o;
}

std::vector<std::pair<size_t, float>> PersistedIndex::query(const std::vector<float>& q, size_t top_k) const {
  std::vector<std::pair<size_t, float>> out;
  const size_t n = embeddings.size();
  if (top_k == 0 || n == 0) return out;
  const size_t len = q.size();
  auto it = by_len_.find(len);
  if (it == by_len_.end()) {
    DeviceRows d;
    std::vector<float> slab;
    for (size_t i = 0; i < n; ++i) {
      if (embeddings[i].size() != len) { d.n_other++; continue; }
      d.index_of.push_back((uint32_t)i);
      slab.insert(slab.end(), embeddings[i].begin(), embeddings[i].end());
    }
    if (!d.index_of.empty() && len > 0) {
      d.dev = std::make_shared<detail::DeviceDense>();
      check(trr_dense_create(default_context(), (uint32_t)len, TRR_METRIC_COSINE, TRR_DTYPE_F32, d.index_of.size(), &d.dev->h));
      check(trr_dense_append(d.dev->h, slab.data(), d.index_of.size()));
    }
    it = by_len_.emplace(len, std::move(d)).first;
  }
  const DeviceRows& d = it->second;
  const size_t k = std::min(top_k, n);
  std::vector<uint32_t> ord(k);
  std::vector<float> sc(k);
  uint32_t got = 0;
  if (d.dev) {
    const uint32_t kk = (uint32_t)std::min(k, d.index_of.size());
    check(trr_dense_search(d.dev->h, q.data(), 1, kk, ord.data(), sc.data(), &got));
  }
  // merge the device list (score desc, index asc) with the rows that score 0.0 by construction (other lengths, or every
  // row when the query is empty), keeping the stable order of the reference's sort (:494-495)
  std::vector<size_t> zeros;
  if (d.n_other || !d.dev) {
    for (size_t i = 0; i < n && zeros.size() < k; ++i)
      if (!d.dev || embeddings[i].size() != len) zeros.push_back(i);
  }
  size_t a = 0, b = 0;
  while (out.size() < k && (a < got || b < zeros.size())) {
    bool take_dev;
    if (a >= got) take_dev = false;
    else if (b >= zeros.size()) take_dev = true;
    else if (sc[a] > 0.0f) take_dev = true;
    else if (sc[a] < 0.0f) take_dev = false;
    else take_dev = d.index_of[ord[a]] < zeros[b];  // both 0.0 (or a NaN score, which partial_cmp treats as Equal)
    if (take_dev) { out.emplace_back(d.index_of[ord[a]], sc[a]); ++a; }
    else { out.emplace_back(zeros[b], 0.0f); ++b; }
  }
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
  // the fused device path identifies documents by ordinal in both indexes; a re-indexed ChunkId would appear under two
  // ordinals where the reference merges by ChunkId, so from then on the general path (fusion over ChunkIds) is used
  if (dense_.get(chunk.id) != nullptr) aligned_ = false;
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
