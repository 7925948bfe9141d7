// zstd_codec.cpp — Zstandard frames for `Compression::Zstd` (reference src/compressed.rs:41-45, 60-64: zstd::encode_all /
// zstd::decode_all of the `zstd` crate 0.13, i.e. standard frames as specified in RFC 8878).
//
// The reference links libzstd; none is available to this build, so this file restates the FORMAT:
//   * zstd_decompress: a complete frame decoder (raw / RLE / compressed blocks, Huffman literals with direct or
//     FSE-compressed weights, treeless literals, FSE sequences in predefined / RLE / described / repeat mode, repeat
//     offsets, multiple and skippable frames, content checksum verified with XXH64).  Dictionaries are not supported.
//   * zstd_compress: a frame writer with real (if modest) compression: LZ77 matches whose sequences are FSE-coded with
//     the predefined distributions, raw literals, RLE / stored blocks where coding does not pay.  Every zstd decoder (the
//     reference's included) reads what it writes; libzstd's own files are smaller (no Huffman literals, no repeat offsets).
//   * zstd_store: raw and RLE blocks only (kept as the trivially correct writer the tests compare against).
// Host-side persistence only; nothing here is on the retrieval path.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../../include/trueno_rag.hpp"

namespace trueno_rag {
namespace {

[[noreturn]] void zfail(const char* what) {
  throw Error(Error::Kind::Serialization, std::string("ZSTD decompression failed: ") + what);
}

inline int highest_set_bit(uint64_t v) {  // v != 0
  return 63 - __builtin_clzll(v);
}

// ---- bit streams ----
// forward, least-significant bit first (FSE table descriptions)
struct FwdBits {
  const uint8_t* p;
  size_t n;
  size_t bit = 0;
  uint32_t read(int nbits) {
    uint32_t v = 0;
    for (int i = 0; i < nbits; ++i, ++bit) {
      if ((bit >> 3) >= n) zfail("truncated table description");
      v |= (uint32_t)((p[bit >> 3] >> (bit & 7)) & 1u) << i;
    }
    return v;
  }
  void rewind(int nbits) { bit -= (size_t)nbits; }
  size_t bytes_used() const { return (bit + 7) >> 3; }
};

// backward: the stream is read from its end; the last byte holds a 1 marker above the payload bits.  `off` is the bit
// offset of the next unread bit; reading past the start yields zeros and leaves `off` negative (the caller checks).
struct RevBits {
  const uint8_t* p;
  int64_t off;
  RevBits(const uint8_t* src, size_t n) : p(src) {
    if (n == 0 || src[n - 1] == 0) zfail("bad bitstream end marker");
    off = (int64_t)(n - 1) * 8 + highest_set_bit(src[n - 1]);
  }
  uint64_t read(int nbits) {
    if (nbits == 0) return 0;
    off -= nbits;
    int64_t start = off;
    int take = nbits;
    if (start < 0) { take += (int)std::max<int64_t>(start, -(int64_t)nbits); start = 0; }
    uint64_t v = 0;
    for (int i = 0; i < take; ++i) {
      const int64_t b = start + i;
      v |= (uint64_t)((p[b >> 3] >> (b & 7)) & 1u) << i;
    }
    return v << (nbits - take);
  }
};

// ---- FSE ----
struct FseTable {
  int accuracy_log = 0;
  std::vector<uint8_t> symbol, num_bits;
  std::vector<uint16_t> new_state_base;
  bool valid = false;
};

void fse_build(FseTable& t, const int16_t* norm, int n_symbols, int accuracy_log) {
  const uint32_t size = 1u << accuracy_log;
  t.accuracy_log = accuracy_log;
  t.symbol.assign(size, 0);
  t.num_bits.assign(size, 0);
  t.new_state_base.assign(size, 0);
  std::vector<uint16_t> next(n_symbols, 0);
  uint32_t high = size;
  for (int s = 0; s < n_symbols; ++s)
    if (norm[s] == -1) { t.symbol[--high] = (uint8_t)s; next[s] = 1; }
  const uint32_t step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
  uint32_t pos = 0;
  for (int s = 0; s < n_symbols; ++s) {
    if (norm[s] <= 0) continue;
    next[s] = (uint16_t)norm[s];
    for (int i = 0; i < norm[s]; ++i) {
      t.symbol[pos] = (uint8_t)s;
      do { pos = (pos + step) & mask; } while (pos >= high);
    }
  }
  if (pos != 0) zfail("corrupt FSE distribution");
  for (uint32_t i = 0; i < size; ++i) {
    const uint16_t x = next[t.symbol[i]]++;
    t.num_bits[i] = (uint8_t)(accuracy_log - highest_set_bit(x));
    t.new_state_base[i] = (uint16_t)(((uint32_t)x << t.num_bits[i]) - size);
  }
  t.valid = true;
}

// reads an FSE table description; returns the bytes it occupies
size_t fse_read_description(FseTable& t, const uint8_t* src, size_t n, int max_accuracy_log, int max_symbols) {
  FwdBits b{src, n};
  const int accuracy_log = 5 + (int)b.read(4);
  if (accuracy_log > max_accuracy_log) zfail("FSE accuracy log too large");
  int32_t remaining = 1 << accuracy_log;
  int16_t norm[256];
  int symb = 0;
  while (remaining > 0 && symb < max_symbols) {
    const int bits = highest_set_bit((uint64_t)remaining + 1) + 1;
    uint32_t val = b.read(bits);
    const uint32_t lower_mask = (1u << (bits - 1)) - 1;
    const uint32_t threshold = (1u << bits) - 1 - (uint32_t)(remaining + 1);
    if ((val & lower_mask) < threshold) {
      b.rewind(1);
      val &= lower_mask;
    } else if (val > lower_mask) {
      val -= threshold;
    }
    const int16_t proba = (int16_t)val - 1;
    remaining -= proba < 0 ? -proba : proba;
    norm[symb++] = proba;
    if (proba == 0) {
      uint32_t repeat = b.read(2);
      while (true) {
        for (uint32_t i = 0; i < repeat && symb < max_symbols; ++i) norm[symb++] = 0;
        if (repeat == 3) repeat = b.read(2); else break;
      }
    }
  }
  if (remaining != 0 || symb > max_symbols) zfail("corrupt FSE table description");
  fse_build(t, norm, symb, accuracy_log);
  return b.bytes_used();
}

void fse_build_rle(FseTable& t, uint8_t sym) {
  t.accuracy_log = 0;
  t.symbol.assign(1, sym);
  t.num_bits.assign(1, 0);
  t.new_state_base.assign(1, 0);
  t.valid = true;
}

// ---- Huffman ----
struct HufTable {
  int max_bits = 0;
  std::vector<uint8_t> symbol, num_bits;
  bool valid = false;
};

void huf_build(HufTable& t, const uint8_t* bits, int n_symbols) {
  int max_bits = 0;
  uint32_t rank_count[17] = {0};
  for (int i = 0; i < n_symbols; ++i) {
    if (bits[i] > 16) zfail("Huffman code too long");
    max_bits = std::max<int>(max_bits, bits[i]);
    rank_count[bits[i]]++;
  }
  if (max_bits == 0 || max_bits > 11) zfail("corrupt Huffman table");
  const uint32_t size = 1u << max_bits;
  t.max_bits = max_bits;
  t.symbol.assign(size, 0);
  t.num_bits.assign(size, 0);
  uint32_t rank_idx[18] = {0};
  rank_idx[max_bits] = 0;
  for (int i = max_bits; i >= 1; --i) {
    rank_idx[i - 1] = rank_idx[i] + rank_count[i] * (1u << (max_bits - i));
    if (rank_idx[i - 1] > size) zfail("corrupt Huffman table");
    for (uint32_t k = rank_idx[i]; k < rank_idx[i - 1]; ++k) t.num_bits[k] = (uint8_t)i;
  }
  if (rank_idx[0] != size) zfail("corrupt Huffman table");
  for (int i = 0; i < n_symbols; ++i) {
    if (bits[i] == 0) continue;
    const uint32_t code = rank_idx[bits[i]], len = 1u << (max_bits - bits[i]);
    for (uint32_t k = 0; k < len; ++k) t.symbol[code + k] = (uint8_t)i;
    rank_idx[bits[i]] += len;
  }
  t.valid = true;
}

void huf_build_from_weights(HufTable& t, uint8_t* weights, int n_weights) {
  // the last weight is implied: the sum of 2^(w-1) must be a power of two
  uint64_t sum = 0;
  for (int i = 0; i < n_weights; ++i) {
    if (weights[i] > 11) zfail("corrupt Huffman weights");
    sum += weights[i] ? (1ull << (weights[i] - 1)) : 0;
  }
  if (sum == 0) zfail("corrupt Huffman weights");
  const int max_bits = highest_set_bit(sum) + 1;
  const uint64_t left = (1ull << max_bits) - sum;
  if (left & (left - 1)) zfail("corrupt Huffman weights");
  if (n_weights >= 256) zfail("too many Huffman weights");
  weights[n_weights] = (uint8_t)(highest_set_bit(left) + 1);
  const int n = n_weights + 1;
  uint8_t bits[256];
  for (int i = 0; i < n; ++i) bits[i] = weights[i] ? (uint8_t)(max_bits + 1 - weights[i]) : 0;
  huf_build(t, bits, n);
}

// reads a Huffman tree description; returns the bytes it occupies
size_t huf_read_description(HufTable& t, const uint8_t* src, size_t n) {
  if (n < 1) zfail("truncated Huffman description");
  const uint8_t header = src[0];
  uint8_t weights[257];
  int n_weights = 0;
  size_t used;
  if (header >= 128) {  // direct: 4 bits per weight
    n_weights = header - 127;
    const size_t bytes = (size_t)(n_weights + 1) / 2;
    if (1 + bytes > n) zfail("truncated Huffman weights");
    for (int i = 0; i < n_weights; ++i) {
      const uint8_t b = src[1 + i / 2];
      weights[i] = (i & 1) ? (b & 15) : (b >> 4);
    }
    used = 1 + bytes;
  } else {  // FSE-compressed weights, two interleaved states over a backward bitstream
    const size_t csize = header;
    if (csize == 0 || 1 + csize > n) zfail("truncated Huffman weights");
    FseTable ft;
    const size_t hdr = fse_read_description(ft, src + 1, csize, 7, 256);
    if (hdr >= csize) zfail("corrupt Huffman weights");
    RevBits rb(src + 1 + hdr, csize - hdr);
    uint32_t s1 = (uint32_t)rb.read(ft.accuracy_log), s2 = (uint32_t)rb.read(ft.accuracy_log);
    if (rb.off < 0) zfail("corrupt Huffman weights");
    while (true) {
      if (n_weights >= 255) zfail("too many Huffman weights");
      weights[n_weights++] = ft.symbol[s1];
      s1 = ft.new_state_base[s1] + (uint32_t)rb.read(ft.num_bits[s1]);
      if (rb.off < 0) { weights[n_weights++] = ft.symbol[s2]; break; }
      if (n_weights >= 255) zfail("too many Huffman weights");
      weights[n_weights++] = ft.symbol[s2];
      s2 = ft.new_state_base[s2] + (uint32_t)rb.read(ft.num_bits[s2]);
      if (rb.off < 0) { weights[n_weights++] = ft.symbol[s1]; break; }
    }
    used = 1 + csize;
  }
  huf_build_from_weights(t, weights, n_weights);
  return used;
}

void huf_decode_stream(const HufTable& t, const uint8_t* src, size_t n, uint8_t* out, size_t out_n) {
  RevBits rb(src, n);
  const uint32_t mask = (1u << t.max_bits) - 1;
  uint32_t state = (uint32_t)rb.read(t.max_bits);
  for (size_t i = 0; i < out_n; ++i) {
    out[i] = t.symbol[state];
    const int nb = t.num_bits[state];
    state = ((state << nb) & mask) | (uint32_t)rb.read(nb);
    if (rb.off < -(int64_t)t.max_bits) zfail("Huffman stream overrun");
  }
  if (rb.off != -(int64_t)t.max_bits) zfail("Huffman stream not fully consumed");
}

// ---- sequences ----
const int16_t LL_DEFAULT[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
const int16_t ML_DEFAULT[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
const int16_t OF_DEFAULT[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
const uint32_t LL_BASE[36] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40,
                              48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
const uint8_t LL_EXTRA[36] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
const uint32_t ML_BASE[53] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
                              30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099,
                              8195, 16387, 32771, 65539};
const uint8_t ML_EXTRA[53] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                              0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};

struct FrameState {
  HufTable huf;
  FseTable ll, of, ml;
  uint64_t rep[3] = {1, 4, 8};
};

size_t read_seq_table(FseTable& t, int mode, const uint8_t* src, size_t n, const int16_t* def, int def_n, int def_log,
                      int max_log, int max_symbols) {
  switch (mode) {
    case 0: fse_build(t, def, def_n, def_log); return 0;
    case 1:
      if (n < 1) zfail("truncated RLE table");
      if (src[0] >= max_symbols) zfail("RLE symbol out of range");
      fse_build_rle(t, src[0]);
      return 1;
    case 2: return fse_read_description(t, src, n, max_log, max_symbols);
    default:
      if (!t.valid) zfail("repeat mode without a previous table");
      return 0;
  }
}

void decode_block(FrameState& fs, const uint8_t* src, size_t n, std::vector<uint8_t>& out, uint64_t max_out) {
  // a block regenerates at most Block_Maximum_Size = 128 KB, whatever the frame declares
  max_out = std::min<uint64_t>(max_out, (uint64_t)out.size() + (128u << 10));
  // ---- literals section ----
  if (n < 1) zfail("truncated block");
  const int lit_type = src[0] & 3, size_format = (src[0] >> 2) & 3;
  std::vector<uint8_t> literals;
  size_t pos;
  if (lit_type <= 1) {  // raw / RLE
    size_t regen;
    if (size_format == 0 || size_format == 2) { regen = src[0] >> 3; pos = 1; }
    else if (size_format == 1) { if (n < 2) zfail("truncated literals header"); regen = (src[0] >> 4) | ((size_t)src[1] << 4); pos = 2; }
    else { if (n < 3) zfail("truncated literals header"); regen = (src[0] >> 4) | ((size_t)src[1] << 4) | ((size_t)src[2] << 12); pos = 3; }
    if (regen > (128u << 10)) zfail("literals too large");
    if (lit_type == 0) {
      if (pos + regen > n) zfail("truncated raw literals");
      literals.assign(src + pos, src + pos + regen);
      pos += regen;
    } else {
      if (pos + 1 > n) zfail("truncated RLE literals");
      literals.assign(regen, src[pos]);
      pos += 1;
    }
  } else {  // Huffman-compressed / treeless
    size_t regen, csize;
    int streams;
    if (size_format <= 1) {
      if (n < 3) zfail("truncated literals header");
      const uint32_t h = src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16);
      regen = (h >> 4) & 0x3FF; csize = (h >> 14) & 0x3FF; streams = size_format == 0 ? 1 : 4; pos = 3;
    } else if (size_format == 2) {
      if (n < 4) zfail("truncated literals header");
      const uint32_t h = src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24);
      regen = (h >> 4) & 0x3FFF; csize = (h >> 18) & 0x3FFF; streams = 4; pos = 4;
    } else {
      if (n < 5) zfail("truncated literals header");
      const uint64_t h = src[0] | ((uint64_t)src[1] << 8) | ((uint64_t)src[2] << 16) | ((uint64_t)src[3] << 24) | ((uint64_t)src[4] << 32);
      regen = (h >> 4) & 0x3FFFF; csize = (h >> 22) & 0x3FFFF; streams = 4; pos = 5;
    }
    if (regen > (128u << 10)) zfail("literals too large");
    if (pos + csize > n) zfail("truncated compressed literals");
    const uint8_t* ls = src + pos;
    size_t ln = csize;
    if (lit_type == 2) {
      const size_t used = huf_read_description(fs.huf, ls, ln);
      ls += used;
      ln -= used;
    } else if (!fs.huf.valid) {
      zfail("treeless literals without a previous Huffman table");
    }
    literals.resize(regen);
    if (streams == 1) {
      huf_decode_stream(fs.huf, ls, ln, literals.data(), regen);
    } else {
      if (ln < 6) zfail("truncated jump table");
      const size_t s1 = ls[0] | ((size_t)ls[1] << 8), s2 = ls[2] | ((size_t)ls[3] << 8), s3 = ls[4] | ((size_t)ls[5] << 8);
      if (6 + s1 + s2 + s3 > ln) zfail("corrupt jump table");
      const size_t s4 = ln - 6 - s1 - s2 - s3;
      const size_t part = (regen + 3) / 4;
      if (part * 3 > regen) zfail("corrupt 4-stream literals");
      const uint8_t* q = ls + 6;
      huf_decode_stream(fs.huf, q, s1, literals.data(), part);
      huf_decode_stream(fs.huf, q + s1, s2, literals.data() + part, part);
      huf_decode_stream(fs.huf, q + s1 + s2, s3, literals.data() + 2 * part, part);
      huf_decode_stream(fs.huf, q + s1 + s2 + s3, s4, literals.data() + 3 * part, regen - 3 * part);
    }
    pos += csize;
  }

  // ---- sequences section ----
  if (pos >= n) zfail("missing sequences section");
  size_t n_seq;
  {
    const uint8_t b0 = src[pos++];
    if (b0 == 0) n_seq = 0;
    else if (b0 < 128) n_seq = b0;
    else if (b0 < 255) { if (pos >= n) zfail("truncated sequence count"); n_seq = ((size_t)(b0 - 128) << 8) + src[pos++]; }
    else { if (pos + 2 > n) zfail("truncated sequence count"); n_seq = (size_t)src[pos] + ((size_t)src[pos + 1] << 8) + 0x7F00; pos += 2; }
  }
  if (n_seq == 0) {
    if (pos != n) zfail("trailing bytes after an empty sequences section");
    if (out.size() + literals.size() > max_out) zfail("output larger than declared");
    out.insert(out.end(), literals.begin(), literals.end());
    return;
  }
  if (pos >= n) zfail("truncated sequences header");
  const uint8_t modes = src[pos++];
  if (modes & 3) zfail("reserved bits set in the sequences header");
  pos += read_seq_table(fs.ll, (modes >> 6) & 3, src + pos, n - pos, LL_DEFAULT, 36, 6, 9, 36);
  pos += read_seq_table(fs.of, (modes >> 4) & 3, src + pos, n - pos, OF_DEFAULT, 29, 5, 8, 32);
  pos += read_seq_table(fs.ml, (modes >> 2) & 3, src + pos, n - pos, ML_DEFAULT, 53, 6, 9, 53);
  if (pos >= n) zfail("missing sequence bitstream");
  RevBits rb(src + pos, n - pos);
  uint32_t ll_s = (uint32_t)rb.read(fs.ll.accuracy_log);
  uint32_t of_s = (uint32_t)rb.read(fs.of.accuracy_log);
  uint32_t ml_s = (uint32_t)rb.read(fs.ml.accuracy_log);
  if (rb.off < 0) zfail("truncated sequence bitstream");
  size_t lit_pos = 0;
  for (size_t i = 0; i < n_seq; ++i) {
    const uint8_t of_code = fs.of.symbol[of_s], ml_code = fs.ml.symbol[ml_s], ll_code = fs.ll.symbol[ll_s];
    if (of_code > 31 || ml_code > 52 || ll_code > 35) zfail("sequence code out of range");
    const uint64_t of_value = (1ull << of_code) + rb.read(of_code);
    const uint64_t match_len = ML_BASE[ml_code] + rb.read(ML_EXTRA[ml_code]);
    const uint64_t lit_len = LL_BASE[ll_code] + rb.read(LL_EXTRA[ll_code]);
    if (rb.off < 0) zfail("sequence bitstream overrun");
    if (i + 1 < n_seq) {  // state updates: literal length, match length, offset
      ll_s = fs.ll.new_state_base[ll_s] + (uint32_t)rb.read(fs.ll.num_bits[ll_s]);
      ml_s = fs.ml.new_state_base[ml_s] + (uint32_t)rb.read(fs.ml.num_bits[ml_s]);
      of_s = fs.of.new_state_base[of_s] + (uint32_t)rb.read(fs.of.num_bits[of_s]);
      if (rb.off < 0) zfail("sequence bitstream overrun");
    }
    uint64_t offset;
    if (of_value > 3) {
      offset = of_value - 3;
      fs.rep[2] = fs.rep[1]; fs.rep[1] = fs.rep[0]; fs.rep[0] = offset;
    } else {
      uint32_t idx = (uint32_t)of_value - 1;
      if (lit_len == 0) idx++;
      if (idx == 0) {
        offset = fs.rep[0];
      } else {
        offset = idx < 3 ? fs.rep[idx] : fs.rep[0] - 1;
        if (idx > 1) fs.rep[2] = fs.rep[1];
        fs.rep[1] = fs.rep[0];
        fs.rep[0] = offset;
      }
    }
    if (lit_len > literals.size() - lit_pos) zfail("sequence needs more literals than the block has");
    if (out.size() + lit_len + match_len > max_out) zfail("output larger than declared");
    out.insert(out.end(), literals.begin() + (long)lit_pos, literals.begin() + (long)(lit_pos + lit_len));
    lit_pos += lit_len;
    if (offset == 0 || offset > out.size()) zfail("match offset out of range");
    const size_t from = out.size() - (size_t)offset;
    for (uint64_t k = 0; k < match_len; ++k) out.push_back(out[from + k]);  // may overlap what it writes
  }
  if (rb.off != 0) zfail("sequence bitstream not fully consumed");
  if (out.size() + (literals.size() - lit_pos) > max_out) zfail("output larger than declared");
  out.insert(out.end(), literals.begin() + (long)lit_pos, literals.end());
}

// ---- XXH64 (content checksum = low 32 bits) ----
inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
uint64_t xxh64(const uint8_t* p, size_t len, uint64_t seed) {
  const uint64_t P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P3 = 1609587929392839161ull,
                 P4 = 9650029242287828579ull, P5 = 2870177450012600261ull;
  auto round = [&](uint64_t acc, uint64_t in) { acc += in * P2; acc = rotl64(acc, 31); return acc * P1; };
  auto merge = [&](uint64_t acc, uint64_t v) { acc ^= round(0, v); return acc * P1 + P4; };
  const uint8_t* end = p + len;
  uint64_t h;
  if (len >= 32) {
    uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    const uint8_t* limit = end - 32;
    do { v1 = round(v1, rd64(p)); v2 = round(v2, rd64(p + 8)); v3 = round(v3, rd64(p + 16)); v4 = round(v4, rd64(p + 24)); p += 32; } while (p <= limit);
    h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
    h = merge(h, v1); h = merge(h, v2); h = merge(h, v3); h = merge(h, v4);
  } else {
    h = seed + P5;
  }
  h += (uint64_t)len;
  while (p + 8 <= end) { h ^= round(0, rd64(p)); h = rotl64(h, 27) * P1 + P4; p += 8; }
  if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * P1; h = rotl64(h, 23) * P2 + P3; p += 4; }
  while (p < end) { h ^= (*p) * P5; h = rotl64(h, 11) * P1; ++p; }
  h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
  return h;
}

// decodes one frame starting at src (after the magic number was recognised); returns the bytes consumed
size_t decode_frame(const uint8_t* src, size_t n, std::vector<uint8_t>& out) {
  size_t pos = 4;
  if (pos >= n) zfail("truncated frame header");
  const uint8_t fhd = src[pos++];
  const int fcs_flag = fhd >> 6, single_segment = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, dict_flag = fhd & 3;
  if (fhd & 0x08) zfail("reserved bit set in the frame header");
  if (!single_segment) {
    if (pos >= n) zfail("truncated frame header");
    pos++;  // window descriptor: the whole output is kept, no window to size
  }
  if (dict_flag) {
    const int dsz = dict_flag == 3 ? 4 : dict_flag;
    if (pos + dsz > n) zfail("truncated frame header");
    uint32_t id = 0;
    for (int i = 0; i < dsz; ++i) id |= (uint32_t)src[pos + i] << (8 * i);
    if (id != 0) zfail("frames that need a dictionary are not supported");
    pos += dsz;
  }
  uint64_t content_size = UINT64_MAX;
  {
    const int fsz = fcs_flag == 0 ? (single_segment ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
    if (pos + fsz > n) zfail("truncated frame header");
    if (fsz) {
      content_size = 0;
      for (int i = 0; i < fsz; ++i) content_size |= (uint64_t)src[pos + i] << (8 * i);
      if (fsz == 2) content_size += 256;
      pos += fsz;
    }
  }
  const size_t out_start = out.size();
  const uint64_t max_out = content_size == UINT64_MAX ? UINT64_MAX : out_start + content_size;
  if (content_size != UINT64_MAX && content_size < ((uint64_t)1 << 32)) out.reserve(out_start + (size_t)content_size);
  FrameState fs;
  while (true) {
    if (pos + 3 > n) zfail("truncated block header");
    const uint32_t bh = src[pos] | ((uint32_t)src[pos + 1] << 8) | ((uint32_t)src[pos + 2] << 16);
    pos += 3;
    const int last = bh & 1, type = (bh >> 1) & 3;
    const size_t bsize = bh >> 3;
    if (type == 3) zfail("reserved block type");
    if (bsize > (128u << 10)) zfail("block larger than 128 KB");
    if (type == 0) {
      if (pos + bsize > n) zfail("truncated raw block");
      if (out.size() + bsize > max_out) zfail("output larger than declared");
      out.insert(out.end(), src + pos, src + pos + bsize);
      pos += bsize;
    } else if (type == 1) {
      if (pos + 1 > n) zfail("truncated RLE block");
      if (out.size() + bsize > max_out) zfail("output larger than declared");
      out.insert(out.end(), bsize, src[pos]);
      pos += 1;
    } else {
      if (pos + bsize > n) zfail("truncated compressed block");
      decode_block(fs, src + pos, bsize, out, max_out);
      pos += bsize;
    }
    if (last) break;
  }
  if (content_size != UINT64_MAX && out.size() - out_start != content_size) zfail("content size mismatch");
  if (checksum) {
    if (pos + 4 > n) zfail("truncated content checksum");
    const uint32_t want = rd32(src + pos);
    pos += 4;
    if ((uint32_t)xxh64(out.data() + out_start, out.size() - out_start, 0) != want) zfail("content checksum mismatch");
  }
  return pos;
}

}  // namespace

// zstd::decode_all: every frame of the input, skippable frames skipped
std::vector<uint8_t> zstd_decompress(const uint8_t* src, size_t n) {
  std::vector<uint8_t> out;
  size_t pos = 0;
  bool any = false;
  while (pos < n) {
    if (n - pos < 4) zfail("truncated magic number");
    const uint32_t magic = rd32(src + pos);
    if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
      if (n - pos < 8) zfail("truncated skippable frame");
      const uint64_t sz = rd32(src + pos + 4);
      if (sz > n - pos - 8) zfail("truncated skippable frame");
      pos += 8 + (size_t)sz;
      continue;
    }
    if (magic != 0xFD2FB528u) zfail("unknown frame descriptor");
    pos += decode_frame(src + pos, n - pos, out);
    any = true;
  }
  if (!any) zfail("no frame");
  return out;
}

// ------------------------------------------------------------------------------------------------
// compressor: LZ77 matches + FSE-coded sequences with the PREDEFINED distributions, raw literals
// ------------------------------------------------------------------------------------------------
namespace {

// forward bit writer (least-significant bit first); the decoder reads the finished stream backward from the end marker
struct BitWriter {
  std::vector<uint8_t> out;
  uint64_t acc = 0;
  int nacc = 0;
  void add(uint64_t v, int nbits) {
    while (nbits > 0) {  // keep the accumulator below 64 bits
      const int take = std::min(nbits, 32);
      acc |= (v & ((1ull << take) - 1)) << nacc;
      nacc += take;
      v >>= take;
      nbits -= take;
      while (nacc >= 8) { out.push_back((uint8_t)acc); acc >>= 8; nacc -= 8; }
    }
  }
  void close() {  // the 1 marker above the payload, then zero padding
    add(1, 1);
    if (nacc) { out.push_back((uint8_t)acc); acc = 0; nacc = 0; }
  }
};

// FSE encoding table of a normalised distribution (the mirror image of fse_build)
struct FseEnc {
  int log = 0;
  std::vector<uint16_t> state_table;                     // [size]
  std::vector<int32_t> delta_nb_bits, delta_find_state;  // [n_symbols]
  void build(const int16_t* norm, int n_symbols, int accuracy_log) {
    log = accuracy_log;
    const uint32_t size = 1u << log, mask = size - 1, step = (size >> 1) + (size >> 3) + 3;
    std::vector<uint8_t> symbol(size, 0);
    std::vector<uint32_t> cumul(n_symbols + 1, 0);
    uint32_t high = size - 1;
    for (int s = 0; s < n_symbols; ++s) {
      if (norm[s] == -1) { cumul[s + 1] = cumul[s] + 1; symbol[high--] = (uint8_t)s; }
      else cumul[s + 1] = cumul[s] + (uint32_t)norm[s];
    }
    uint32_t pos = 0;
    for (int s = 0; s < n_symbols; ++s)
      for (int i = 0; i < norm[s]; ++i) {
        symbol[pos] = (uint8_t)s;
        do { pos = (pos + step) & mask; } while (pos > high);
      }
    state_table.assign(size, 0);
    {
      std::vector<uint32_t> c(cumul);
      for (uint32_t u = 0; u < size; ++u) state_table[c[symbol[u]]++] = (uint16_t)(size + u);
    }
    delta_nb_bits.assign(n_symbols, 0);
    delta_find_state.assign(n_symbols, 0);
    int total = 0;
    for (int s = 0; s < n_symbols; ++s) {
      if (norm[s] == 0) {
        delta_nb_bits[s] = ((log + 1) << 16) - (1 << log);
      } else if (norm[s] == -1 || norm[s] == 1) {
        delta_nb_bits[s] = (log << 16) - (1 << log);
        delta_find_state[s] = total - 1;
        total += 1;
      } else {
        const int max_bits_out = log - highest_set_bit((uint64_t)(norm[s] - 1));
        const int min_state_plus = norm[s] << max_bits_out;
        delta_nb_bits[s] = (max_bits_out << 16) - min_state_plus;
        delta_find_state[s] = total - norm[s];
        total += norm[s];
      }
    }
  }
  uint32_t init(int sym) const {
    const int nb = (delta_nb_bits[sym] + (1 << 15)) >> 16;
    const int value = (nb << 16) - delta_nb_bits[sym];
    return state_table[(value >> nb) + delta_find_state[sym]];
  }
  void encode(BitWriter& w, uint32_t& state, int sym) const {
    const int nb = (int)((state + (uint32_t)delta_nb_bits[sym]) >> 16);
    w.add(state, nb);
    state = state_table[(state >> nb) + delta_find_state[sym]];
  }
  void flush(BitWriter& w, uint32_t state) const { w.add(state, log); }
};

template <size_t N>
int code_of(const uint32_t (&base)[N], uint32_t v) {  // the largest code whose baseline is <= v
  int c = 0;
  for (size_t i = 0; i < N; ++i) if (base[i] <= v) c = (int)i;
  return c;
}

struct Seq { uint32_t lit_len, match_len, offset; };

// encoding tables of the three predefined distributions (built once; C++11 static initialisation is thread-safe)
struct SeqEnc {
  FseEnc ll, of, ml;
  SeqEnc() {
    ll.build(LL_DEFAULT, 36, 6);
    of.build(OF_DEFAULT, 29, 5);
    ml.build(ML_DEFAULT, 53, 6);
  }
};
const SeqEnc& seq_enc() {
  static const SeqEnc e;
  return e;
}

// one block of at most 128 KB: returns false when coding it does not pay (the caller then stores it)
bool compress_block(const uint8_t* base, size_t bs, size_t be, std::vector<int64_t>& table, std::vector<uint8_t>& out) {
  const size_t n = be - bs;
  std::vector<Seq> seqs;
  std::vector<uint8_t> lits;
  lits.reserve(n);
  size_t anchor = bs, i = bs;
  if (n >= 16) {
    const size_t limit = be - 8;
    while (i < limit) {
      const uint32_t v = rd32(base + i);
      const uint32_t h = (v * 2654435761u) >> 15;  // 17 bits
      const int64_t cand = table[h];
      table[h] = (int64_t)i;
      if (cand >= 0 && i - (size_t)cand < ((size_t)1 << 27) && rd32(base + cand) == v) {
        size_t ml = 4;
        while (i + ml < be && base[(size_t)cand + ml] == base[i + ml]) ++ml;
        seqs.push_back(Seq{(uint32_t)(i - anchor), (uint32_t)ml, (uint32_t)(i - (size_t)cand)});
        lits.insert(lits.end(), base + anchor, base + i);
        // index a few positions inside the match so that later data finds it
        for (size_t k = i + 1; k + 4 <= be && k < i + ml; k += 3) table[(rd32(base + k) * 2654435761u) >> 15] = (int64_t)k;
        i += ml;
        anchor = i;
      } else {
        ++i;
      }
    }
  }
  if (seqs.empty()) return false;
  lits.insert(lits.end(), base + anchor, base + be);

  const FseEnc &ll_enc = seq_enc().ll, &of_enc = seq_enc().of, &ml_enc = seq_enc().ml;
  std::vector<uint8_t> blk;
  // literals section: raw
  const size_t L = lits.size();
  if (L < 32) blk.push_back((uint8_t)(L << 3));
  else if (L < 4096) { blk.push_back((uint8_t)((L << 4) | 4)); blk.push_back((uint8_t)(L >> 4)); }
  else { blk.push_back((uint8_t)((L << 4) | 12)); blk.push_back((uint8_t)(L >> 4)); blk.push_back((uint8_t)(L >> 12)); }
  blk.insert(blk.end(), lits.begin(), lits.end());
  // sequences section: count, modes (all predefined), bitstream written last sequence first
  const size_t ns = seqs.size();
  if (ns < 128) blk.push_back((uint8_t)ns);
  else if (ns < 0x7F00) { blk.push_back((uint8_t)((ns >> 8) + 128)); blk.push_back((uint8_t)ns); }
  else { blk.push_back(255); blk.push_back((uint8_t)(ns - 0x7F00)); blk.push_back((uint8_t)((ns - 0x7F00) >> 8)); }
  blk.push_back(0);
  BitWriter w;
  auto codes = [&](const Seq& q, int& llc, int& mlc, int& ofc, uint32_t& ofv) {
    llc = code_of(LL_BASE, q.lit_len);
    mlc = code_of(ML_BASE, q.match_len);
    ofv = q.offset + 3;  // never a repeat-offset code
    ofc = highest_set_bit(ofv);
  };
  int llc, mlc, ofc;
  uint32_t ofv;
  codes(seqs[ns - 1], llc, mlc, ofc, ofv);
  uint32_t ml_state = ml_enc.init(mlc), of_state = of_enc.init(ofc), ll_state = ll_enc.init(llc);
  w.add(seqs[ns - 1].lit_len - LL_BASE[llc], LL_EXTRA[llc]);
  w.add(seqs[ns - 1].match_len - ML_BASE[mlc], ML_EXTRA[mlc]);
  w.add(ofv - (1u << ofc), ofc);
  for (size_t k = ns - 1; k-- > 0;) {
    codes(seqs[k], llc, mlc, ofc, ofv);
    of_enc.encode(w, of_state, ofc);
    ml_enc.encode(w, ml_state, mlc);
    ll_enc.encode(w, ll_state, llc);
    w.add(seqs[k].lit_len - LL_BASE[llc], LL_EXTRA[llc]);
    w.add(seqs[k].match_len - ML_BASE[mlc], ML_EXTRA[mlc]);
    w.add(ofv - (1u << ofc), ofc);
  }
  ml_enc.flush(w, ml_state);
  of_enc.flush(w, of_state);
  ll_enc.flush(w, ll_state);
  w.close();
  blk.insert(blk.end(), w.out.begin(), w.out.end());
  if (blk.size() >= n) return false;
  out.swap(blk);
  return true;
}

}  // namespace

// zstd::encode_all: one frame (single segment, content size declared, no checksum); blocks are compressed (LZ77 matches,
// sequences FSE-coded with the predefined distributions, raw literals), RLE, or stored when coding does not pay
std::vector<uint8_t> zstd_compress(const uint8_t* src, size_t n) {
  std::vector<uint8_t> out;
  out.reserve(n / 2 + 64);
  const uint32_t magic = 0xFD2FB528u;
  out.insert(out.end(), reinterpret_cast<const uint8_t*>(&magic), reinterpret_cast<const uint8_t*>(&magic) + 4);
  int fcs_flag, fsz;
  uint64_t field = n;
  if (n < 256) { fcs_flag = 0; fsz = 1; }
  else if (n < 65536 + 256) { fcs_flag = 1; fsz = 2; field = n - 256; }
  else if (n < ((uint64_t)1 << 32)) { fcs_flag = 2; fsz = 4; }
  else { fcs_flag = 3; fsz = 8; }
  out.push_back((uint8_t)((fcs_flag << 6) | (1 << 5)));
  for (int i = 0; i < fsz; ++i) out.push_back((uint8_t)(field >> (8 * i)));
  auto header = [&](int type, size_t size, bool last) {
    const uint32_t h = (uint32_t)(last ? 1 : 0) | ((uint32_t)type << 1) | ((uint32_t)size << 3);
    out.push_back((uint8_t)h); out.push_back((uint8_t)(h >> 8)); out.push_back((uint8_t)(h >> 16));
  };
  if (n == 0) { header(0, 0, true); return out; }
  const size_t BLOCK = 128u << 10;
  std::vector<int64_t> table((size_t)1 << 17, -1);
  std::vector<uint8_t> blk;
  for (size_t bs = 0; bs < n; bs += BLOCK) {
    const size_t be = std::min(n, bs + BLOCK), len = be - bs;
    const bool last = be == n;
    bool same = true;
    for (size_t k = bs + 1; k < be && same; ++k) same = src[k] == src[bs];
    if (same && len > 1) {
      header(1, len, last);
      out.push_back(src[bs]);
    } else if (compress_block(src, bs, be, table, blk)) {
      header(2, blk.size(), last);
      out.insert(out.end(), blk.begin(), blk.end());
    } else {
      header(0, len, last);
      out.insert(out.end(), src + bs, src + be);
    }
  }
  return out;
}

// A valid frame made of raw and RLE blocks only (single segment, content size declared, no checksum).
std::vector<uint8_t> zstd_store(const uint8_t* src, size_t n) {
  std::vector<uint8_t> out;
  out.reserve(n + n / (128u << 10) * 3 + 32);
  const uint32_t magic = 0xFD2FB528u;
  out.insert(out.end(), reinterpret_cast<const uint8_t*>(&magic), reinterpret_cast<const uint8_t*>(&magic) + 4);
  int fcs_flag, fsz;
  uint64_t field = n;
  if (n < 256) { fcs_flag = 0; fsz = 1; }
  else if (n < 65536 + 256) { fcs_flag = 1; fsz = 2; field = n - 256; }
  else if (n < ((uint64_t)1 << 32)) { fcs_flag = 2; fsz = 4; }
  else { fcs_flag = 3; fsz = 8; }
  out.push_back((uint8_t)((fcs_flag << 6) | (1 << 5)));  // single segment
  for (int i = 0; i < fsz; ++i) out.push_back((uint8_t)(field >> (8 * i)));
  const size_t BLOCK = 128u << 10, MIN_RUN = 32;
  auto block = [&](int type, size_t size, bool last) {
    const uint32_t h = (uint32_t)(last ? 1 : 0) | ((uint32_t)type << 1) | ((uint32_t)size << 3);
    out.push_back((uint8_t)h); out.push_back((uint8_t)(h >> 8)); out.push_back((uint8_t)(h >> 16));
  };
  if (n == 0) { block(0, 0, true); return out; }
  size_t i = 0;
  while (i < n) {
    // a run of one byte value?
    size_t run = 1;
    while (i + run < n && run < BLOCK && src[i + run] == src[i]) ++run;
    if (run >= MIN_RUN) {
      block(1, run, i + run == n);
      out.push_back(src[i]);
      i += run;
      continue;
    }
    // raw bytes up to the next long run (or the block limit)
    size_t j = i + run;
    while (j < n && j - i < BLOCK) {
      size_t r = 1;
      while (j + r < n && r < MIN_RUN && src[j + r] == src[j]) ++r;
      if (r >= MIN_RUN) break;
      j += r;
    }
    const size_t len = std::min(j - i, BLOCK);
    block(0, len, i + len == n);
    out.insert(out.end(), src + i, src + i + len);
    i += len;
  }
  return out;
}

}  // namespace trueno_rag
