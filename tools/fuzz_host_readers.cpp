// tools/fuzz_host_readers.cpp - mutation fuzzing of the host-side readers (ZSTD frames, LZ4 blocks, bincode BM25Index,
// the CLI's index.json) under AddressSanitizer / UBSan.  Development helper, not part of the library or the test suite.
//
//   python tools/fuzz_seeds.py /tmp/zf          # writes seed files (needs pyarrow for the zstd frames)
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=all -Iinclude tools/fuzz_host_readers.cpp \
//       trueno_rag_b200/csrc/host/host_mirror.cpp trueno_rag_b200/csrc/host/zstd_codec.cpp \
//       -Ltrueno_rag_b200 -ltrueno_rag_b200 -Wl,-rpath,$PWD/trueno_rag_b200 -o /tmp/zf/fuzz && /tmp/zf/fuzz /tmp/zf
//
// Round 1: 20000 mutations (1-3 bit flips / byte overwrites, 1 in 8 truncated) per seed, 8 seeds: no report.
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../include/trueno_rag.hpp"

namespace trueno_rag {
std::vector<uint8_t> zstd_decompress(const uint8_t*, size_t);
std::vector<uint8_t> zstd_compress(const uint8_t*, size_t);
}
using namespace trueno_rag;

static std::vector<uint8_t> load(const std::string& p) {
  std::vector<uint8_t> d;
  FILE* f = fopen(p.c_str(), "rb");
  if (!f) { fprintf(stderr, "missing seed %s\n", p.c_str()); return d; }
  d.resize(1 << 22);
  d.resize(fread(d.data(), 1, d.size(), f));
  fclose(f);
  return d;
}

int main(int argc, char** argv) {
  const std::string dir = argc > 1 ? argv[1] : "/tmp/zf";
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  auto mutate = [&](std::vector<uint8_t> m) {
    const int nf = 1 + (int)(rnd() % 3);
    for (int k = 0; k < nf; ++k) {
      const size_t p = rnd() % m.size();
      if (rnd() & 1) m[p] ^= (uint8_t)(1u << (rnd() % 8)); else m[p] = (uint8_t)rnd();
    }
    if (rnd() % 8 == 0) m.resize(rnd() % m.size());
    return m;
  };
  struct Target { const char* file; int kind; };
  const Target targets[] = {{"f0.zst", 0}, {"f1.zst", 0}, {"f2.zst", 0}, {"f3.zst", 0}, {"ix.bin", 1}, {"ix.lz4", 2}, {"ix.json", 3}};
  for (const Target& t : targets) {
    const std::vector<uint8_t> seed = load(dir + "/" + t.file);
    if (seed.empty()) continue;
    long ok = 0, err = 0;
    for (int it = 0; it < 20000; ++it) {
      const std::vector<uint8_t> m = mutate(seed);
      try {
        switch (t.kind) {
          case 0: { auto o = zstd_decompress(m.data(), m.size()); if (o.size() > (64u << 20)) { printf("runaway output\n"); return 1; } break; }
          case 1: { auto x = BM25Index::from_bytes(m.data(), m.size()); auto b = x.to_bytes(); break; }
          case 2: { auto x = decompress(Compression::Lz4, m.data(), m.size()); break; }
          default: {
            auto x = PersistedIndex::from_json(reinterpret_cast<const char*>(m.data()), m.size());
            const std::string j = x.to_json();
            bool finite = true;  // (a non-finite number is written as null, which - as in the reference - does not parse back)
            for (const auto& row : x.embeddings) for (float v : row) finite = finite && v == v && v - v == 0.0f;
            if (finite && PersistedIndex::from_json(j.data(), j.size()).to_json() != j) { printf("json writer not idempotent\n"); return 1; }
            break;
          }
        }
        ok++;
      } catch (const Error&) {
        err++;
      }
    }
    printf("%-8s accepted %ld rejected %ld\n", t.file, ok, err);
  }
  // round trips of the two writers on generated data (runs, repeats of earlier content, noise)
  long trips = 0;
  for (int it = 0; it < 3000; ++it) {
    std::vector<uint8_t> d;
    const int parts = 1 + (int)(rnd() % 12);
    for (int k = 0; k < parts; ++k) {
      const int kind = (int)(rnd() % 4);
      const size_t len = rnd() % (it % 50 == 0 ? 200000 : 3000);
      if (kind == 0) for (size_t i = 0; i < len; ++i) d.push_back((uint8_t)rnd());
      else if (kind == 1) d.insert(d.end(), len, (uint8_t)rnd());
      else if (kind == 2 && !d.empty()) { const size_t from = rnd() % d.size(); for (size_t i = 0; i < len; ++i) d.push_back(d[from + i % (d.size() - from)]); }
      else for (size_t i = 0; i < len; ++i) d.push_back((uint8_t)("abcab "[rnd() % 6]));
    }
    const auto z = zstd_compress(d.data(), d.size());
    if (zstd_decompress(z.data(), z.size()) != d) { printf("zstd round trip mismatch at %d\n", it); return 1; }
    const auto l = compress(Compression::Lz4, d.data(), d.size());
    if (decompress(Compression::Lz4, l.data(), l.size()) != d) { printf("lz4 round trip mismatch at %d\n", it); return 1; }
    trips++;
  }
  printf("round trips ok: %ld\n", trips);
  return 0;
}
