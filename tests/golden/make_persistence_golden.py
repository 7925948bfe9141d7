#!/usr/bin/env python
"""Generates tests/golden/bm25_index.{json,zst,lz4.hex}: one small BM25Index in the reference's on-disk formats.

  bm25_index.json     the index content (terms, postings, lengths, parameters) as plain JSON - the expectation
  bm25_index.zst      bincode of that index compressed by LIBZSTD (through pyarrow) at the reference's level 3
                      (src/compressed.rs:43): a third-party frame the library's decoder must read
  bm25_index.lz4      the same bincode as an lz4_flex-style size-prepended block written by the independent pure-Python
                      encoder of tests/test_persistence_format.py (greedy matcher, not the library's)

    python tests/golden/make_persistence_golden.py     # rewrites the fixtures; needs pyarrow
"""
import json
import os
import random
import struct
import sys
import uuid

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.test_persistence_format import lz4_block_encode, write_bm25  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    import pyarrow as pa
    rng = random.Random(20261018)
    vocab = ["retrieval", "dense", "sparse", "fusion", "rank", "query", "index", "vector", "token", "score", "chunk",
             "embed", "cosine", "bm25", "hybrid", "recall", "latency", "batch", "shard", "gpu"]
    ids = [uuid.UUID(int=rng.getrandbits(128)) for _ in range(40)]
    inv, lens = {}, {}
    for c in ids:
        terms = rng.sample(vocab, rng.randint(2, 9))
        tfs = [rng.randint(1, 4) for _ in terms]
        lens[c] = sum(tfs)
        for t, tf in zip(terms, tfs):
            inv.setdefault(t, []).append((c, tf))
    avg = struct.unpack("<f", struct.pack("<f", sum(lens.values()) / len(ids)))[0]
    d = dict(inv=inv, dfs={t: len(pl) for t, pl in inv.items()}, lens=lens, avg=avg, count=len(ids), k1=1.2, b=0.75,
             lowercase=1, stop={"the", "a", "of"})
    k1 = struct.unpack("<f", struct.pack("<f", 1.2))[0]
    raw = write_bm25(d, random.Random(1))
    open(os.path.join(HERE, "bm25_index.zst"), "wb").write(pa.Codec("zstd", compression_level=3).compress(raw, asbytes=True))
    open(os.path.join(HERE, "bm25_index.lz4"), "wb").write(struct.pack("<I", len(raw)) + lz4_block_encode(raw))
    expect = {"inv": {t: [[c.hex, tf] for c, tf in sorted(pl)] for t, pl in sorted(inv.items())},
              "dfs": dict(sorted(d["dfs"].items())), "lens": {c.hex: v for c, v in sorted(lens.items())},
              "avg_bits": struct.unpack("<I", struct.pack("<f", avg))[0], "count": len(ids),
              "k1_bits": struct.unpack("<I", struct.pack("<f", k1))[0], "b": 0.75, "lowercase": 1, "stop": sorted(d["stop"]),
              "bincode_len": len(raw)}
    json.dump(expect, open(os.path.join(HERE, "bm25_index.json"), "w"), indent=1, sort_keys=True)
    print("wrote fixtures:", len(raw), "bytes of bincode")


if __name__ == "__main__":
    main()
