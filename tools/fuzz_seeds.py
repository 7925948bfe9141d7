"""Seed files for tools/fuzz_host_readers.cpp:  python tools/fuzz_seeds.py <dir>"""
import json
import os
import random
import sys

import numpy as np
import pyarrow as pa

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from trueno_rag_b200 import api  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/zf"
os.makedirs(out, exist_ok=True)
rng = random.Random(5)
words = [bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(2, 9))) for _ in range(300)]
text = b" ".join(rng.choice(words) for _ in range(8000))
skew = bytes(rng.choice(b"aaaaaaaabbbbccd") for _ in range(30000))
period = (b"0123456789abcdef" * 3 + b"XY") * 800
for i, (d, level) in enumerate([(text, 3), (skew, 19), (period, 1), (text[:3000], 9)]):
    open(f"{out}/f{i}.zst", "wb").write(pa.Codec("zstd", compression_level=level).compress(d, asbytes=True))
ix = api.BM25Index()
for i in range(40):
    ix.add(api.Chunk(" ".join("w%02d" % rng.randrange(50) for _ in range(rng.randint(3, 15)))))
open(f"{out}/ix.bin", "wb").write(ix.to_bytes())
open(f"{out}/ix.lz4", "wb").write(ix.to_compressed_bytes(api.Compression.Lz4))
E = np.random.default_rng(0).standard_normal((6, 5)).astype(np.float32)
doc = {"chunks": [{"content": 'h\u00e9llo "w"\n\u2603 \U0001F600 %d' % i, "title": None, "source": "a.md"} for i in range(6)],
       "embeddings": [[float(x) for x in r] for r in E], "dimension": 5, "embedder_type": "tfidf", "model_name": None,
       "extra": {"x": [1, 2, {"y": None}]}}
open(f"{out}/ix.json", "w").write(json.dumps(doc, ensure_ascii=True))
print("seeds in", out)
