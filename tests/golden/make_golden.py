#!/usr/bin/env python
"""Generates tests/golden/hybrid_small.npz: outputs of the CPU oracle (oracle/trr_oracle.c, pinned to the reference's
known-answer tests by tests/test_oracle_kat.py) on one small seeded hybrid workload.  The reference itself is Rust and
cannot run in this image, so these vectors are ORACLE outputs frozen at commit time: tests/test_golden.py checks that the
oracle still reproduces them (CPU) and that the CUDA path reproduces them (GPU) without calling the oracle.

    python tests/golden/make_golden.py        # rewrites the fixture (only after a deliberate change of the oracle)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

SEED, N, D, V, B, C_, K = 0x5EED0010, 6000, 96, 1200, 12, 50, 10
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hybrid_small.npz")


def workload():
    f, b = O.synth_corpus(SEED, 0, N, D, bf16=True, dups=True)
    q = O.synth_queries(SEED, 0, B, D, N, corpus_bf16=True, dups=True)
    u = q.view(np.uint32)
    q = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32).view(np.float32)   # bf16-exact queries
    cdf = O.zipf_cdf(V)
    doc_off, toks = O.synth_doc_tokens(SEED, cdf, 0, N)
    q_off, q_terms = O.synth_query_terms(SEED, cdf, 0, B)
    return b, q, doc_off, toks, q_off, q_terms


def oracle_outputs():
    b, q, doc_off, toks, q_off, q_terms = workload()
    oix = O.BM25(n_terms=V, doc_off=doc_off, tokens=toks)
    d_ord, d_sc, d_n = O.dense_search_batch(b, q, C_)
    s_ord, s_sc, s_n = oix.search_batch(q_terms, q_off, C_)
    out = {"dense_ord": d_ord, "dense_score": d_sc, "dense_n": d_n, "sparse_ord": s_ord, "sparse_score": s_sc, "sparse_n": s_n}
    for name, strat, param in (("rrf", O.RRF, 60.0), ("linear", O.LINEAR, 0.7), ("dbsf", O.DBSF, 0.0)):
        ids = np.full((B, K), 0xFFFFFFFF, np.uint32)
        fused = np.zeros((B, K), np.float32)
        cnt = np.zeros(B, np.uint32)
        for i in range(B):
            o, fz, dd, ss = O.hybrid_assemble(strat, np.float32(param), (d_ord[i, :d_n[i]], d_sc[i, :d_n[i]]),
                                              (s_ord[i, :s_n[i]], s_sc[i, :s_n[i]]), K)
            cnt[i] = len(o)
            ids[i, :len(o)] = o
            fused[i, :len(o)] = fz
        out[f"{name}_ord"], out[f"{name}_fused"], out[f"{name}_n"] = ids, fused, cnt
    return out


if __name__ == "__main__":
    np.savez_compressed(OUT, **oracle_outputs(), params=np.array([SEED, N, D, V, B, C_, K], np.int64))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
