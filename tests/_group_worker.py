"""Worker of tests/test_gpu_hybrid.py::test_group_two_ranks: one process per GPU, rank r holds shard r of a small seeded
corpus and calls the sharded entry point (reference src/retrieve.rs:175-220 over the whole corpus in ONE call per rank);
every rank compares what it got with the oracle's result over the unsharded corpus.

    python tests/_group_worker.py <rank> <world> <id file> <exchange: 0 nccl | 1 peer>
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O                      # noqa: E402  (checker only)
from trueno_rag_b200 import api, shard             # noqa: E402
from tests.test_gpu_hybrid import (SEED, assert_hybrid, bf16_round, make_bm25, make_dense, oracle_hybrid)  # noqa: E402


def main():
    rank, world, id_file, exchange = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
    N, D, V, B = 30000, 128, 6000, 40
    f, b = O.synth_corpus(SEED, 0, N, D, bf16=True, dups=True)
    Q = bf16_round(O.synth_queries(SEED, 0, B, D, N, corpus_bf16=True, dups=True))
    cdf = O.zipf_cdf(V)
    doc_off, toks = O.synth_doc_tokens(SEED, cdf, 0, N)
    q_off, q_terms = O.synth_query_terms(SEED, cdf, 0, B)
    c = dict(N=N, D=D, V=V, B=B, f=f, b=b, Q=Q, doc_off=doc_off, toks=toks, q_off=q_off, q_terms=q_terms,
             oix=O.BM25(n_terms=V, doc_off=doc_off, tokens=toks))
    ctx = api.Context(rank)
    if rank == 0:
        uid = api.group_unique_id()
        with open(id_file + ".tmp", "wb") as fh:
            fh.write(uid)
        os.replace(id_file + ".tmp", id_file)
    else:
        t0 = time.time()
        while not os.path.exists(id_file):
            if time.time() - t0 > 120:
                raise RuntimeError("no rendezvous id")
            time.sleep(0.05)
        uid = open(id_file, "rb").read()
    group = api.Group(ctx, rank, world, uid, exchange)
    lo, hi = shard.shard_range(N, rank, world)
    dense, bm = make_dense(api, ctx, c, lo, hi), make_bm25(api, ctx, c, lo, hi)
    dense.set_mode(2 if rank % 2 == 0 else 1)        # both dense kernels across the ranks
    tot = group.allreduce_u64(np.array([hi - lo, rank], np.uint64))
    assert int(tot[0]) == N and int(tot[1]) == world * (world - 1) // 2, tot
    mx = group.allreduce_u64(np.array([rank + 7], np.uint64), op_max=True)
    assert int(mx[0]) == world + 6, mx
    # consecutive calls alternate the exchange buffers: three strategies, twice each
    for strategy, param, C_, k in ((O.RRF, 60.0, 50, 10), (O.LINEAR, 0.7, 50, 10), (O.DBSF, 0.0, 20, 40)) * 2:
        got = group.search(dense, bm, Q, q_terms, q_off, C_, strategy, param, k)
        assert_hybrid(got, oracle_hybrid(c, strategy, np.float32(param), C_, k))
    got = group.search(dense, bm, Q, q_terms, q_off, 50, O.RRF, 60.0, 10, use_sparse=False)
    assert_hybrid(got, oracle_hybrid(c, O.RRF, np.float32(60.0), 50, 10, use_sparse=False))
    print(f"rank {rank}: exchange in use {group.exchange}, OK", flush=True)
    group.sync()
    group.close(); dense.close(); bm.close(); ctx.close()


if __name__ == "__main__":
    main()
