"""world_size-2 test (gloo, CPU) of the document-sharded path's host plumbing: contiguous shard ranges, global BM25
statistics, exchange-record layout and the all-gather.  The shard-local lists and the final merge are computed by the
ORACLE here (no GPU in this tier); the GPU tests run the same flow with the CUDA kernels on logical shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from trueno_rag_b200 import shard

N, D, B, C_, K = 600, 32, 6, 10, 5
SEED = 0x5EED0001


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    rows, _ = O.synth_corpus(SEED, 0, N, D)
    rows[N // 2 + 3] = rows[5]                       # an exact tie that straddles the shard boundary
    q = O.synth_queries(SEED, 0, B, D, N)
    q[0] = rows[5]
    cdf = O.zipf_cdf(300)
    doc_off, toks = O.synth_doc_tokens(SEED, cdf, 0, N)
    q_off, q_terms = O.synth_query_terms(SEED, cdf, 0, B)
    return rows, q, doc_off, toks, q_off, q_terms


def _unsharded():
    rows, q, doc_off, toks, q_off, q_terms = _inputs()
    ix = O.BM25(n_terms=300, doc_off=doc_off, tokens=toks)
    out = []
    for b in range(B):
        d = O.dense_search(rows, q[b], C_)
        s = ix.search(q_terms[q_off[b]:q_off[b + 1]], C_)
        out.append(O.hybrid_assemble(O.RRF, 60.0, d, s, K))
    return out


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows, q, doc_off, toks, q_off, q_terms = _inputs()
        lo, hi = shard.shard_range(N, rank, world)
        # shard-local dense top-C with GLOBAL ordinals
        d_ord = np.zeros((B, C_), np.uint32); d_sc = np.zeros((B, C_), np.float32); d_n = np.zeros(B, np.uint32)
        s_ord = np.zeros((B, C_), np.uint32); s_sc = np.zeros((B, C_), np.float32); s_n = np.zeros(B, np.uint32)
        # shard-local BM25 must use GLOBAL statistics (N, df, avgdl): build the global index, keep this shard's docs
        gix = O.BM25(n_terms=300, doc_off=doc_off, tokens=toks)
        for b in range(B):
            o, s = O.dense_search(rows[lo:hi], q[b], C_)
            d_n[b] = len(o); d_ord[b, :len(o)] = o + lo; d_sc[b, :len(o)] = s
            o, s = gix.search(q_terms[q_off[b]:q_off[b + 1]], N)        # all scored docs, then restrict to the shard
            keep = (o >= lo) & (o < hi)
            o, s = o[keep][:C_], s[keep][:C_]
            s_n[b] = len(o); s_ord[b, :len(o)] = o; s_sc[b, :len(o)] = s
        rec = shard.pack_exchange((d_ord, d_sc, d_n), (s_ord, s_sc, s_n), B, C_)
        t = torch.from_numpy(rec.view(np.int32).copy())
        gathered = shard.all_gather_records(t, world).numpy().view(np.uint32)
        assert gathered.shape == (world, shard.exchange_words(B, C_))
        # merge (oracle): global top-C per source, then fuse
        results = []
        for b in range(B):
            per_src = []
            for src in range(2):
                pairs = []
                for g in range(world):
                    ords, scores, n = shard.unpack_exchange(gathered[g], B, C_)
                    pairs += [(float(scores[src, b, i]), int(ords[src, b, i])) for i in range(n[src, b])]
                pairs.sort(key=lambda p: (-p[0], p[1]))
                pairs = pairs[:C_]
                per_src.append(([p[1] for p in pairs], [p[0] for p in pairs]))
            results.append(O.hybrid_assemble(O.RRF, 60.0, per_src[0], per_src[1], K))
        ret[rank] = [tuple(np.asarray(x).tolist() for x in r) for r in results]
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sharded_flow_matches_unsharded_oracle():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    ref = _unsharded()
    for rank in range(world):
        for b in range(B):
            got = ret[rank][b]
            assert got[0] == ref[b][0].tolist()                               # ids, bit-exact
            assert got[1] == ref[b][1].tolist()                               # fused scores
            for gi, ri in ((got[2], ref[b][2]), (got[3], ref[b][3])):          # per-source scores (NaN-aware)
                assert np.array_equal(np.array(gi, np.float32), ri, equal_nan=True)
