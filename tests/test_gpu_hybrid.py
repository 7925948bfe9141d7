"""End-to-end hybrid path (dense top-C + BM25 top-C + fusion + take k) and the document-sharded variant.
The multi-GPU flow is exercised on ONE GPU with G logical shards: each shard has its own dense slab and CSR
(global BM25 statistics, local doc ids, global ordinals out), the exchange records are concatenated in device memory
exactly as an all-gather would leave them, and the merge+fusion kernel consumes them.  1/2/4/8 shards must return
identical ids and scores, equal to the oracle's."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from trueno_rag_b200 import shard

pytestmark = pytest.mark.gpu
F32 = np.float32
SEED = 0x5EED0004


@pytest.fixture(scope="module")
def api(ctx):
    from trueno_rag_b200 import api as a
    return a


def bf16_round(x):
    u = np.ascontiguousarray(x, F32).view(np.uint32)
    return ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32).view(F32)


@pytest.fixture(scope="module")
def corpus():
    N, D, V, B = 40000, 256, 8000, 48
    f, b = O.synth_corpus(SEED, 0, N, D, bf16=True, dups=True)
    Q = bf16_round(O.synth_queries(SEED, 0, B, D, N, corpus_bf16=True, dups=True))
    cdf = O.zipf_cdf(V)
    doc_off, toks = O.synth_doc_tokens(SEED, cdf, 0, N)
    q_off, q_terms = O.synth_query_terms(SEED, cdf, 0, B)
    oix = O.BM25(n_terms=V, doc_off=doc_off, tokens=toks)
    return dict(N=N, D=D, V=V, B=B, f=f, b=b, Q=Q, doc_off=doc_off, toks=toks, q_off=q_off, q_terms=q_terms, oix=oix)


def oracle_hybrid(c, strategy, param, C_, k, use_dense=True, use_sparse=True):
    out = []
    for b in range(c["B"]):
        d = O.dense_search(c["b"], c["Q"][b], C_, literal=False) if use_dense else ([], [])
        s = c["oix"].search(c["q_terms"][c["q_off"][b]:c["q_off"][b + 1]], C_) if use_sparse else ([], [])
        out.append(O.hybrid_assemble(strategy, param, d, s, k))
    return out


def assert_hybrid(got, exp):
    o_ord, o_f, o_d, o_s, o_n = got
    for b, (i, f, d, s) in enumerate(exp):
        n = int(o_n[b])
        assert n == len(i), (b, n, len(i))
        assert np.array_equal(o_ord[b, :n], i), (b, o_ord[b, :n], i)
        assert np.array_equal(o_f[b, :n], f), b
        assert np.array_equal(o_d[b, :n], d, equal_nan=True) and np.array_equal(o_s[b, :n], s, equal_nan=True), b


def make_dense(api, ctx, c, lo, hi):
    ix = api.DenseIndex(ctx, c["D"], api.COSINE, api.BF16, base=lo)
    ix.append(c["b"][lo:hi])
    return ix


def make_bm25(api, ctx, c, lo, hi):
    """Shard CSR restricted to docs [lo, hi) with LOCAL doc ids and GLOBAL statistics (N, df, avgdl, idf)."""
    term_off, post_doc, post_tf, doc_len, df = c["oix"].csr()
    keep = (post_doc >= lo) & (post_doc < hi)
    term_of = np.repeat(np.arange(c["V"]), np.diff(term_off).astype(np.int64))
    cnt = np.bincount(term_of[keep], minlength=c["V"])
    s_off = np.zeros(c["V"] + 1, np.uint64)
    np.cumsum(cnt, out=s_off[1:])
    return api.Bm25Device(ctx, hi - lo, s_off, post_doc[keep] - lo, post_tf[keep], doc_len[lo:hi], c["oix"].avgdl,
                          api.bm25_idf_host(c["N"], df), doc_base=lo)


@pytest.mark.parametrize("strategy,param,C_,k", [(O.RRF, 60.0, 50, 10), (O.LINEAR, 0.7, 50, 10), (O.DBSF, 0.0, 20, 40),
                                                  (O.UNION, 0.0, 10, 20), (O.INTERSECTION, 0.0, 50, 10),
                                                  (O.CONVEX, 0.25, 5, 3)])
@pytest.mark.parametrize("mode", [1, 2])
def test_hybrid_single_shard(api, ctx, corpus, strategy, param, C_, k, mode):
    c = corpus
    dense, bm = make_dense(api, ctx, c, 0, c["N"]), make_bm25(api, ctx, c, 0, c["N"])
    dense.set_mode(mode)
    got = api.hybrid_search(dense, bm, c["Q"], c["q_terms"], c["q_off"], C_, strategy, param, k)
    assert_hybrid(got, oracle_hybrid(c, strategy, np.float32(param), C_, k))
    dense.close(); bm.close()


def test_hybrid_dense_only_and_sparse_only(api, ctx, corpus):
    c = corpus
    dense, bm = make_dense(api, ctx, c, 0, c["N"]), make_bm25(api, ctx, c, 0, c["N"])
    got = api.hybrid_search(dense, bm, c["Q"], c["q_terms"], c["q_off"], 50, O.RRF, 60.0, 10, use_sparse=False)
    assert_hybrid(got, oracle_hybrid(c, O.RRF, 60.0, 50, 10, use_sparse=False))
    got = api.hybrid_search(dense, bm, c["Q"], c["q_terms"], c["q_off"], 50, O.RRF, 60.0, 10, use_dense=False)
    assert_hybrid(got, oracle_hybrid(c, O.RRF, 60.0, 50, 10, use_dense=False))
    dense.close(); bm.close()


@pytest.mark.parametrize("G", [1, 2, 4, 8])
@pytest.mark.parametrize("strategy,param", [(O.RRF, 60.0), (O.LINEAR, 0.7)])
def test_sharded_flow_on_logical_shards(api, ctx, corpus, G, strategy, param):
    c = corpus
    B, C_, k = c["B"], 50, 10
    rec_bytes = api.exchange_bytes(B, C_)
    assert rec_bytes == shard.exchange_words(B, C_) * 4
    gathered = torch.zeros(G * rec_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()                                         # the library enqueues on its own stream
    handles = []
    for g in range(G):
        lo, hi = shard.shard_range(c["N"], g, G)
        dense, bm = make_dense(api, ctx, c, lo, hi), make_bm25(api, ctx, c, lo, hi)
        dense.set_mode(2 if g % 2 == 0 else 1)                      # mix both dense kernels across shards
        api.hybrid_local(dense, bm, c["Q"], c["q_terms"], c["q_off"], C_, gathered.data_ptr() + g * rec_bytes)
        handles += [dense, bm]
    torch.cuda.synchronize()
    got = api.hybrid_merge(ctx, gathered.data_ptr(), G, B, C_, strategy, param, k)
    assert_hybrid(got, oracle_hybrid(c, strategy, np.float32(param), C_, k))
    for h in handles:
        h.close()


def test_cfg5_shape_linear_top100_eight_shards(api, ctx):
    """BASELINE.json configs[4] at reduced N: 4096-dim bf16, C = 100 (set explicitly, src/retrieve.rs:128-131), linear
    fusion dense_weight 0.7 (examples/hybrid_search.rs:27-29), top-100, 8 document shards, tensor-core dense pass."""
    N, D, V, B, C_, k, G = 24000, 4096, 6000, 40, 100, 100, 8
    f, b = O.synth_corpus(SEED + 5, 0, N, D, bf16=True)
    Q = bf16_round(O.synth_queries(SEED + 5, 0, B, D, N, corpus_bf16=True))
    cdf = O.zipf_cdf(V)
    doc_off, toks = O.synth_doc_tokens(SEED + 5, cdf, 0, N)
    q_off, q_terms = O.synth_query_terms(SEED + 5, cdf, 0, B)
    oix = O.BM25(n_terms=V, doc_off=doc_off, tokens=toks)
    c = dict(N=N, D=D, V=V, B=B, f=f, b=b, Q=Q, q_off=q_off, q_terms=q_terms, oix=oix)
    rec_bytes = api.exchange_bytes(B, C_)
    gathered = torch.zeros(G * rec_bytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    handles = []
    for g in range(G):
        lo, hi = shard.shard_range(N, g, G)
        dense, bm = make_dense(api, ctx, c, lo, hi), make_bm25(api, ctx, c, lo, hi)
        dense.set_mode(2)
        api.hybrid_local(dense, bm, Q, q_terms, q_off, C_, gathered.data_ptr() + g * rec_bytes)
        assert dense.stats().mode_used == 2
        handles += [dense, bm]
    torch.cuda.synchronize()
    got = api.hybrid_merge(ctx, gathered.data_ptr(), G, B, C_, O.LINEAR, 0.7, k)
    assert_hybrid(got, oracle_hybrid(c, O.LINEAR, np.float32(0.7), C_, k))
    for h in handles:
        h.close()


def test_concurrent_searches_on_one_handle(api, ctx):
    """The header promises that search entry points are re-entrant on one handle (the reference's query methods take &self
    and SparseIndex is Send + Sync, src/index.rs:8): four host threads hammer trr_hybrid_search / trr_dense_search /
    trr_bm25_search on the same handles and every result must equal the oracle's."""
    import threading
    n, d, n_terms, B, Cn, K = 30000, 128, 4000, 24, 20, 7
    f, bf = O.synth_corpus(SEED + 21, 0, n, d, bf16=True)
    rows = (bf.astype(np.uint32) << 16).view(np.float32)
    cdf = O.zipf_cdf(n_terms)
    doc_off, toks = O.synth_doc_tokens(SEED + 21, cdf, 0, n)
    oix = O.BM25(n_terms=n_terms, doc_off=doc_off, tokens=toks)
    term_off, post_doc, post_tf, doc_len, df = oix.csr()
    dense = api.DenseIndex(ctx, d, api.COSINE, api.BF16)
    dense.append(bf)
    bm = api.Bm25Device(ctx, n, term_off, post_doc, post_tf, doc_len, oix.avgdl, api.bm25_idf_host(n, df))
    errors = []

    def worker(tid):
        try:
            for it in range(6):
                q = O.synth_queries(SEED + 100 * tid + it, 0, B, d, n, corpus_bf16=True)
                q_off, q_terms = O.synth_query_terms(SEED + 100 * tid + it, cdf, 0, B)
                d_exp = O.dense_search_batch(rows, q, Cn)
                s_exp = oix.search_batch(q_terms, q_off, Cn)
                o_ord, o_f, o_d, o_s, o_n = api.hybrid_search(dense, bm, q, q_terms, q_off, Cn, api.RRF, 60.0, K)
                for b in range(B):
                    i, fsc, dd, ss = O.hybrid_assemble(O.RRF, 60.0, (d_exp[0][b, :d_exp[2][b]], d_exp[1][b, :d_exp[2][b]]),
                                                       (s_exp[0][b, :s_exp[2][b]], s_exp[1][b, :s_exp[2][b]]), K)
                    m = int(o_n[b])
                    assert m == len(i) and np.array_equal(o_ord[b, :m], i) and np.array_equal(o_f[b, :m], fsc)
                g = dense.search(q, Cn)
                assert np.array_equal(g[2], d_exp[2]) and np.array_equal(g[0], d_exp[0]) and np.array_equal(g[1], d_exp[1])
                s = bm.search(q_terms, q_off, Cn)
                assert np.array_equal(s[2], s_exp[2])
                for b in range(B):
                    m = int(s_exp[2][b])
                    assert np.array_equal(s[0][b, :m], s_exp[0][b, :m]) and np.array_equal(s[1][b, :m], s_exp[1][b, :m])
        except Exception as ex:  # noqa: BLE001
            errors.append((tid, repr(ex)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    dense.close(); bm.close()
    assert not errors, errors


def test_hybrid_input_validation(api, ctx):
    """A query of the wrong width raises DimensionMismatch (VectorStore::search, src/index.rs:387-392) instead of reading
    past the buffer; a 1-D query is one query; q_off must have B + 1 entries."""
    from trueno_rag_b200 import _lib
    dense = api.DenseIndex(ctx, 32)
    dense.append(np.random.default_rng(0).standard_normal((100, 32)).astype(np.float32))
    with pytest.raises(api.TrrError) as e:
        api.hybrid_search(dense, None, np.zeros((3, 31), np.float32), None, None, 5, api.RRF, 60.0, 3, use_sparse=False)
    assert e.value.status == _lib.TRR_ERR_DIM_MISMATCH
    out = api.hybrid_search(dense, None, np.ones(32, np.float32), None, None, 5, api.RRF, 60.0, 3, use_sparse=False)
    assert out[0].shape == (1, 3) and out[4][0] == 3
    dense.close()


@pytest.mark.parametrize("exchange", [0, 1])
def test_group_single_rank_matches_oracle(api, ctx, corpus, exchange):
    """trr_group_* + trr_hybrid_search_sharded with a group of one rank (what a single-GPU caller of the sharded entry point
    gets): equal to the oracle, for both exchanges, over consecutive calls (the exchange buffers alternate)."""
    c = corpus
    group = api.Group(ctx, 0, 1, api.group_unique_id(), exchange)
    dense, bm = make_dense(api, ctx, c, 0, c["N"]), make_bm25(api, ctx, c, 0, c["N"])
    assert int(group.allreduce_u64(np.array([41], np.uint64))[0]) == 41
    for strategy, param, C_, k in ((O.RRF, 60.0, 50, 10), (O.LINEAR, 0.7, 50, 10), (O.UNION, 0.0, 10, 20)):
        got = group.search(dense, bm, c["Q"], c["q_terms"], c["q_off"], C_, strategy, param, k)
        assert_hybrid(got, oracle_hybrid(c, strategy, np.float32(param), C_, k))
    group.close(); dense.close(); bm.close()


@pytest.mark.parametrize("exchange", [0, 1])
def test_group_two_ranks(exchange, tmp_path):
    """Two processes, two GPUs, one shard each (skipped on a single-GPU box): tests/_group_worker.py."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    id_file = str(tmp_path / "group_id")
    procs = [subprocess.Popen([sys.executable, os.path.join(root, "tests", "_group_worker.py"), str(r), "2", id_file,
                               str(exchange)], cwd=root, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"rank {r}:" in out and "OK" in out, out[-3000:]
