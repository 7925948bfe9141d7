"""Persistence -> device load: snapshots of the device stores restore bit-identical search results."""
import os

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
SEED = 0x5EED0006


@pytest.fixture(scope="module")
def api(ctx):
    from trueno_rag_b200 import api as a
    return a


def bf16_round(x):
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    return ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("dtype,metric", [(0, 0), (1, 0), (0, 1), (1, 2)])
def test_dense_snapshot_round_trip(api, ctx, tmp_path, dtype, metric):
    n, d = 21000, 256
    f, b = O.synth_corpus(SEED, 0, n, d, bf16=bool(dtype), dups=True)
    rows = b if dtype else f
    ix = api.DenseIndex(ctx, d, metric, dtype, base=5000)
    ix.append(rows)
    for dead in (3, 77, n - 1):
        ix.remove(dead)
    Q = O.synth_queries(SEED, 0, 40, d, n, corpus_bf16=bool(dtype), dups=True)
    if dtype:
        Q = bf16_round(Q)
    before = ix.search(Q, 25)
    path = os.path.join(tmp_path, "dense.trr")
    ix.save(path)
    ix.close()
    ix2 = api.DenseIndex.load(ctx, path)
    assert (ix2.dim, ix2.metric, ix2.dtype) == (d, metric, dtype) and len(ix2) == n - 3
    for mode in ((1, 2) if metric != 1 else (1,)):
        ix2.set_mode(mode)
        after = ix2.search(Q, 25)
        assert all(np.array_equal(x, y) for x, y in zip(before, after)), mode
    alive = np.ones(n, bool)
    alive[[3, 77, n - 1]] = False
    eo, es, en = O.dense_search_batch(rows, Q, 25, metric=metric, alive=alive)
    assert np.array_equal(after[0], eo + 5000) and np.array_equal(after[1], es)
    ix2.append(rows[:10])                                              # a restored store keeps accepting inserts
    assert len(ix2) == n - 3 + 10
    ix2.close()


def test_bm25_snapshot_round_trip(api, ctx, tmp_path):
    cdf = O.zipf_cdf(4000)
    doc_off, toks = O.synth_doc_tokens(SEED, cdf, 0, 50000)
    oix = O.BM25(n_terms=4000, doc_off=doc_off, tokens=toks)
    term_off, post_doc, post_tf, doc_len, df = oix.csr()
    dev = api.Bm25Device(ctx, 50000, term_off, post_doc, post_tf, doc_len, oix.avgdl, api.bm25_idf_host(50000, df),
                         doc_base=700)
    q_off, q_terms = O.synth_query_terms(SEED, cdf, 0, 64)
    before = dev.search(q_terms, q_off, 50)
    path = os.path.join(tmp_path, "bm25.trr")
    dev.save(path)
    n_post = dev.n_postings
    dev.close()
    dev2 = api.Bm25Device.load(ctx, path)
    assert dev2.n_postings == n_post
    after = dev2.search(q_terms, q_off, 50)
    assert all(np.array_equal(x, y) for x, y in zip(before, after))
    eo, es, en = oix.search_batch(q_terms, q_off, 50)
    assert np.array_equal(after[2], en)
    for b in range(64):
        m = int(en[b])
        assert np.array_equal(after[0][b, :m], eo[b, :m] + 700) and np.array_equal(after[1][b, :m], es[b, :m])
    dev2.close()


def test_bm25_remove_save_load_append_keeps_dead_postings_accounted(api, ctx, tmp_path):
    """remove -> save -> load -> append: the snapshot carries the number of tombstoned postings; without it the threshold
    bootstrap after the append takes a term's posting count (dead ones included) for its live document count and drops valid
    hits.  The restored index must keep returning exactly what an index built from the surviving + appended documents returns."""
    V, n0, n1 = 800, 30000, 4000
    cdf = O.zipf_cdf(V)
    doc_off, toks = O.synth_doc_tokens(SEED + 41, cdf, 0, n0 + n1)
    first = O.BM25(n_terms=V, doc_off=doc_off[:n0 + 1], tokens=toks[:int(doc_off[n0])])
    t_off, pd, ptf, dl, df = first.csr()
    dev = api.Bm25Device(ctx, n0, t_off, pd, ptf, dl, first.avgdl, api.bm25_idf_host(n0, df))
    # remove nearly every document that contains term t_star: afterwards its live df (3) is far below k while its posting
    # count (dead postings included) stays far above
    t_star = int(np.argmin(np.abs(df.astype(np.int64) - 400)))
    holders = pd[int(t_off[t_star]):int(t_off[t_star + 1])]
    assert len(holders) > 150
    gone = np.unique(holders[3:])
    keep = np.setdiff1d(np.arange(n0, dtype=np.uint32), gone)
    lens = np.diff(doc_off).astype(np.int64)

    def oracle_over(ids):
        off = np.zeros(len(ids) + 1, np.uint64)
        np.cumsum(lens[ids], out=off[1:])
        tk = np.concatenate([toks[int(doc_off[d]):int(doc_off[d + 1])] for d in ids])
        return O.BM25(n_terms=V, doc_off=off, tokens=tk)

    kept = oracle_over(keep)
    df_kept = kept.csr()[4]
    assert dev.remove(gone, kept.avgdl, api.bm25_idf_host(len(keep), df_kept)) > 0
    path = os.path.join(tmp_path, "bm25_dead.trr")
    dev.save(path)
    dev.close()
    dev2 = api.Bm25Device.load(ctx, path)
    # append the documents n0 .. n0 + n1 with the statistics of (survivors + appended)
    ids_all = np.concatenate([keep, np.arange(n0, n0 + n1, dtype=np.uint32)])
    allo = oracle_over(ids_all)
    df_all = allo.csr()[4]
    sub_off = (doc_off[n0:n0 + n1 + 1] - doc_off[n0]).astype(np.uint64)
    delta = O.BM25(n_terms=V, doc_off=sub_off, tokens=toks[int(doc_off[n0]):int(doc_off[n0 + n1])])
    d_off, d_pd, d_ptf, d_dl, _ = delta.csr()
    dev2.append(n1, d_off, d_pd, d_ptf, d_dl, allo.avgdl, api.bm25_idf_host(len(ids_all), df_all))
    qs = [[t_star], [t_star, 5], [t_star, t_star, 7, 9], [11, 12, 13]]
    q_off = np.cumsum([0] + [len(q) for q in qs]).astype(np.uint32)
    q_terms = np.array([x for q in qs for x in q], np.uint32)
    for k in (5, 50):
        ords, scores, cnt = dev2.search(q_terms, q_off, k)
        eo, es, en = allo.search_batch(q_terms, q_off, k)
        assert np.array_equal(cnt, en), (cnt, en)
        for b in range(len(qs)):
            m = int(en[b])
            assert np.array_equal(ords[b, :m], ids_all[eo[b, :m]]), (k, b)
            assert np.array_equal(scores[b, :m], es[b, :m]), (k, b)
    dev2.close()


def test_snapshot_rejects_foreign_and_truncated_files(api, ctx, tmp_path):
    bad = os.path.join(tmp_path, "bad.trr")
    open(bad, "wb").write(b"not a snapshot at all, just some bytes" * 4)
    with pytest.raises(api.TrrError):
        api.DenseIndex.load(ctx, bad)
    with pytest.raises(api.TrrError):
        api.Bm25Device.load(ctx, bad)
    with pytest.raises(api.TrrError):
        api.DenseIndex.load(ctx, os.path.join(tmp_path, "missing.trr"))
    ix = api.DenseIndex(ctx, 64)
    ix.append(np.ones((5000, 64), np.float32))
    good = os.path.join(tmp_path, "good.trr")
    ix.save(good)
    ix.close()
    data = open(good, "rb").read()
    open(bad, "wb").write(data[: len(data) // 2])
    with pytest.raises(api.TrrError):
        api.DenseIndex.load(ctx, bad)


def test_bm25_snapshot_with_corrupt_indices_is_refused(api, ctx, tmp_path):
    """A file that passes the header checks but whose skip rows / document ids would send the search kernels out of
    bounds is refused at load time (the loaded arrays are device indices)."""
    cdf = O.zipf_cdf(500)
    doc_off, toks = O.synth_doc_tokens(SEED, cdf, 0, 40000)
    oix = O.BM25(n_terms=500, doc_off=doc_off, tokens=toks)
    term_off, post_doc, post_tf, doc_len, df = oix.csr()
    dev = api.Bm25Device(ctx, 40000, term_off, post_doc, post_tf, doc_len, oix.avgdl, api.bm25_idf_host(40000, df))
    good = os.path.join(tmp_path, "good.trr")
    dev.save(good)
    n_post = dev.n_postings
    dev.close()
    data = bytearray(open(good, "rb").read())
    header = 8 + 6 * 4 + 8 + 8 + 2 * 4                      # Bm25SnapHeader (csrc/capi.cu)
    ok = api.Bm25Device.load(ctx, good)
    assert ok.n_postings == n_post
    ok.close()
    # a document id beyond the shard
    bad_doc = bytearray(data)
    bad_doc[header + 8 * 1000: header + 8 * 1000 + 4] = (0x7FFFFFF0).to_bytes(4, "little")
    p1 = os.path.join(tmp_path, "bad_doc.trr")
    open(p1, "wb").write(bad_doc)
    with pytest.raises(api.TrrError):
        api.Bm25Device.load(ctx, p1)
    # a skip entry beyond the postings (second column of the row of term 3)
    n_ranges = (40000 + 32767) // 32768
    skip0 = header + 8 * (n_post + 2)
    bad_skip = bytearray(data)
    off = skip0 + 4 * (3 * (n_ranges + 1) + 1)
    bad_skip[off: off + 4] = (0xFFFFFF00).to_bytes(4, "little")
    p2 = os.path.join(tmp_path, "bad_skip.trr")
    open(p2, "wb").write(bad_skip)
    with pytest.raises(api.TrrError):
        api.Bm25Device.load(ctx, p2)
