"""BM25 parity: CUDA posting-list kernel (through the C ABI) vs the oracle.  Bit-exact ids and scores."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
F32 = np.float32
SEED = 0x5EED0003


@pytest.fixture(scope="module")
def api(ctx):
    from trueno_rag_b200 import api as a
    return a


def build(api, ctx, oix, n_docs, doc_base=0, k1=1.2, b=0.75):
    term_off, post_doc, post_tf, doc_len, df = oix.csr()
    return api.Bm25Device(ctx, n_docs, term_off, post_doc, post_tf, doc_len, oix.avgdl, api.bm25_idf_host(n_docs, df),
                          k1, b, doc_base)


def compare(dev, oix, q_terms, q_off, k, base=0):
    ords, scores, n = dev.search(q_terms, q_off, k)
    eo, es, en = oix.search_batch(q_terms, q_off, k)
    assert np.array_equal(n, en), (n, en)
    for b in range(len(q_off) - 1):
        m = int(n[b])
        assert np.array_equal(ords[b, :m], eo[b, :m] + base), (b, ords[b, :m], eo[b, :m])
        assert np.array_equal(scores[b, :m], es[b, :m]), (b, scores[b, :m] - es[b, :m])


def test_impacts_are_bit_exact(api, ctx):
    rng = np.random.default_rng(0)
    docs = [list(rng.integers(0, 50, int(rng.integers(1, 40)))) for _ in range(500)]
    oix = O.BM25(docs, 50, 1.5, 0.6)
    dev = build(api, ctx, oix, 500, k1=1.5, b=0.6)
    term_off, post_doc, post_tf, doc_len, df = oix.csr()
    exp = np.zeros(oix.n_postings, F32)
    for t in range(50):
        for p in range(int(term_off[t]), int(term_off[t + 1])):
            exp[p] = O.bm25_score_term(int(post_tf[p]), int(df[t]), 500, int(doc_len[post_doc[p]]), oix.avgdl, 1.5, 0.6)
    assert np.array_equal(dev.impacts(), exp)
    dev.close()


def test_random_small_indexes_vs_literal_oracle(api, ctx):
    rng = np.random.default_rng(1)
    for trial in range(12):
        n_docs, n_terms = int(rng.integers(1, 400)), int(rng.integers(1, 60))
        docs = [list(rng.integers(0, n_terms, int(rng.integers(0, 30)))) for _ in range(n_docs)]
        oix = O.BM25(docs, n_terms)
        dev = build(api, ctx, oix, n_docs)
        qs = [list(rng.integers(0, n_terms + 3, int(rng.integers(0, 9)))) for _ in range(7)]
        qs = [[x if x < n_terms else 0xFFFFFFFF for x in q] for q in qs]
        q_off = np.cumsum([0] + [len(q) for q in qs]).astype(np.uint32)
        q_terms = np.array([x for q in qs for x in q], np.uint32)
        for k in (1, 7, 1000):
            ords, scores, n = dev.search(q_terms, q_off, k)
            for b, q in enumerate(qs):
                lo, ls = oix.search(q, k, literal=True)                 # the reference's literal algorithm
                assert n[b] == len(lo)
                assert np.array_equal(ords[b, :n[b]], lo) and np.array_equal(scores[b, :n[b]], ls)
        dev.close()


@pytest.mark.parametrize("n_docs,n_terms", [(3000, 500), (40000, 5000), (70000, 20000)])
def test_zipf_corpus(api, ctx, n_docs, n_terms):
    cdf = O.zipf_cdf(n_terms)
    doc_off, toks = O.synth_doc_tokens(SEED, cdf, 0, n_docs)
    oix = O.BM25(n_terms=n_terms, doc_off=doc_off, tokens=toks)
    dev = build(api, ctx, oix, n_docs, doc_base=123456)
    q_off, q_terms = O.synth_query_terms(SEED, cdf, 0, 40)
    for k in (10, 100):
        compare(dev, oix, q_terms, q_off, k, base=123456)
    dev.close()


def test_duplicate_terms_and_massive_ties(api, ctx):
    # every document has the same length and term 0 once -> thousands of exact score ties
    n_docs = 40000
    docs = [[0, 1 + (i % 7), 8 + (i % 3)] for i in range(n_docs)]
    oix = O.BM25(docs, 11)
    dev = build(api, ctx, oix, n_docs)
    qs = [[0], [0, 0], [0, 3, 0, 9], [5, 5, 5], [10, 0xFFFFFFFF, 2]]
    q_off = np.cumsum([0] + [len(q) for q in qs]).astype(np.uint32)
    q_terms = np.array([x for q in qs for x in q], np.uint32)
    for k in (1, 50, 1000):
        compare(dev, oix, q_terms, q_off, k)
    # thousands of documents tie exactly at the k-th score: neither integer fast pass can prove its selection, so these
    # queries must have gone through both fallback levels (32-bit cells, then the exact kernel)
    compare(dev, oix, q_terms, q_off, 50)
    st = dev.stats()
    assert st.n_guard_fallbacks >= 1 and st.n_exact_fallbacks >= 1, (st.n_guard_fallbacks, st.n_exact_fallbacks)
    dev.close()


def test_fast_pass_proves_ordinary_queries(api, ctx):
    # on a Zipfian corpus the 16-bit fast pass + exact re-scoring must settle (nearly) every query without a fallback
    n_docs, n_terms = 120000, 30000
    cdf = O.zipf_cdf(n_terms)
    doc_off, toks = O.synth_doc_tokens(SEED + 7, cdf, 0, n_docs)
    oix = O.BM25(n_terms=n_terms, doc_off=doc_off, tokens=toks)
    dev = build(api, ctx, oix, n_docs)
    q_off, q_terms = O.synth_query_terms(SEED + 7, cdf, 0, 300)
    compare(dev, oix, q_terms, q_off, 50)
    st = dev.stats()
    assert st.mode_used == 2 and st.n_exact_fallbacks == 0 and st.n_guard_fallbacks <= 30, (st.mode_used, st.n_guard_fallbacks, st.n_exact_fallbacks)
    dev.close()


def test_long_posting_lists_exceed_the_stage(api, ctx):
    # terms 0..3 occur in (almost) every document of every 16K range: segments longer than the 2048-entry stage
    rng = np.random.default_rng(2)
    n_docs = 50000
    docs = []
    for i in range(n_docs):
        d = [0, 1, 2] + ([3] if i % 5 else []) + list(rng.integers(4, 200, int(rng.integers(0, 6))))
        docs.append(d)
    oix = O.BM25(docs, 200)
    dev = build(api, ctx, oix, n_docs)
    qs = [[0, 1, 2, 3], [3, 50, 0], [150, 151, 152, 2, 153, 154], [0] * 12]
    q_off = np.cumsum([0] + [len(q) for q in qs]).astype(np.uint32)
    q_terms = np.array([x for q in qs for x in q], np.uint32)
    compare(dev, oix, q_terms, q_off, 100)
    dev.close()


def test_many_query_terms_and_limits(api, ctx):
    cdf = O.zipf_cdf(3000)
    doc_off, toks = O.synth_doc_tokens(SEED + 1, cdf, 0, 20000)
    oix = O.BM25(n_terms=3000, doc_off=doc_off, tokens=toks)
    dev = build(api, ctx, oix, 20000)
    rng = np.random.default_rng(3)
    q = rng.integers(0, 3000, 512).astype(np.uint32)                    # 512 terms: the per-query limit
    compare(dev, oix, q, np.array([0, 512], np.uint32), 20)
    with pytest.raises(api.TrrError):
        dev.search(np.zeros(513, np.uint32), np.array([0, 513], np.uint32), 5)
    o, s, n = dev.search(np.zeros(0, np.uint32), np.array([0, 0, 0], np.uint32), 5)   # two empty queries
    assert list(n) == [0, 0]
    dev.close()


def test_bench_shape_topic_keywords(api, ctx):                          # benches/retrieval.rs:45-69
    from oracle.tokenizer import TextIndex
    ti = TextIndex()
    for i in range(1000):
        ti.add(f"Document {i} about topic {i % 100} with keywords")
    oix = ti.build()
    dev = build(api, ctx, oix, 1000)
    q = np.array(ti.query_ids("topic keywords"), np.uint32)
    for k in (10, 100):
        compare(dev, oix, q, np.array([0, len(q)], np.uint32), k)
    dev.close()


@pytest.mark.parametrize("k1,b", [(0.0, 0.75), (1.2, 0.0), (1.2, 1.0), (2.5, 0.3), (-0.5, 0.75), (1.2, -2.0)])
def test_unusual_parameters(api, ctx, k1, b):
    """BM25Index::with_params (src/index.rs:78-84) accepts any floats: k1 = 0 makes every impact idf, negative values can
    make impacts negative or non-finite (those documents are dropped by `score > 0.0`, src/index.rs:236) and switch the
    threshold bootstrap off."""
    cdf = O.zipf_cdf(2000)
    doc_off, toks = O.synth_doc_tokens(SEED + 7, cdf, 0, 70000)
    oix = O.BM25(n_terms=2000, doc_off=doc_off, tokens=toks, k1=k1, b=b)
    dev = build(api, ctx, oix, 70000, k1=k1, b=b)
    q_off, q_terms = O.synth_query_terms(SEED + 7, cdf, 0, 24)
    for k in (10, 100):
        compare(dev, oix, q_terms, q_off, k)
    dev.close()


def test_range_boundaries_and_many_queries(api, ctx):
    """Documents exactly at the 32768-document range boundaries and at the 2048-document warp sub-range boundaries; enough
    queries that every query is one work item (no range chunking) and few enough documents that most ranges are sparse."""
    n_docs = 3 * 32768 + 1
    docs = [[] for _ in range(n_docs)]
    for d in (0, 2047, 2048, 32767, 32768, 65535, 65536, 98303, 98304):
        docs[d] = [1, 2, 2, 3]
    for d in range(0, n_docs, 997):
        docs[d] = docs[d] + [3, 4]
    oix = O.BM25(docs, 6)
    dev = build(api, ctx, oix, n_docs)
    rng = np.random.default_rng(5)
    qs = [list(rng.integers(0, 6, int(rng.integers(1, 6)))) for _ in range(400)]
    q_off = np.cumsum([0] + [len(q) for q in qs]).astype(np.uint32)
    q_terms = np.array([x for q in qs for x in q], np.uint32)
    compare(dev, oix, q_terms, q_off, 20)
    dev.close()


def test_more_than_4096_queries_keep_submission_order(api, ctx):
    """Batches above 4096 queries skip the posting-volume sort (identity order); results must not depend on it."""
    cdf = O.zipf_cdf(1500)
    doc_off, toks = O.synth_doc_tokens(SEED + 11, cdf, 0, 9000)
    oix = O.BM25(n_terms=1500, doc_off=doc_off, tokens=toks)
    dev = build(api, ctx, oix, 9000)
    q_off, q_terms = O.synth_query_terms(SEED + 11, cdf, 0, 4500)
    compare(dev, oix, q_terms, q_off, 10)
    dev.close()


def _csr_of(doc_off, toks, n_terms, lo, hi):
    """CSR over term ids of documents [lo, hi) with doc ids relative to lo, plus the documents' lengths."""
    sub_off = (doc_off[lo:hi + 1] - doc_off[lo]).astype(np.uint64)
    sub = O.BM25(n_terms=n_terms, doc_off=sub_off, tokens=toks[int(doc_off[lo]):int(doc_off[hi])])
    term_off, post_doc, post_tf, doc_len, df = sub.csr()
    return term_off, post_doc, post_tf, doc_len, df


def test_append_reweights_on_device_and_equals_full_build(api, ctx):
    """BM25Index::add after the first search (src/index.rs:176-204): trr_bm25_append merges the postings of the new
    documents on the device and re-weights everything with the new N / df / avgdl; impacts, ids and scores must equal
    a from-scratch build over all documents (and therefore the oracle) bit for bit.  The vocabulary grows on the way."""
    V_all, n_all = 3000, 60000
    cdf = O.zipf_cdf(V_all)
    doc_off, toks = O.synth_doc_tokens(SEED + 21, cdf, 0, n_all)
    cuts = [0, 20000, 20001, 45000, n_all]                    # a single-document append included
    vocab = [1200, 1200, 2500, V_all]                         # terms >= vocab[i] do not exist yet in step i
    toks = toks.copy()
    for i in range(len(cuts) - 1):                            # restrict early documents to the early vocabulary
        seg = slice(int(doc_off[cuts[i]]), int(doc_off[cuts[i + 1]]))
        toks[seg] = toks[seg] % vocab[i]
    q_off, q_terms = O.synth_query_terms(SEED + 21, cdf, 0, 48)
    dev = None
    df_tot = np.zeros(V_all, np.uint64)
    len_tot = 0
    for i in range(len(cuts) - 1):
        lo, hi, V = cuts[i], cuts[i + 1], vocab[i]
        t_off, pd, ptf, dl, df = _csr_of(doc_off, toks, V, lo, hi)
        df_tot[:V] += df
        len_tot += int(dl.sum())
        n_now = hi
        avgdl = float(np.float32(np.uint32(len_tot & 0xFFFFFFFF)) / np.float32(n_now))
        idf = api.bm25_idf_host(n_now, df_tot[:V].astype(np.uint32))
        if dev is None:
            dev = api.Bm25Device(ctx, hi - lo, t_off, pd, ptf, dl, avgdl, idf)
        else:
            dev.append(hi - lo, t_off, pd, ptf, dl, avgdl, idf)
        full = O.BM25(n_terms=V, doc_off=doc_off[:hi + 1], tokens=toks[:int(doc_off[hi])])
        assert abs(full.avgdl - avgdl) == 0.0
        qt = np.where(q_terms < V, q_terms, 0xFFFFFFFF).astype(np.uint32)
        compare(dev, full, qt, q_off, 50)
        f_off, f_pd, f_ptf, f_dl, f_df = full.csr()
        ref = build(api, ctx, full, hi)
        assert np.array_equal(dev.impacts(), ref.impacts())
        assert dev.n_postings == ref.n_postings
        ref.close()
    dev.close()


def test_remove_in_place_equals_rebuild_without_the_documents(api, ctx):
    """BM25Index::remove (src/index.rs:245-275) on the device: the removed documents' postings are tombstoned (tf = 0) and
    everything is re-weighted with the new N / df / avgdl; ids and scores must equal an index built without them."""
    V, n = 1500, 40000
    cdf = O.zipf_cdf(V)
    doc_off, toks = O.synth_doc_tokens(SEED + 31, cdf, 0, n)
    full = O.BM25(n_terms=V, doc_off=doc_off, tokens=toks)
    dev = build(api, ctx, full, n, doc_base=100)
    q_off, q_terms = O.synth_query_terms(SEED + 31, cdf, 0, 32)
    top = dev.search(q_terms, q_off, 5)[0]
    rng = np.random.default_rng(4)
    gone = np.unique(np.concatenate([top[:, :2].ravel() - 100, rng.integers(0, n, 3000).astype(np.uint32)]))
    keep = np.setdiff1d(np.arange(n, dtype=np.uint32), gone)
    # the oracle without the removed documents (ordinals renumbered; `keep` maps them back, order preserved)
    lens = np.diff(doc_off).astype(np.int64)
    k_off = np.zeros(len(keep) + 1, np.uint64)
    np.cumsum(lens[keep], out=k_off[1:])
    k_toks = np.concatenate([toks[int(doc_off[d]):int(doc_off[d + 1])] for d in keep])
    kept = O.BM25(n_terms=V, doc_off=k_off, tokens=k_toks)
    _, _, _, _, df_kept = kept.csr()
    dead = dev.remove(gone, kept.avgdl, api.bm25_idf_host(len(keep), df_kept))
    assert 0 < dead < dev.n_postings
    for k in (10, 100):
        ords, scores, cnt = dev.search(q_terms, q_off, k)
        eo, es, en = kept.search_batch(q_terms, q_off, k)
        assert np.array_equal(cnt, en)
        for b in range(len(q_off) - 1):
            m = int(en[b])
            assert np.array_equal(ords[b, :m], keep[eo[b, :m]] + 100), b
            assert np.array_equal(scores[b, :m], es[b, :m]), b
    dev.close()

