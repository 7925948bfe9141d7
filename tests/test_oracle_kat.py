"""Pins the CPU oracle against the reference's own known-answer tests and against an
independent numpy-float32 restatement.  CPU only."""
import math

import numpy as np
import pytest

from oracle import oracle as O
from oracle.tokenizer import TextIndex, tokenize

F32 = np.float32


# ---------------------------------------------------------------- numpy cross-check of the C oracle
def np_seq_sum(x):
    s = F32(0.0)
    for v in x.astype(F32):
        s = F32(s + v)
    return s


def np_cosine(a, b):
    a, b = a.astype(F32), b.astype(F32)
    dot = np_seq_sum(a * b)
    na, nb = np.sqrt(np_seq_sum(a * a)), np.sqrt(np_seq_sum(b * b))
    if na == 0 or nb == 0:
        return F32(0.0)
    return F32(dot / F32(na * nb))


def test_distance_functions_match_numpy_f32():
    rng = np.random.default_rng(1)
    for d in (1, 2, 3, 7, 64, 384):
        a, b = rng.standard_normal(d).astype(F32), rng.standard_normal(d).astype(F32)
        assert O.cosine(a, b) == float(np_cosine(a, b))
        assert O.dot(a, b) == float(np_seq_sum(a * b))
        diff = (a - b).astype(F32)
        assert O.euclidean(a, b) == float(np.sqrt(np_seq_sum(diff * diff)))


# ---------------------------------------------------------------- reference src/embed.rs:1339-1409 (public copies)
def test_cosine_similarity_kats():
    assert abs(O.cosine([1, 2, 3], [1, 2, 3]) - 1.0) < 1e-5       # identical
    assert abs(O.cosine([1, 0], [0, 1])) < 1e-5                    # orthogonal
    assert abs(O.cosine([1, 0], [-1, 0]) + 1.0) < 1e-5             # opposite
    assert O.cosine([0, 0, 0], [1, 2, 3]) == 0.0                   # zero vector
    assert O.cosine([1, 2], [1, 2, 3]) == 0.0                      # length mismatch (src/embed.rs:313-315)
    assert O.dot([1, 2, 3], [4, 5, 6]) == 32.0
    assert O.euclidean([0, 0], [3, 4]) == 5.0


def test_cosine_bounded_property():  # tests/property_tests.rs:87-96
    rng = np.random.default_rng(2)
    for _ in range(200):
        d = int(rng.integers(1, 64))
        a, b = rng.uniform(-10, 10, d).astype(F32), rng.uniform(-10, 10, d).astype(F32)
        assert -1.0 - 1e-5 <= O.cosine(a, b) <= 1.0 + 1e-5


# ---------------------------------------------------------------- reference src/index.rs:742-867 (VectorStore)
def test_vector_store_search_cosine():
    rows = np.array([[1, 0, 0], [0, 1, 0], [math.sqrt(0.5), math.sqrt(0.5), 0]], F32)
    ords, scores = O.dense_search(rows, [0.9, 0.1, 0.0], 10)
    assert list(ords) == [0, 2, 1]                                  # north, diagonal, east
    assert [float(s) for s in scores] == [float(F32(0.9938837)), float(F32(0.78086877)), float(F32(0.11043153))]


def test_vector_store_search_top_k_and_ties():
    rows = np.array([[i, 0, 0] for i in range(10)], F32)
    ords, scores = O.dense_search(rows, [9, 0, 0], 3)
    assert len(ords) == 3
    assert list(ords) == [1, 2, 3] and all(s == 1.0 for s in scores)  # nine exact ties -> ordinal order; row 0 scores 0.0


def test_distance_metric_euclidean_and_dot():
    rows = np.array([[0, 0], [1, 0], [10, 0]], F32)
    ords, scores = O.dense_search(rows, [0, 0], 10, metric=O.EUCLIDEAN)
    assert list(ords) == [0, 1, 2] and list(scores) == [0.0, -1.0, -10.0]
    ords, _ = O.dense_search(np.array([[1, 0], [10, 0]], F32), [1, 0], 10, metric=O.DOT)
    assert ords[0] == 1


def test_dense_literal_equals_bounded_selection():
    rng = np.random.default_rng(3)
    rows = rng.standard_normal((500, 16)).astype(F32)
    rows[100] = rows[7]
    rows[300] = rows[7]                                             # exact ties
    q = rng.standard_normal(16).astype(F32)
    for k in (1, 5, 50, 500, 600):
        a = O.dense_search(rows, q, k, literal=True)
        b = O.dense_search(rows, q, k, literal=False)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    c = O.dense_search_par(rows, q, 50, threads=4)
    a = O.dense_search(rows, q, 50)
    assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1])


# ---------------------------------------------------------------- reference src/index.rs:498-634 (BM25)
def test_bm25_tokenize():
    toks = tokenize("Hello World! This is a test.")
    assert "hello" in toks and "world" in toks and "test" in toks
    assert "this" not in toks and "is" not in toks and "a" not in toks
    assert tokenize("HELLO World") == ["hello", "world"]
    assert tokenize("x y z I") == []                                # byte length < 2


def _index(texts, k1=1.2, b=0.75):
    ti = TextIndex(k1, b)
    for t in texts:
        ti.add(t)
    return ti, ti.build()


def test_bm25_search_ranking_kat():
    ti, ix = _index(["python programming language", "python python python programming"])
    ords, scores = ix.search(ti.query_ids("python programming"), 10)
    assert list(ords) == [1, 0]
    assert abs(O.bm25_idf(2, 2) - 0.18232156) < 1e-7                    # ln((2 - 2 + 0.5) / (2 + 0.5) + 1), src/index.rs:147
    assert abs(scores[0] - 0.45025012) < 1e-6 and abs(scores[1] - 0.38727623) < 1e-6
    # literal form gives the same bits
    lo, ls = ix.search(ti.query_ids("python programming"), 10, literal=True)
    assert np.array_equal(lo, ords) and np.array_equal(ls, scores)


def test_bm25_search_edge_cases():
    ti, ix = _index(["Test document"])
    assert len(ix.search(ti.query_ids(""), 10)[0]) == 0                    # empty query
    assert len(ix.search(ti.query_ids("the a an"), 10)[0]) == 0            # stopwords only
    ti, ix = _index(["Cats and dogs"])
    assert len(ix.search(ti.query_ids("quantum physics"), 10)[0]) == 0     # no match
    ti, ix = _index([f"document {i} about rust" for i in range(10)])
    assert len(ix.search(ti.query_ids("rust"), 3)[0]) == 3                 # top-k cap
    ti, ix = _index(["Machine learning algorithms", "Deep learning neural networks", "Natural language processing"])
    ords, _ = ix.search(ti.query_ids("machine learning"), 10)
    assert 0 in ords and ords[0] == 0


def test_bm25_idf_rare_vs_common():
    ti, ix = _index(["common rare", "common word", "common term"])
    rare = ix.search(ti.query_ids("rare"), 10)
    common = ix.search(ti.query_ids("common"), 10)
    assert len(rare[0]) == 1 and len(common[0]) == 3 and rare[1][0] > common[1][0]


def test_bm25_duplicate_query_terms_count_twice():
    ti, ix = _index(["rust systems", "python scripts", "rust rust"])
    one = ix.search(ti.query_ids("rust"), 10)
    two = ix.search(ti.query_ids("rust rust"), 10)
    assert np.array_equal(one[0], two[0])
    assert np.array_equal(two[1], (one[1] + one[1]).astype(F32))


def test_bm25_scores_non_negative_and_literal_equals_fast():
    rng = np.random.default_rng(4)
    for trial in range(20):
        n_docs, n_terms = int(rng.integers(3, 60)), int(rng.integers(2, 30))
        docs = [list(rng.integers(0, n_terms, int(rng.integers(0, 12)))) for _ in range(n_docs)]
        ix = O.BM25(docs, n_terms)
        q = list(rng.integers(0, n_terms + 2, int(rng.integers(1, 8))))   # includes unknown ids
        q = [x if x < n_terms else 0xFFFFFFFF for x in q]
        for k in (1, 5, 100):
            a = ix.search(q, k, literal=True)
            b = ix.search(q, k, literal=False)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
            assert (a[1] >= 0).all() and len(a[0]) <= k


def test_bm25_avgdl_is_u32_sum_over_count():
    ti, ix = _index(["short text", "this is a longer piece of text about programming"])
    term_off, post_doc, post_tf, doc_len, df = ix.csr()
    assert list(doc_len) == [2, 5]
    assert ix.avgdl == float(F32(7) / F32(2))


# ---------------------------------------------------------------- reference src/fusion.rs:273-540
def ids(*xs):
    return list(xs)


def test_rrf_kats():
    assert len(O.fuse(O.RRF, 60.0, ([], []), ([], []))[0]) == 0
    i, s = O.fuse(O.RRF, 60.0, ([1, 2], [0.9, 0.8]), ([], []))
    assert list(i) == [1, 2]
    i, s = O.fuse(O.RRF, 60.0, ([], []), ([1, 2], [0.9, 0.8]))
    assert len(i) == 2
    i, s = O.fuse(O.RRF, 60.0, ([1, 2], [0.9, 0.8]), ([1, 3], [0.9, 0.8]))
    assert len(i) == 3 and i[0] == 1
    assert list(i[1:]) == [2, 3] and s[1] == s[2]                    # exact RRF tie -> ordinal order
    i, s = O.fuse(O.RRF, 60.0, ([1], [1.0]), ([1], [1.0]))
    assert abs(s[0] - 2.0 / 61.0) < 1e-3 and s[0] == F32(F32(1.0) / F32(61.0)) + F32(F32(1.0) / F32(61.0))


def test_linear_convex_kats():
    assert len(O.fuse(O.LINEAR, 0.5, ([], []), ([], []))[0]) == 0
    assert len(O.fuse(O.LINEAR, 0.7, ([1, 2], [1.0, 0.5]), ([], []))[0]) == 2
    i, s = O.fuse(O.LINEAR, 0.5, ([1], [1.0]), ([1], [1.0]))
    assert abs(s[0] - 1.0) < 0.01
    i, s = O.fuse(O.LINEAR, 0.9, ([1, 2], [1.0, 0.0]), ([2, 1], [1.0, 0.0]))
    assert i[0] == 1
    a = O.fuse(O.LINEAR, 0.6, ([1, 2], [0.9, 0.5]), ([2, 3], [0.8, 0.4]))
    b = O.fuse(O.CONVEX, 0.6, ([1, 2], [0.9, 0.5]), ([2, 3], [0.8, 0.4]))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    for w in np.linspace(0, 0.99, 12):                               # prop_linear_weights_sum_to_one
        _, s = O.fuse(O.LINEAR, float(w), ([1], [1.0]), ([1], [1.0]))
        assert abs(s[0] - 1.0) < 0.01


def test_normalizer_kats():
    assert O.min_max([]).size == 0
    assert abs(O.min_max([5.0])[0] - 1.0) < 1e-3
    assert list(O.min_max([10, 5, 0])) == [1.0, 0.5, 0.0]
    assert O.z_score([]).size == 0
    assert all(abs(v) < 1e-3 for v in O.z_score([5, 5, 5]))


def test_dbsf_union_intersection_kats():
    assert len(O.fuse(O.DBSF, 0, ([], []), ([], []))[0]) == 0
    i, _ = O.fuse(O.DBSF, 0, ([1, 2, 3], [10, 5, 0]), ([1, 2, 3], [100, 50, 0]))
    assert i[0] == 1
    assert len(O.fuse(O.UNION, 0, ([1], [0.9]), ([2], [0.8]))[0]) == 2
    i, s = O.fuse(O.UNION, 0, ([1, 2], [0.9, 0.8]), ([1, 3], [0.7, 0.6]))
    assert list(i) == [1, 2, 3] and list(s) == [F32(0.9), F32(0.8), F32(0.6)]
    i, s = O.fuse(O.UNION, 0, ([1], [0.9]), ([1], [0.5]))
    assert abs(s[0] - 0.9) < 1.2e-7
    assert len(O.fuse(O.INTERSECTION, 0, ([1], [0.9]), ([2], [0.8]))[0]) == 0
    i, _ = O.fuse(O.INTERSECTION, 0, ([1, 2], [0.8, 0.6]), ([2, 3], [0.9, 0.5]))
    assert list(i) == [2]
    _, s = O.fuse(O.INTERSECTION, 0, ([1], [0.8]), ([1], [0.4]))
    assert abs(s[0] - 0.6) < 1e-3


def test_fusion_properties():
    rng = np.random.default_rng(5)
    for _ in range(50):
        nd, ns = int(rng.integers(1, 10)), int(rng.integers(1, 10))
        d = (list(range(nd)), [1.0 - i * 0.1 for i in range(nd)])
        s = (list(range(100, 100 + ns)), [1.0 - i * 0.1 for i in range(ns)])
        _, sc = O.fuse(O.RRF, 60.0, d, s)
        assert (sc > 0).all()                                       # prop_rrf_scores_positive
        di = list(rng.integers(0, 100, nd))
        si = list(rng.integers(0, 100, ns))
        i, _ = O.fuse(O.INTERSECTION, 0, (di, [1.0] * nd), (si, [1.0] * ns))
        assert set(i) <= (set(di) & set(si)) and set(i) == (set(di) & set(si))   # prop_intersection_subset_of_inputs
        a = O.fuse(O.RRF, 60.0, d, s)
        b = O.fuse(O.RRF, 60.0, d, s)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])         # prop_fusion_deterministic


# ---------------------------------------------------------------- reference src/retrieve.rs:175-220
def test_hybrid_assemble():
    i, f, d, s = O.hybrid_assemble(O.RRF, 60.0, ([1, 2], [0.9, 0.8]), ([1, 3], [0.7, 0.6]), 10)
    assert list(i) == [1, 2, 3]
    assert d[0] == F32(0.9) and s[0] == F32(0.7)
    assert d[1] == F32(0.8) and math.isnan(s[1])
    assert math.isnan(d[2]) and s[2] == F32(0.6)
    i, *_ = O.hybrid_assemble(O.RRF, 60.0, ([1, 2], [0.9, 0.8]), ([1, 3], [0.7, 0.6]), 2)
    assert len(i) == 2                                             # respects k
    i, f, d, s = O.hybrid_assemble(O.RRF, 60.0, ([1, 2], [0.9, 0.8]), ([], []), 10)
    assert all(math.isnan(x) for x in s) and not any(math.isnan(x) for x in d)   # dense only


# ---------------------------------------------------------------- synthetic generators
def test_synth_is_deterministic_and_normalised():
    a, ab = O.synth_corpus(0x5EED0004, 10, 8, 64, bf16=True)
    b, bb = O.synth_corpus(0x5EED0004, 12, 2, 64, bf16=True)
    assert np.array_equal(a[2:4], b) and np.array_equal(ab[2:4], bb)
    f, _ = O.synth_corpus(0x5EED0002, 0, 16, 384)
    assert np.allclose(np.linalg.norm(f, axis=1), 1.0, atol=1e-5)
    assert np.array_equal((ab.astype(np.uint32) << 16).view(F32), a)
    q = O.synth_queries(0x5EED0004, 0, 300, 64, 1000, corpus_bf16=True)
    assert np.allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-5)
    cdf = O.zipf_cdf(5000)
    off, toks = O.synth_doc_tokens(7, cdf, 0, 100)
    lens = np.diff(off)
    assert lens.min() >= 32 and lens.max() <= 64 and toks.max() < 5000
    qoff, qt = O.synth_query_terms(7, cdf, 0, 50)
    ql = np.diff(qoff)
    assert ql.min() >= 8 and ql.max() <= 32
