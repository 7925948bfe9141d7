"""The reference's known-answer tests replayed through the drop-in surface on the GPU:
VectorStore / BM25Index / FusionStrategy / HybridRetriever (C++ host mirror -> C ABI -> CUDA kernels), each checked
against the oracle bit for bit.  Test names follow the reference's."""
import math

import numpy as np
import pytest

from oracle import oracle as O
from oracle.tokenizer import TextIndex

pytestmark = pytest.mark.gpu
F32 = np.float32


@pytest.fixture(scope="module")
def api(ctx):
    from trueno_rag_b200 import api as a
    return a


def chunk(api, content, emb=None):
    return api.Chunk(content, embedding=emb)


# ------------------------------------------------------------------ src/index.rs:742-867
def test_vector_store_search_cosine(api):
    store = api.VectorStore.with_dimension(3)
    c1 = chunk(api, "north", [1.0, 0.0, 0.0])
    c2 = chunk(api, "east", [0.0, 1.0, 0.0])
    c3 = chunk(api, "diagonal", [math.sqrt(0.5), math.sqrt(0.5), 0.0])
    for c in (c1, c2, c3):
        store.insert(c)
    res = store.search([0.9, 0.1, 0.0], 10)
    assert len(res) == 3
    assert res[0][0] == c1.id and res[1][0] == c3.id and res[2][0] == c2.id
    rows = np.array([c1.embedding, c2.embedding, c3.embedding], F32)
    _, exp = O.dense_search(rows, [0.9, 0.1, 0.0], 10)
    assert [r[1] for r in res] == [float(x) for x in exp]           # bit-exact scores


def test_vector_store_search_top_k(api):
    store = api.VectorStore.with_dimension(3)
    cs = [chunk(api, f"chunk {i}", [float(i), 0.0, 0.0]) for i in range(10)]
    store.insert_batch(cs)
    res = store.search([9.0, 0.0, 0.0], 3)
    assert len(res) == 3
    assert [r[0] for r in res] == [cs[1].id, cs[2].id, cs[3].id]    # nine exact ties -> insertion order
    assert all(r[1] == 1.0 for r in res)


def test_vector_store_search_wrong_dimension(api):
    store = api.VectorStore.with_dimension(3)
    with pytest.raises(api.Error) as e:
        store.search([1.0, 0.0], 10)
    assert e.value.kind == "DimensionMismatch" and e.value.expected == 3 and e.value.actual == 2


def test_vector_store_remove(api):
    store = api.VectorStore.with_dimension(3)
    c = chunk(api, "test", [1.0, 0.0, 0.0])
    keep = chunk(api, "keep", [0.0, 1.0, 0.0])
    store.insert(c)
    store.insert(keep)
    assert len(store) == 2
    assert store.remove(c.id) is True
    assert len(store) == 1 and store.get(c.id) is None
    res = store.search([1.0, 0.0, 0.0], 10)
    assert [r[0] for r in res] == [keep.id]


def test_vector_store_reinsert_replaces_vector(api):               # HashMap::insert semantics (src/index.rs:372)
    store = api.VectorStore.with_dimension(2)
    c = chunk(api, "a", [1.0, 0.0])
    store.insert(c)
    store.insert(api.Chunk("a2", id=c.id, embedding=[0.0, 1.0]))
    assert len(store) == 1
    res = store.search([0.0, 1.0], 10)
    assert len(res) == 1 and res[0][0] == c.id and res[0][1] == 1.0


def test_distance_metric_euclidean(api):
    store = api.VectorStore(2, api.EUCLIDEAN)
    c1, c2, c3 = chunk(api, "origin", [0.0, 0.0]), chunk(api, "near", [1.0, 0.0]), chunk(api, "far", [10.0, 0.0])
    store.insert_batch([c1, c2, c3])
    res = store.search([0.0, 0.0], 10)
    assert [r[0] for r in res] == [c1.id, c2.id, c3.id]
    assert [r[1] for r in res] == [0.0, -1.0, -10.0]


def test_distance_metric_dot_product(api):
    store = api.VectorStore(2, api.DOT)
    c1, c2 = chunk(api, "small", [1.0, 0.0]), chunk(api, "large", [10.0, 0.0])
    store.insert_batch([c1, c2])
    res = store.search([1.0, 0.0], 10)
    assert res[0][0] == c2.id and res[0][1] == 10.0


def test_prop_vector_store_search_returns_stored(api):              # src/index.rs:913-935
    rng = np.random.default_rng(0)
    for _ in range(8):
        dim, n = int(rng.integers(2, 10)), int(rng.integers(1, 20))
        store = api.VectorStore.with_dimension(dim)
        ids = []
        for i in range(n):
            e = [0.0] * dim
            e[i % dim] = 1.0
            c = chunk(api, f"chunk {i}", e)
            ids.append(c.id)
            store.insert(c)
        res = store.search([1.0] * dim, n)
        assert len(res) == n and all(r[0] in ids for r in res)
        assert [r[0] for r in res] == ids                            # all tie -> insertion order


def test_vector_store_clone_is_independent(api):
    store = api.VectorStore.with_dimension(2)
    a, b = chunk(api, "a", [1.0, 0.0]), chunk(api, "b", [0.0, 1.0])
    store.insert_batch([a, b])
    twin = store.clone()
    store.remove(a.id)
    assert len(twin) == 2 and len(store) == 1
    assert [r[0] for r in twin.search([1.0, 0.0], 2)] == [a.id, b.id]


# ------------------------------------------------------------------ src/index.rs:547-670 (BM25)
def bm25(api, texts):
    ix = api.BM25Index()
    cs = [chunk(api, t) for t in texts]
    ix.add_batch(cs)
    ti = TextIndex()
    for t in texts:
        ti.add(t)
    return ix, cs, ti


def assert_same(api_res, cs, oracle_res):
    o, s = oracle_res
    assert [r[0] for r in api_res] == [cs[i].id for i in o]
    assert [r[1] for r in api_res] == [float(x) for x in s]


def test_bm25_search_basic(api):
    ix, cs, ti = bm25(api, ["Machine learning algorithms", "Deep learning neural networks", "Natural language processing"])
    res = ix.search("machine learning", 10)
    assert res and any(r[0] == cs[0].id for r in res)
    assert_same(res, cs, ti.build().search(ti.query_ids("machine learning"), 10))


def test_bm25_search_empty_query_stopwords_no_match(api):
    ix, cs, _ = bm25(api, ["Test document"])
    assert ix.search("", 10) == []
    assert ix.search("the a an", 10) == []
    ix, cs, _ = bm25(api, ["Cats and dogs"])
    assert ix.search("quantum physics", 10) == []


def test_bm25_search_ranking(api):
    ix, cs, ti = bm25(api, ["python programming language", "python python python programming"])
    res = ix.search("python programming", 10)
    assert len(res) == 2 and res[0][0] == cs[1].id
    assert_same(res, cs, ti.build().search(ti.query_ids("python programming"), 10))


def test_bm25_search_top_k(api):
    ix, cs, ti = bm25(api, [f"document {i} about rust" for i in range(10)])
    res = ix.search("rust", 3)
    assert len(res) == 3
    assert_same(res, cs, ti.build().search(ti.query_ids("rust"), 3))   # ten exact ties -> insertion order


def test_bm25_remove(api):
    ix = api.BM25Index()
    c = chunk(api, "Test document")
    ix.add(c)
    assert len(ix) == 1
    ix.remove(c.id)
    assert len(ix) == 0 and ix.search("test", 10) == []


def test_bm25_remove_updates_statistics(api):
    texts = ["rust memory safety", "python data science", "rust compiler borrow rules", "go concurrency"]
    ix, cs, _ = bm25(api, texts)
    ix.remove(cs[1].id)
    ti = TextIndex()
    for i, t in enumerate(texts):
        if i != 1:
            ti.add(t)
    o, s = ti.build().search(ti.query_ids("rust compiler"), 10)
    kept = [cs[0], cs[2], cs[3]]
    res = ix.search("rust compiler", 10)
    assert [r[0] for r in res] == [kept[i].id for i in o]
    assert [r[1] for r in res] == [float(x) for x in s]


def test_bm25_incremental_adds_between_searches(api):
    """RagPipeline::index_document interleaves with queries (src/pipeline.rs:333-347): every add after a search goes
    through trr_bm25_append (device-side merge + re-weighting) and must equal the oracle rebuilt from scratch, including
    new vocabulary, a removal in the middle (full rebuild) and adds after it."""
    words = "rust memory safety python data science compiler borrow rules go concurrency tensor kernel index search".split()
    rng = np.random.default_rng(9)
    ix, ti, cs = api.BM25Index(), TextIndex(), []
    removed = set()
    for step in range(60):
        text = " ".join(rng.choice(words[: 4 + step // 4], int(rng.integers(2, 9))))
        c = chunk(api, text)
        ix.add(c)
        cs.append(c)
        if step == 30:
            ix.remove(cs[7].id)
            removed.add(7)
        if step % 3 == 0 or step in (30, 31, 32):
            ti2, kept = TextIndex(), []
            for i, cc in enumerate(cs):
                if i not in removed:
                    ti2.add(cc.content)
                    kept.append(cc)
            for q in ("rust compiler", "python data data", "tensor kernel search index"):
                o, sc = ti2.build().search(ti2.query_ids(q), 10)
                res = ix.search(q, 10)
                assert [r[0] for r in res] == [kept[i].id for i in o], (step, q)
                assert [r[1] for r in res] == [float(x) for x in sc], (step, q)


def test_bm25_avg_doc_length_and_idf(api):
    ix, cs, ti = bm25(api, ["short text", "this is a longer piece of text about programming"])
    assert ix.avg_doc_length == ti.build().avgdl > 0
    ix, cs, ti = bm25(api, ["common rare", "common word", "common term"])
    rare, common = ix.search("rare", 10), ix.search("common", 10)
    assert len(rare) == 1 and len(common) == 3 and rare[0][1] > common[0][1]


def test_prop_bm25_scores_non_negative_and_within_k(api):
    rng = np.random.default_rng(1)
    words = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta", "iota", "kappa"]
    for _ in range(6):
        texts = [" ".join(rng.choice(words, int(rng.integers(1, 9)))) for _ in range(int(rng.integers(3, 12)))]
        ix, cs, ti = bm25(api, texts)
        q = " ".join(rng.choice(words + ["omega"], 3))
        for k in (1, 4, 100):
            res = ix.search(q, k)
            assert len(res) <= k and all(r[1] >= 0 for r in res)
            assert_same(res, cs, ti.build().search(ti.query_ids(q), k))


# ------------------------------------------------------------------ src/fusion.rs:273-540
def cid(api, n):
    return api.ChunkId.from_u128(n)


def lst(api, pairs):
    return [(cid(api, i), s) for i, s in pairs]


def test_rrf(api):
    f = api.FusionStrategy.RRF(60.0)
    assert f.fuse([], []) == []
    r = f.fuse(lst(api, [(1, 0.9), (2, 0.8)]), [])
    assert [x[0] for x in r] == [cid(api, 1), cid(api, 2)]
    assert len(f.fuse([], lst(api, [(1, 0.9), (2, 0.8)]))) == 2
    r = f.fuse(lst(api, [(1, 0.9), (2, 0.8)]), lst(api, [(1, 0.9), (3, 0.8)]))
    assert len(r) == 3 and r[0][0] == cid(api, 1)
    r = f.fuse(lst(api, [(1, 1.0)]), lst(api, [(1, 1.0)]))
    assert abs(r[0][1] - 2.0 / 61.0) < 1e-3
    assert r[0][1] == float(O.fuse(O.RRF, 60.0, ([1], [1.0]), ([1], [1.0]))[1][0])


def test_linear_and_convex(api):
    assert api.FusionStrategy.Linear(0.5).fuse([], []) == []
    assert len(api.FusionStrategy.Linear(0.7).fuse(lst(api, [(1, 1.0), (2, 0.5)]), [])) > 0
    r = api.FusionStrategy.Linear(0.5).fuse(lst(api, [(1, 1.0)]), lst(api, [(1, 1.0)]))
    assert abs(r[0][1] - 1.0) < 0.01
    r = api.FusionStrategy.Linear(0.9).fuse(lst(api, [(1, 1.0), (2, 0.0)]), lst(api, [(2, 1.0), (1, 0.0)]))
    assert r[0][0] == cid(api, 1)
    a = api.FusionStrategy.Linear(0.6).fuse(lst(api, [(1, 0.9), (2, 0.5)]), lst(api, [(2, 0.8), (3, 0.4)]))
    b = api.FusionStrategy.Convex(0.6).fuse(lst(api, [(1, 0.9), (2, 0.5)]), lst(api, [(2, 0.8), (3, 0.4)]))
    assert a == b
    for w in np.linspace(0.0, 0.99, 7):
        r = api.FusionStrategy.Linear(float(w)).fuse(lst(api, [(1, 1.0)]), lst(api, [(1, 1.0)]))
        assert abs(r[0][1] - 1.0) < 0.01


def test_dbsf_union_intersection(api):
    assert api.FusionStrategy.DBSF().fuse([], []) == []
    r = api.FusionStrategy.DBSF().fuse(lst(api, [(1, 10.0), (2, 5.0), (3, 0.0)]), lst(api, [(1, 100.0), (2, 50.0), (3, 0.0)]))
    assert r[0][0] == cid(api, 1)
    u = api.FusionStrategy.Union()
    assert len(u.fuse(lst(api, [(1, 0.9)]), lst(api, [(2, 0.8)]))) == 2
    assert len(u.fuse(lst(api, [(1, 0.9), (2, 0.8)]), lst(api, [(1, 0.7), (3, 0.6)]))) == 3
    r = u.fuse(lst(api, [(1, 0.9)]), lst(api, [(1, 0.5)]))
    assert abs(r[0][1] - 0.9) < 1.2e-7
    i = api.FusionStrategy.Intersection()
    assert i.fuse(lst(api, [(1, 0.9)]), lst(api, [(2, 0.8)])) == []
    r = i.fuse(lst(api, [(1, 0.8), (2, 0.6)]), lst(api, [(2, 0.9), (3, 0.5)]))
    assert len(r) == 1 and r[0][0] == cid(api, 2)
    r = i.fuse(lst(api, [(1, 0.8)]), lst(api, [(1, 0.4)]))
    assert abs(r[0][1] - 0.6) < 1e-3


# ------------------------------------------------------------------ src/retrieve.rs:464-531,669-718 + examples/hybrid_search.rs
DOCS = [
    "Rust programming language provides memory safety guarantees without garbage collection.",
    "Python is excellent for data science and machine learning applications.",
    "Go provides fast compilation and built-in concurrency primitives.",
    "Memory management in systems programming is crucial for performance.",
    "The Rust compiler enforces strict borrowing rules at compile time.",
]
QUERY = "memory safe programming language"


def fake_embed(text, dim=384):
    """Deterministic stand-in for MockEmbedder (embedders are out of scope: embeddings are inputs to the path)."""
    h = abs(hash(text)) % (2 ** 31)
    return O.synth_corpus(0xE0BED, h, 1, dim)[0][0]


def embed_stable(text, dim=384):
    import zlib
    return O.synth_corpus(0xE0BED, zlib.crc32(text.encode()), 1, dim)[0][0]


def build_retriever(api, fusion, C_=50, use_dense=True, use_sparse=True, dim=384):
    r = api.HybridRetriever(api.VectorStore.with_dimension(dim), api.BM25Index(), lambda t: embed_stable(t, dim),
                            api.HybridRetrieverConfig(C_, fusion, use_dense, use_sparse))
    cs = [api.Chunk(t, embedding=embed_stable(t, dim)) for t in DOCS]
    r.index_batch(cs)
    return r, cs


def oracle_hybrid(strategy, param, k, C_=50, use_dense=True, use_sparse=True, dim=384):
    rows = np.array([embed_stable(t, dim) for t in DOCS], F32)
    ti = TextIndex()
    for t in DOCS:
        ti.add(t)
    d = O.dense_search(rows, embed_stable(QUERY, dim), C_) if use_dense else ([], [])
    s = ti.build().search(ti.query_ids(QUERY), C_) if use_sparse else ([], [])
    return O.hybrid_assemble(strategy, param, d, s, k)


@pytest.mark.parametrize("name,fusion,strategy,param", [
    ("rrf", lambda a: a.FusionStrategy.RRF(60.0), O.RRF, 60.0),
    ("linear0.7", lambda a: a.FusionStrategy.Linear(0.7), O.LINEAR, 0.7),
    ("linear0.3", lambda a: a.FusionStrategy.Linear(0.3), O.LINEAR, 0.3),
    ("dbsf", lambda a: a.FusionStrategy.DBSF(), O.DBSF, 0.0),
    ("union", lambda a: a.FusionStrategy.Union(), O.UNION, 0.0),
    ("intersection", lambda a: a.FusionStrategy.Intersection(), O.INTERSECTION, 0.0),
])
def test_hybrid_search_example_all_strategies(api, name, fusion, strategy, param):
    r, cs = build_retriever(api, fusion(api))
    assert len(r) == 5
    for k in (3, 6, 20):                                             # pipeline.query(q, 3) -> retrieve(q, 6)
        res = r.retrieve(QUERY, k)
        i, f, d, s = oracle_hybrid(strategy, np.float32(param), k)
        assert len(res) <= k
        assert [x.chunk_id for x in res] == [cs[j].id for j in i]
        assert [x.fused_score for x in res] == [float(v) for v in f]
        assert [x.dense_score for x in res] == [None if math.isnan(v) else float(v) for v in d]
        assert [x.sparse_score for x in res] == [None if math.isnan(v) else float(v) for v in s]


def test_retrieve_dense_and_sparse_only(api):
    r, cs = build_retriever(api, api.FusionStrategy.RRF(60.0))
    rd = r.retrieve_dense(QUERY, 3)
    assert len(rd) == 3 and all(x.dense_score is not None and x.sparse_score is None and x.fused_score is None for x in rd)
    rs = r.retrieve_sparse(QUERY, 3)
    assert rs and all(x.sparse_score is not None and x.dense_score is None for x in rs)
    r2, _ = build_retriever(api, api.FusionStrategy.RRF(60.0), use_sparse=False)
    res = r2.retrieve(QUERY, 3)
    assert res and all(x.dense_score is not None and x.sparse_score is None for x in res)
    r3, _ = build_retriever(api, api.FusionStrategy.RRF(60.0), use_dense=False)
    res = r3.retrieve(QUERY, 3)
    assert res and all(x.sparse_score is not None and x.dense_score is None for x in res)
    i, f, d, s = oracle_hybrid(O.RRF, 60.0, 3, use_dense=False)
    assert [x.fused_score for x in res] == [float(v) for v in f]


def test_hybrid_retriever_respects_k(api):                          # prop_hybrid_retriever_respects_k
    r, _ = build_retriever(api, api.FusionStrategy.RRF(60.0))
    for k in (1, 2, 5, 9):
        assert len(r.retrieve(QUERY, k)) <= k


def test_hybrid_index_requires_embedding(api):
    r = api.HybridRetriever(api.VectorStore.with_dimension(4), api.BM25Index(), lambda t: [0.0] * 4)
    with pytest.raises(api.Error):
        r.index(api.Chunk("no embedding"))


def test_nemotron_example_ranking(api, ctx):
    """examples/nemotron_embeddings.rs:60-92: a handful of 4096-dimensional document embeddings ranked against one query by
    cosine similarity (stable sort, descending) - through the exact scan kernel, bit for bit the oracle's order and scores."""
    rng = np.random.default_rng(79)
    docs = rng.standard_normal((6, 4096)).astype(np.float32)
    docs[4] = docs[1]                                                    # an exact tie keeps document order
    docs /= np.linalg.norm(docs, axis=1, keepdims=True).astype(np.float32)
    q = (docs[2] + 0.5 * rng.standard_normal(4096)).astype(np.float32)
    idx, sims = api.rank_by_cosine(ctx, q, docs)
    eo, es, en = O.dense_search_batch(docs, q[None, :], 6)
    assert int(en[0]) == 6 and np.array_equal(idx, eo[0]) and np.array_equal(sims, es[0])
    assert idx[0] == 2 and list(idx).index(1) < list(idx).index(4)
