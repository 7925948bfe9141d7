"""Dense path parity: CUDA (through the C ABI) vs the oracle on the same seeded inputs.
IDs bit-exact under the canonical (score desc, ordinal asc) order; scores bit-exact (the CUDA path reproduces the
reference's f32 operation order: tolerance 0, stricter than north_star's 1e-5)."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
F32 = np.float32
SEED = 0x5EED0002


@pytest.fixture(scope="module")
def api(ctx):
    from trueno_rag_b200 import api as a
    return a


def check(api, ctx, rows, Q, k, metric=0, dtype=0, mode=1, alive=None, base=0, expect_mode=None):
    ix = api.DenseIndex(ctx, rows.shape[1], metric, dtype, base=base)
    try:
        ix.append(rows)
        if alive is not None:
            for i in np.nonzero(~alive)[0]:
                ix.remove(int(i))
        ix.set_mode(mode)
        ords, scores, n = ix.search(Q, k)
        st = ix.stats()
        if expect_mode is not None:
            assert st.mode_used == expect_mode
        eo, es, en = O.dense_search_batch(rows, Q, k, metric=metric, alive=alive, literal=False)
        assert np.array_equal(n, en)
        for b in range(Q.shape[0]):
            m = int(n[b])
            assert np.array_equal(ords[b, :m], eo[b, :m] + base), (b, ords[b, :m], eo[b, :m])
            assert np.array_equal(scores[b, :m], es[b, :m]), (b, scores[b, :m] - es[b, :m])
        return st
    finally:
        ix.close()


def bf16_round(x):
    u = np.ascontiguousarray(x, F32).view(np.uint32)
    return ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32).view(F32)


# ------------------------------------------------------------------ generic kernel (tiny stores, odd dimensions)
@pytest.mark.parametrize("n,d", [(1, 1), (3, 3), (10, 2), (33, 5), (257, 7), (1000, 128), (3000, 6)])
@pytest.mark.parametrize("metric", [0, 1, 2])
def test_scan_generic_small(api, ctx, n, d, metric):
    rng = np.random.default_rng(n * 31 + d)
    rows = rng.standard_normal((n, d)).astype(F32)
    Q = rng.standard_normal((3, d)).astype(F32)
    for k in (1, 10, n + 5):
        check(api, ctx, rows, Q, min(k, 1024), metric)


def test_scan_bench_shape_one_hot(api, ctx):                       # benches/retrieval.rs:71-94
    rows = np.zeros((1000, 128), F32)
    rows[np.arange(1000), np.arange(1000) % 128] = 1.0
    Q = np.ones((1, 128), F32)
    check(api, ctx, rows, Q, 10)
    check(api, ctx, rows, Q, 100)                                    # all 1000 scores tie -> ordinals 0..99


def test_scan_zero_vectors_and_zero_query(api, ctx):
    rng = np.random.default_rng(5)
    rows = rng.standard_normal((300, 16)).astype(F32)
    rows[[0, 17, 299]] = 0.0
    check(api, ctx, rows, np.zeros((1, 16), F32), 20)                # cosine 0.0 for every row -> ordinal order
    check(api, ctx, rows, rng.standard_normal((2, 16)).astype(F32), 300)


# ------------------------------------------------------------------ bulk (TMA) kernel
@pytest.mark.parametrize("n,d,dtype", [(5000, 384, 0), (20011, 384, 0), (8192, 768, 1), (30001, 768, 1), (6000, 4096, 1),
                                        (5000, 1024, 0), (4500, 8, 1), (4097, 4, 0)])
def test_scan_bulk(api, ctx, n, d, dtype):
    f, b = O.synth_corpus(SEED, 0, n, d, bf16=bool(dtype), dups=True)
    rows = b if dtype else f
    Q = O.synth_queries(SEED, 0, 3, d, n, corpus_bf16=bool(dtype), dups=True)
    Q[0] = f[n // 2]                                                  # an exact corpus row
    for k in (1, 10, 100):
        check(api, ctx, rows, Q, k, 0, dtype, expect_mode=1)


@pytest.mark.parametrize("metric", [1, 2])
def test_scan_bulk_metrics(api, ctx, metric):
    rng = np.random.default_rng(7)
    rows = (rng.standard_normal((9000, 384)) * rng.uniform(0.5, 2.0, (9000, 1))).astype(F32)
    Q = rng.standard_normal((2, 384)).astype(F32)
    check(api, ctx, rows, Q, 25, metric)
    check(api, ctx, bf16_round(rows).view(np.uint32).__rshift__(16).astype(np.uint16), Q, 25, metric, 1)


def test_scan_bulk_exact_ties_and_tombstones(api, ctx):
    n, d = 12000, 384
    f, _ = O.synth_corpus(SEED, 0, n, d)
    f[5000:5040] = f[77]                                              # 41 identical rows
    Q = np.stack([f[77], f[3]])
    alive = np.ones(n, bool)
    alive[[77, 5003, 5004, 11999, 0]] = False
    check(api, ctx, f, Q, 50, alive=alive)
    check(api, ctx, f, Q, 50, base=1_000_000)


def test_scan_k_limits(api, ctx):
    f, _ = O.synth_corpus(SEED, 0, 5000, 64)
    Q = O.synth_queries(SEED, 0, 1, 64, 5000)
    check(api, ctx, f, Q, 1024)
    ix = api.DenseIndex(ctx, 64)
    ix.append(f)
    with pytest.raises(api.TrrError):
        ix.search(Q, 1025)
    o, s, n = ix.search(Q, 0)
    assert n[0] == 0
    ix.close()


def test_empty_store_and_norms(api, ctx):
    ix = api.DenseIndex(ctx, 8)
    o, s, n = ix.search(np.ones((2, 8), F32), 5)
    assert list(n) == [0, 0]
    rng = np.random.default_rng(11)
    rows = rng.standard_normal((100, 8)).astype(F32)
    ix.append(rows[:40])
    ix.append(rows[40:])
    assert len(ix) == 100
    exp = np.array([np.sqrt(np.float32(sum_seq(r * r))) for r in rows], F32)
    assert np.array_equal(ix.norms(100), exp)
    ix.close()


def sum_seq(x):
    s = F32(0)
    for v in x.astype(F32):
        s = F32(s + v)
    return s


def test_incremental_append_between_searches(api, ctx):
    f, _ = O.synth_corpus(SEED, 0, 9000, 128)
    Q = O.synth_queries(SEED, 0, 2, 128, 9000)
    ix = api.DenseIndex(ctx, 128)
    ix.append(f[:4000])
    ix.search(Q, 5)
    ix.append(f[4000:])
    o, s, n = ix.search(Q, 5)
    eo, es, en = O.dense_search_batch(f, Q, 5)
    assert np.array_equal(o, eo) and np.array_equal(s, es)
    ix.close()


def test_device_synth_matches_oracle_generator(api, ctx):
    for dtype in (0, 1):
        ix = api.DenseIndex(ctx, 96, api.COSINE, dtype)
        ix.append_synth(SEED, 1000, 5000, dups=True)
        f, b = O.synth_corpus(SEED, 1000, 5000, 96, bf16=bool(dtype), dups=True)
        Q = O.synth_queries(SEED, 0, 2, 96, 5000, corpus_bf16=bool(dtype), dups=True)
        o, s, n = ix.search(Q, 10)
        eo, es, en = O.dense_search_batch(b if dtype else f, Q, 10)
        assert np.array_equal(o, eo) and np.array_equal(s, es)
        ix.close()


# ------------------------------------------------------------------ tensor-core fast pass (K2) + exact rescoring
def test_gemm_fast_scores_match_matmul(api, ctx):
    n, d, B = 1000, 256, 40
    f, b = O.synth_corpus(SEED, 0, n, d, bf16=True)
    Q = bf16_round(O.synth_queries(SEED, 0, B, d, n, corpus_bf16=True))
    ix = api.DenseIndex(ctx, d, api.DOT, api.BF16)
    ix.append(b)
    got = ix.debug_gemm_scores(Q, 1024)[:, :n]
    exp = Q.astype(np.float64) @ f.astype(np.float64).T
    assert np.abs(got - exp).max() < 1e-5, np.abs(got - exp).max()
    ix.close()


@pytest.mark.parametrize("n,d,dtype,B,k", [(20000, 768, 1, 200, 10), (50000, 768, 1, 130, 50), (20000, 384, 0, 64, 10),
                                            (16500, 100, 1, 17, 5), (33000, 4096, 1, 128, 10), (17000, 72, 0, 16, 3),
                                            (60000, 4096, 1, 96, 100), (40000, 256, 1, 1500, 100), (40000, 128, 1, 3000, 50)])
def test_gemm_path_matches_oracle(api, ctx, n, d, dtype, B, k):
    f, b = O.synth_corpus(SEED + 3, 0, n, d, bf16=bool(dtype), dups=True)
    rows = b if dtype else f
    Q = O.synth_queries(SEED + 3, 0, B, d, n, corpus_bf16=bool(dtype), dups=True)
    if dtype:
        Q = bf16_round(Q)
    st = check(api, ctx, rows, Q, k, 0, dtype, mode=2, expect_mode=2)
    assert st.max_fast_exact_gap <= st.eps_bound + 1e-3 * (dtype == 0)


def test_batches_above_4096_queries_are_served_in_pieces(api, ctx):
    n, d, B = 17000, 64, 4500
    f, b = O.synth_corpus(SEED + 6, 0, n, d, bf16=True)
    Q = bf16_round(O.synth_queries(SEED + 6, 0, B, d, n, corpus_bf16=True))
    check(api, ctx, b, Q, 7, 0, 1, mode=0)


@pytest.mark.parametrize("B", [300, 256, 512])
def test_gemm_one_and_two_cta_kernels_match_oracle(api, ctx, B):
    """Both tensor-core kernels must return the same bit-exact results: an even number of 128-query blocks (B = 256, 512)
    takes the cta_group::2 kernel (two SMs of a TPC share one 256 x 256 tile, the default at bench size), an odd number
    below nine (B = 300: three blocks) the cta_group::1 kernel."""
    n, d = 40000, 768
    f, b = O.synth_corpus(SEED + 4, 0, n, d, bf16=True, dups=True)
    Q = bf16_round(O.synth_queries(SEED + 4, 0, B, d, n, corpus_bf16=True, dups=True))
    st = check(api, ctx, b, Q, 50, 0, 1, mode=2, expect_mode=2)
    assert st.n_guard_fallbacks <= 2


def test_gemm_path_dot_metric_and_f32_queries(api, ctx):
    n, d, B = 30000, 512, 96
    rng = np.random.default_rng(3)
    f = (rng.standard_normal((n, d)) * rng.uniform(0.2, 3.0, (n, 1))).astype(F32)
    b = (bf16_round(f).view(np.uint32) >> 16).astype(np.uint16)
    Q = rng.standard_normal((B, d)).astype(F32)                       # NOT bf16-exact: the proof widens, fallbacks allowed
    st = check(api, ctx, b, Q, 10, 2, 1, mode=2, expect_mode=2)
    st = check(api, ctx, b, bf16_round(Q), 10, 2, 1, mode=2, expect_mode=2)


def test_gemm_path_adversarial_ties_fall_back_to_scan(api, ctx):
    n, d, B = 20000, 256, 20
    f, b = O.synth_corpus(SEED + 9, 0, n, d, bf16=True)
    b[1000:1200] = b[5]                                               # 201 identical rows: the proof must fail
    Q = bf16_round(O.synth_queries(SEED + 9, 0, B, d, n, corpus_bf16=True))
    Q[0] = (b[5].astype(np.uint32) << 16).view(F32)
    st = check(api, ctx, b, Q, 50, 0, 1, mode=2, expect_mode=2)
    assert st.n_guard_fallbacks >= 1


def test_gemm_tombstones_and_base(api, ctx):
    n, d, B = 18000, 128, 33
    f, b = O.synth_corpus(SEED + 5, 0, n, d, bf16=True)
    Q = bf16_round(O.synth_queries(SEED + 5, 0, B, d, n, corpus_bf16=True))
    alive = np.ones(n, bool)
    top = O.dense_search_batch(b, Q, 3)[0]
    alive[np.unique(top.ravel())] = False                             # remove every current winner
    check(api, ctx, b, Q, 10, 0, 1, mode=2, alive=alive, base=7_000_000)


@pytest.mark.parametrize("dtype", [0, 1])
def test_gemm_operands_follow_appends_and_removes(api, ctx, dtype):
    """VectorStore::insert / remove between batched searches (src/index.rs:359-383, 421-424): the tensor-core operands
    (bf16 shadow, scale/bias, TMA descriptors) are rebuilt lazily and the results track the oracle."""
    n, d, B = 30000, 128, 40
    f, b = O.synth_corpus(SEED + 13, 0, n, d, bf16=bool(dtype))
    rows = b if dtype else f
    Q = O.synth_queries(SEED + 13, 0, B, d, n, corpus_bf16=bool(dtype))
    if dtype:
        Q = bf16_round(Q)
    ix = api.DenseIndex(ctx, d, api.COSINE, dtype)
    ix.set_mode(2)
    alive = np.zeros(n, bool)
    for lo, hi in ((0, 17000), (17000, 17001), (17001, n)):
        ix.append(rows[lo:hi])
        alive[lo:hi] = True
        got = ix.search(Q, 10)
        assert ix.stats().mode_used == 2
        exp = O.dense_search_batch(rows[:hi], Q, 10, alive=alive[:hi])
        assert all(np.array_equal(x, y) for x, y in zip(got, exp)), hi
        victim = int(got[0][0, 0])                                     # remove the current winner of query 0
        ix.remove(victim)
        alive[victim] = False
        got = ix.search(Q, 10)
        exp = O.dense_search_batch(rows[:hi], Q, 10, alive=alive[:hi])
        assert all(np.array_equal(x, y) for x, y in zip(got, exp)), ("after remove", hi)
    ix.close()


def test_auto_mode_picks_scan_for_single_query_and_gemm_for_batches(api, ctx):
    n, d = 20000, 128
    f, b = O.synth_corpus(SEED, 0, n, d, bf16=True)
    ix = api.DenseIndex(ctx, d, api.COSINE, api.BF16)
    ix.append(b)
    ix.search(bf16_round(O.synth_queries(SEED, 0, 1, d, n, True)), 10)
    assert ix.stats().mode_used == 1
    ix.search(bf16_round(O.synth_queries(SEED, 0, 64, d, n, True)), 10)
    assert ix.stats().mode_used == 2
    ix.close()


# ------------------------------------------------------------------ tensor-core path: Euclidean metric and large k
@pytest.mark.parametrize("dtype", [0, 1])
def test_gemm_path_euclidean(api, ctx, dtype):
    """-euclidean_distance (src/index.rs:400,452-458) through the tensor-core pass: the fast pass ranks by 2 q.d - |d|^2,
    the survivors are re-scored with the reference's sequential sum of squared differences, and the candidate proof works
    on squared distances.  Bit-exact ids and scores; rows of varying norms; f32 and bf16 stores; f32 (not bf16-exact) queries."""
    n, d, B = 40000, 256, 200
    rng = np.random.default_rng(17)
    f = (rng.standard_normal((n, d)) * rng.uniform(0.5, 2.0, (n, 1))).astype(F32)
    rows = bf16_round(f) if dtype == 0 else (bf16_round(f).view(np.uint32) >> 16).astype(np.uint16)
    Q = (rng.standard_normal((B, d)) * rng.uniform(0.5, 2.0, (B, 1))).astype(F32)
    Q[:5] = f[100:105] + 0.01 * rng.standard_normal((5, d)).astype(F32)      # near-duplicates of stored rows
    for k in (1, 10, 50):
        st = check(api, ctx, rows, Q, k, 1, dtype, mode=2, expect_mode=2)
        assert st.n_guard_fallbacks <= B // 4


def test_gemm_path_euclidean_ties_and_removed_rows(api, ctx):
    n, d, B = 30000, 64, 64
    rng = np.random.default_rng(18)
    f = bf16_round(rng.standard_normal((n, d)).astype(F32))
    f[2000:2100] = f[7]                                                         # exact distance ties
    rows = (f.view(np.uint32) >> 16).astype(np.uint16)
    alive = np.ones(n, bool)
    alive[rng.integers(0, n, 500)] = False
    Q = bf16_round(rng.standard_normal((B, d)).astype(F32))
    Q[0] = f[7]
    check(api, ctx, rows, Q, 20, 1, 1, mode=2, alive=alive, expect_mode=2)


@pytest.mark.parametrize("k,B", [(200, 300), (1024, 128), (1000, 40)])
def test_gemm_path_large_k(api, ctx, k, B):
    """k up to 1024 on the tensor-core path (the reference's k is unbounded, src/index.rs:386-412): the re-scoring width
    grows with k."""
    n, d = 60000, 128
    f, b = O.synth_corpus(SEED + 31, 0, n, d, bf16=True)
    Q = bf16_round(O.synth_queries(SEED + 31, 0, B, d, n, corpus_bf16=True))
    check(api, ctx, b, Q, k, 0, 1, mode=2, expect_mode=2)
    check(api, ctx, b, Q, k, 2, 1, mode=2, expect_mode=2)


def test_auto_mode_takes_the_tensor_path_for_euclidean_batches(api, ctx):
    n, d, B = 50000, 96, 33
    rng = np.random.default_rng(19)
    f = rng.standard_normal((n, d)).astype(F32)
    Q = rng.standard_normal((B, d)).astype(F32)
    st = check(api, ctx, f, Q, 10, 1, 0, mode=0, expect_mode=2)                 # f32 store (bf16 shadow), AUTO
