"""Fusion kernel parity (all six strategies) vs the oracle; bit-exact."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
F32 = np.float32


@pytest.fixture(scope="module")
def api(ctx):
    from trueno_rag_b200 import api as a
    return a


def rand_list(rng, n, universe, ties=False, dup=False):
    ids = rng.choice(universe, n, replace=dup).astype(np.uint32) if n else np.zeros(0, np.uint32)
    sc = np.sort(rng.random(n).astype(F32))[::-1].copy()
    if ties and n > 3:
        sc[1] = sc[2] = sc[3]
    return ids, sc


@pytest.mark.parametrize("strategy,param", [(O.RRF, 60.0), (O.RRF, 1.0), (O.LINEAR, 0.7), (O.LINEAR, 0.0), (O.CONVEX, 0.35),
                                             (O.DBSF, 0.0), (O.UNION, 0.0), (O.INTERSECTION, 0.0)])
def test_fuse_random_lists(api, ctx, strategy, param):
    rng = np.random.default_rng(strategy * 7 + 1)
    dense, sparse = [], []
    for b in range(64):
        nd, ns = int(rng.integers(0, 51)), int(rng.integers(0, 51))
        if b == 0:
            nd = ns = 0
        if b == 1:
            ns = 0
        dense.append(rand_list(rng, nd, 80, ties=b % 3 == 0, dup=b % 11 == 5))
        sparse.append(rand_list(rng, ns, 80, ties=b % 4 == 0, dup=b % 13 == 7))
    got = api.fuse(ctx, strategy, param, dense, sparse)
    for b in range(64):
        ei, ef = O.fuse(strategy, param, dense[b], sparse[b])
        hi, hf, hd, hs = O.hybrid_assemble(strategy, param, dense[b], sparse[b], 10 ** 6)
        gi, gf, gd, gs = got[b]
        assert np.array_equal(gi, ei), (b, gi, ei)
        assert np.array_equal(gf, ef), (b, gf - ef)
        assert np.array_equal(gd, hd, equal_nan=True) and np.array_equal(gs, hs, equal_nan=True)


def test_fuse_constant_and_single_element_lists(api, ctx):
    d = (np.array([1, 2, 3], np.uint32), np.array([0.5, 0.5, 0.5], F32))     # range < EPSILON -> all 1.0 / z-score 0.0
    s = (np.array([3], np.uint32), np.array([7.0], F32))
    for strategy, param in ((O.LINEAR, 0.4), (O.DBSF, 0.0), (O.RRF, 60.0)):
        gi, gf, _, _ = api.fuse(ctx, strategy, param, [d], [s])[0]
        ei, ef = O.fuse(strategy, param, d, s)
        assert np.array_equal(gi, ei) and np.array_equal(gf, ef)


def test_fuse_large_candidate_lists(api, ctx):
    rng = np.random.default_rng(9)
    d, s = rand_list(rng, 100, 150), rand_list(rng, 100, 150)
    for strategy, param in ((O.LINEAR, 0.7), (O.RRF, 60.0)):
        gi, gf, _, _ = api.fuse(ctx, strategy, param, [d], [s], k_out=100)[0]
        ei, ef = O.fuse(strategy, param, d, s)
        assert np.array_equal(gi, ei[:100]) and np.array_equal(gf, ef[:100])
