"""Opt-in kernels that are not the default yet (runs last in the GPU suite on purpose)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.test_gpu_bm25 import SEED, build, compare

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(ctx):
    from trueno_rag_b200 import api as a
    return a


@pytest.mark.parametrize("n_docs,n_terms,k", [(3000, 500, 10), (70000, 20000, 100), (200000, 3000, 50)])
def test_v2_kernel_opt_in_is_bit_exact(api, ctx, monkeypatch, n_docs, n_terms, k):
    """TRR_BM25_V2=1: the warp-autonomous kernel (2048-document sub-ranges, fine skip table for frequent terms; opt-in
    because it does not beat the default kernel yet) must return exactly what the default kernel and the oracle return,
    also after an append invalidates its fine skip table"""
    cdf = O.zipf_cdf(n_terms)
    doc_off, toks = O.synth_doc_tokens(SEED + 7, cdf, 0, n_docs)
    oix = O.BM25(n_terms=n_terms, doc_off=doc_off, tokens=toks)
    dev = build(api, ctx, oix, n_docs, doc_base=77)
    q_off, q_terms = O.synth_query_terms(SEED + 7, cdf, 0, 300 if n_docs < 100000 else 40)
    v1 = dev.search(q_terms, q_off, k)
    monkeypatch.setenv("TRR_BM25_V2", "1")
    v2 = dev.search(q_terms, q_off, k)
    for a_, b_ in zip(v1, v2):
        assert np.array_equal(a_, b_)
    compare(dev, oix, q_terms, q_off, k, base=77)
    dev.close()
