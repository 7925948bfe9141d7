#!/usr/bin/env python
"""Per-config throughput table for BASELINE.json's configs[1..3] on one GPU (development helper; bench.py is the contract).

  cfg2  dense cosine 1M x 384 f32, top-10, batch 1 (K1 exact scan) and batch 256 (K2 through the bf16 shadow + exact rescoring)
  cfg3  BM25 only, Zipf corpus (vocab 1M), 8-32-term queries, top-100, batch 1 / 64 / 1024, PROBE_DOCS documents (default 10M)
Device times are the library's own CUDA events (trr_*_last_stats); wall times include the host-buffer call."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from trueno_rag_b200 import _lib, api  # noqa: E402
from trueno_rag_b200._lib import u32p, u64p  # noqa: E402


def wall(fn, n=5):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n


def cfg2(ctx):
    n, d = 1_000_000, 384
    ix = api.DenseIndex(ctx, d, api.COSINE, api.F32, capacity=n)
    ix.append_synth(0x5EED0002, 0, n)
    Q = O.synth_queries(0x5EED0002, 0, 256, d, n)
    ix.set_mode(api.MODE_SCAN)
    w = wall(lambda: ix.search(Q[:1], 10))
    st = ix.stats()
    print(f"cfg2 B=1   K1 scan : kernel {st.ms_main_kernel*1e3:7.1f} us ({n*d*4/st.ms_main_kernel/1e6:5.0f} GB/s)  "
          f"device call {st.ms_total*1e3:7.1f} us  host call {w*1e6:7.1f} us  -> {1/w:8.0f} q/s")
    ix.set_mode(api.MODE_GEMM)
    w = wall(lambda: ix.search(Q, 10))
    st = ix.stats()
    print(f"cfg2 B=256 K2 gemm : kernel {st.ms_main_kernel*1e3:7.1f} us ({2*256*n*d/st.ms_main_kernel/1e9:5.0f} TFLOP/s)  "
          f"device call {st.ms_total*1e3:7.1f} us  host call {w*1e6:7.1f} us  -> {256/w:8.0f} q/s  fallbacks {st.n_guard_fallbacks}")
    ix.set_mode(api.MODE_SCAN)
    w = wall(lambda: ix.search(Q[:16], 10), 2)
    print(f"cfg2 B=16  K1 scan (4 queries per pass): host call {w*1e6:7.1f} us -> {16/w:8.0f} q/s")
    ix.close()


def cfg3(ctx):
    L = _lib.load()
    N, V = int(os.environ.get("PROBE_DOCS", "10000000")), 1_000_000
    seed = 0x5EED0003
    cdf = O.zipf_cdf(V)
    df = np.zeros(V, np.uint32); dl = np.zeros(N, np.uint32); tot = C.c_uint64()
    api._check(L.trr_synth_bm25_count(seed, cdf.ctypes.data_as(u64p), V, 0, N, df.ctypes.data_as(u32p), dl.ctypes.data_as(u32p), C.byref(tot)))
    term_off = np.zeros(V + 1, np.uint64); np.cumsum(df, out=term_off[1:])
    P = int(term_off[-1])
    pd = np.zeros(P, np.uint32); ptf = np.zeros(P, np.uint32)
    api._check(L.trr_synth_bm25_fill(seed, cdf.ctypes.data_as(u64p), V, 0, N, term_off.ctypes.data_as(u64p), pd.ctypes.data_as(u32p), ptf.ctypes.data_as(u32p)))
    avgdl = float(np.float32(np.uint32(tot.value & 0xFFFFFFFF)) / np.float32(N))
    dev = api.Bm25Device(ctx, N, term_off, pd, ptf, dl, avgdl, api.bm25_idf_host(N, df))
    q_off, q_terms = O.synth_query_terms(seed, cdf, 0, 1024)
    dfl = np.diff(term_off)
    for B in (1, 64, 1024):
        qt, qo = q_terms[:q_off[B]], q_off[:B + 1]
        vol = int(dfl[qt].sum())
        w = wall(lambda: dev.search(qt, qo, 100), 3)
        st = dev.stats()
        print(f"cfg3 B={B:<4d} K3 bm25 top-100 ({N} docs, {P} postings): kernel {st.ms_main_kernel*1e3:9.1f} us "
              f"({8*vol/st.ms_main_kernel/1e6:5.0f} GB/s algorithmic)  host call {w*1e6:9.1f} us -> {B/w:8.0f} q/s")
    # parity on the first 4 queries against the oracle over the same CSR
    oix = O.BM25.from_csr(N, V, term_off, pd, ptf, dl, df, avgdl)
    got = dev.search(q_terms[:q_off[4]], q_off[:5], 100)
    exp = oix.search_batch(q_terms[:q_off[4]], q_off[:5], 100)
    ok = all(np.array_equal(g, e) for g, e in zip(got, exp))
    print("cfg3 parity vs oracle (4 queries, ids + scores + counts bit-exact):", ok)
    dev.close()


def hybrid_small_batches(ctx):
    """cfg4 corpus (10M x 768 bf16 + Zipf BM25), the reference's own call shape: HybridRetriever::retrieve for ONE query
    (B = 1), and small batches, through the host-buffer entry point trr_hybrid_search (H2D + kernels + D2H + sync)."""
    L = _lib.load()
    N, D, V, Cn, K = int(os.environ.get("PROBE_DOCS", "10000000")), 768, 1_000_000, 50, 10
    seed = 0x5EED0004
    dense = api.DenseIndex(ctx, D, api.COSINE, api.BF16, capacity=N)
    dense.append_synth(seed, 0, N)
    cdf = O.zipf_cdf(V)
    df = np.zeros(V, np.uint32); dl = np.zeros(N, np.uint32); tot = C.c_uint64()
    api._check(L.trr_synth_bm25_count(seed, cdf.ctypes.data_as(u64p), V, 0, N, df.ctypes.data_as(u32p), dl.ctypes.data_as(u32p), C.byref(tot)))
    term_off = np.zeros(V + 1, np.uint64); np.cumsum(df, out=term_off[1:])
    pd = np.zeros(int(term_off[-1]), np.uint32); ptf = np.zeros(int(term_off[-1]), np.uint32)
    api._check(L.trr_synth_bm25_fill(seed, cdf.ctypes.data_as(u64p), V, 0, N, term_off.ctypes.data_as(u64p), pd.ctypes.data_as(u32p), ptf.ctypes.data_as(u32p)))
    avgdl = float(np.float32(np.uint32(tot.value & 0xFFFFFFFF)) / np.float32(N))
    bm = api.Bm25Device(ctx, N, term_off, pd, ptf, dl, avgdl, api.bm25_idf_host(N, df))
    del pd, ptf
    Q = O.synth_queries(seed, 0, 64, D, N, corpus_bf16=True)
    u = Q.view(np.uint32)
    Q = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32).view(np.float32)
    q_off, q_terms = O.synth_query_terms(seed, cdf, 0, 64)
    for B in (1, 8, 64):
        qo, qt = q_off[:B + 1], q_terms[:q_off[B]]
        w = wall(lambda: api.hybrid_search(dense, bm, Q[:B], qt, qo, Cn, api.RRF, 60.0, K), 5)
        print(f"cfg4 hybrid B={B:<3d} (10M x 768 bf16 + BM25, RRF k=60, C=50, top-10), host-buffer call: {w*1e3:8.3f} ms -> {B/w:9.1f} q/s "
              f"(dense path: {'scan' if dense.stats().mode_used == 1 else 'gemm'})")
    dense.close(); bm.close()


def append(ctx):
    """trr_bm25_append vs a full rebuild: PROBE_DOCS documents on the device (default 10M), then 100k more."""
    L = _lib.load()
    N, V, dN = int(os.environ.get("PROBE_DOCS", "10000000")), 1_000_000, 100_000
    seed = 0x5EED0003
    cdf = O.zipf_cdf(V)

    def csr(lo, hi):
        n = hi - lo
        df = np.zeros(V, np.uint32); dl = np.zeros(n, np.uint32); tot = C.c_uint64()
        api._check(L.trr_synth_bm25_count(seed, cdf.ctypes.data_as(u64p), V, lo, hi, df.ctypes.data_as(u32p), dl.ctypes.data_as(u32p), C.byref(tot)))
        off = np.zeros(V + 1, np.uint64); np.cumsum(df, out=off[1:])
        P = int(off[-1])
        pd = np.zeros(max(P, 1), np.uint32); ptf = np.zeros(max(P, 1), np.uint32)
        api._check(L.trr_synth_bm25_fill(seed, cdf.ctypes.data_as(u64p), V, lo, hi, off.ctypes.data_as(u64p), pd.ctypes.data_as(u32p), ptf.ctypes.data_as(u32p)))
        return off, pd[:P], ptf[:P], dl, df, int(tot.value)

    off0, pd0, ptf0, dl0, df0, tot0 = csr(0, N)
    off1, pd1, ptf1, dl1, df1, tot1 = csr(N, N + dN)
    avg0 = float(np.float32(np.uint32(tot0 & 0xFFFFFFFF)) / np.float32(N))
    t0 = time.perf_counter()
    dev = api.Bm25Device(ctx, N, off0, pd0, ptf0, dl0, avg0, api.bm25_idf_host(N, df0))
    t_build = time.perf_counter() - t0
    df_all = (df0.astype(np.uint64) + df1).astype(np.uint32)
    avg1 = float(np.float32(np.uint32((tot0 + tot1) & 0xFFFFFFFF)) / np.float32(N + dN))
    idf1 = api.bm25_idf_host(N + dN, df_all)
    t0 = time.perf_counter()
    dev.append(dN, off1, pd1, ptf1, dl1, avg1, idf1)
    t_app = time.perf_counter() - t0
    q_off, q_terms = O.synth_query_terms(seed, cdf, 0, 8)
    got = dev.search(q_terms, q_off, 50)
    dev.close()
    # the same index from scratch (what the host layer had to do before): concatenate on the host, upload everything
    t0 = time.perf_counter()
    offA = np.zeros(V + 1, np.uint64); np.cumsum(df_all, out=offA[1:])
    pdA = np.zeros(int(offA[-1]), np.uint32); ptfA = np.zeros(int(offA[-1]), np.uint32)
    api._check(L.trr_synth_bm25_fill(seed, cdf.ctypes.data_as(u64p), V, 0, N + dN, offA.ctypes.data_as(u64p), pdA.ctypes.data_as(u32p), ptfA.ctypes.data_as(u32p)))
    t_host = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = api.Bm25Device(ctx, N + dN, offA, pdA, ptfA, np.concatenate([dl0, dl1]), avg1, idf1)
    t_full = time.perf_counter() - t0
    exp = ref.search(q_terms, q_off, 50)
    ref.close()
    ok = all(np.array_equal(a, b) for a, b in zip(got, exp))
    print(f"append {dN} docs to {N}: trr_bm25_append {t_app*1e3:.1f} ms  vs  full trr_bm25_build {t_full*1e3:.1f} ms "
          f"(+ {t_host*1e3:.0f} ms host CSR assembly); initial build {t_build*1e3:.1f} ms; identical results: {ok}")


if __name__ == "__main__":
    ctx = api.Context(0)
    which = sys.argv[1:] or ["cfg2", "cfg3"]
    if "cfg2" in which:
        cfg2(ctx)
    if "cfg3" in which:
        cfg3(ctx)
    if "append" in which:
        append(ctx)
    if "hybrid" in which:
        hybrid_small_batches(ctx)
