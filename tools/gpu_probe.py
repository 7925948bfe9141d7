"""Step-by-step GPU bring-up probe.  Each step runs in its own process (tools/run_probe.sh) so that a faulting
kernel cannot take the later steps down.  Prints compact diagnostics; used during development only."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from trueno_rag_b200 import api  # noqa: E402

F32 = np.float32
SEED = 0x5EED0002


def bf16_round(x):
    u = np.ascontiguousarray(x, F32).view(np.uint32)
    return ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32).view(F32)


def cmp_lists(tag, got, exp):
    (o, s, n), (eo, es, en) = got, exp
    bad_n = int((n != en).sum())
    bad_o = bad_s = 0
    for b in range(len(n)):
        m = int(min(n[b], en[b]))
        bad_o += int((o[b, :m] != eo[b, :m]).sum())
        bad_s += int((s[b, :m] != es[b, :m]).sum())
    print(f"[{tag}] queries={len(n)} bad_n={bad_n} bad_ids={bad_o} bad_scores={bad_s}", flush=True)
    if bad_o or bad_s or bad_n:
        b = 0
        print("   got ", o[b, :8], s[b, :8], n[b])
        print("   exp ", eo[b, :8], es[b, :8], en[b])
    return not (bad_n or bad_o or bad_s)


def step_ctx():
    ctx = api.Context(0)
    print("sm_count", ctx.sm_count, "stream", hex(ctx.stream))


def step_scan_small():
    ctx = api.Context(0)
    rng = np.random.default_rng(0)
    for (n, d) in [(3, 3), (1000, 128), (3000, 6)]:
        rows = rng.standard_normal((n, d)).astype(F32)
        Q = rng.standard_normal((3, d)).astype(F32)
        for metric in (0, 1, 2):
            ix = api.DenseIndex(ctx, d, metric)
            ix.append(rows)
            got = ix.search(Q, 10)
            cmp_lists(f"scan-generic n={n} d={d} m={metric}", got, O.dense_search_batch(rows, Q, 10, metric=metric))
            ix.close()


def step_scan_bulk():
    ctx = api.Context(0)
    for (n, d, dt) in [(20011, 384, 0), (30001, 768, 1), (6000, 4096, 1)]:
        f, b = O.synth_corpus(SEED, 0, n, d, bf16=bool(dt), dups=True)
        rows = b if dt else f
        Q = O.synth_queries(SEED, 0, 3, d, n, corpus_bf16=bool(dt), dups=True)
        ix = api.DenseIndex(ctx, d, 0, dt)
        ix.append(rows)
        ix.set_mode(1)
        for k in (10, 100):
            got = ix.search(Q, k)
            st = ix.stats()
            cmp_lists(f"scan-bulk n={n} d={d} dt={dt} k={k} ms={st.ms_main_kernel:.3f}", got, O.dense_search_batch(rows, Q, k))
        ix.close()


def step_gemm_debug():
    ctx = api.Context(0)
    n, d, B = 1000, 256, 40
    f, b = O.synth_corpus(SEED, 0, n, d, bf16=True)
    Q = bf16_round(O.synth_queries(SEED, 0, B, d, n, corpus_bf16=True))
    ix = api.DenseIndex(ctx, d, api.DOT, api.BF16)
    ix.append(b)
    got = ix.debug_gemm_scores(Q, 1024)[:, :n]
    exp = (Q.astype(np.float64) @ f.astype(np.float64).T)
    err = np.abs(got - exp)
    print("gemm-debug max_err", err.max(), "mean_err", err.mean(), "got[0,:4]", got[0, :4], "exp[0,:4]", exp[0, :4])
    if err.max() > 1e-4:
        bad = np.argwhere(err > 1e-4)
        print("  bad count", len(bad), "first", bad[:10].tolist())
        print("  rows with errors", np.unique(bad[:, 0])[:20], "cols", np.unique(bad[:, 1])[:40])
        # does got match some permutation / partial K?
        for kk in (16, 32, 64, 128):
            part = Q[:, :kk].astype(np.float64) @ f[:, :kk].astype(np.float64).T
            print(f"  vs partial K={kk}: max_err {np.abs(got - part).max():.4g}")
    ix.close()


def step_gemm_path():
    ctx = api.Context(0)
    for (n, d, dt, B, k) in [(20000, 768, 1, 200, 10), (50000, 768, 1, 130, 50), (20000, 384, 0, 64, 10)]:
        f, b = O.synth_corpus(SEED + 3, 0, n, d, bf16=bool(dt), dups=True)
        rows = b if dt else f
        Q = O.synth_queries(SEED + 3, 0, B, d, n, corpus_bf16=bool(dt), dups=True)
        if dt:
            Q = bf16_round(Q)
        ix = api.DenseIndex(ctx, d, 0, dt)
        ix.append(rows)
        ix.set_mode(2)
        got = ix.search(Q, k)
        st = ix.stats()
        cmp_lists(f"gemm n={n} d={d} dt={dt} B={B} k={k} fallbacks={st.n_guard_fallbacks} gap={st.max_fast_exact_gap:.2e} "
                  f"eps={st.eps_bound:.2e} ms={st.ms_main_kernel:.3f}", got, O.dense_search_batch(rows, Q, k))
        ix.close()


def step_bm25():
    ctx = api.Context(0)
    for (n_docs, n_terms) in [(3000, 500), (70000, 20000)]:
        cdf = O.zipf_cdf(n_terms)
        doc_off, toks = O.synth_doc_tokens(SEED, cdf, 0, n_docs)
        oix = O.BM25(n_terms=n_terms, doc_off=doc_off, tokens=toks)
        term_off, post_doc, post_tf, doc_len, df = oix.csr()
        dev = api.Bm25Device(ctx, n_docs, term_off, post_doc, post_tf, doc_len, oix.avgdl, api.bm25_idf_host(n_docs, df))
        q_off, q_terms = O.synth_query_terms(SEED, cdf, 0, 40)
        for k in (10, 100):
            got = dev.search(q_terms, q_off, k)
            st = dev.stats()
            cmp_lists(f"bm25 docs={n_docs} k={k} ms={st.ms_main_kernel:.3f}", got, oix.search_batch(q_terms, q_off, k))
        dev.close()


def step_fusion():
    ctx = api.Context(0)
    rng = np.random.default_rng(1)
    ok = True
    for strategy, param in ((0, 60.0), (1, 0.7), (3, 0.0), (4, 0.0), (5, 0.0)):
        dense, sparse = [], []
        for b in range(32):
            nd, ns = int(rng.integers(0, 51)), int(rng.integers(0, 51))
            dense.append((rng.choice(80, nd, replace=False).astype(np.uint32), np.sort(rng.random(nd).astype(F32))[::-1].copy()))
            sparse.append((rng.choice(80, ns, replace=False).astype(np.uint32), np.sort(rng.random(ns).astype(F32))[::-1].copy()))
        got = api.fuse(ctx, strategy, param, dense, sparse)
        bad = 0
        for b in range(32):
            ei, ef = O.fuse(strategy, param, dense[b], sparse[b])
            gi, gf, _, _ = got[b]
            bad += int(not (np.array_equal(gi, ei) and np.array_equal(gf, ef)))
        print(f"[fusion strategy={strategy}] bad_queries={bad}")
        ok = ok and bad == 0


def step_perf_scan():
    ctx = api.Context(0)
    n, d = int(os.environ.get("PROBE_DOCS", "1000000")), int(os.environ.get("PROBE_DIM", "384"))
    dt = int(os.environ.get("PROBE_DTYPE", "0"))
    ix = api.DenseIndex(ctx, d, 0, dt, capacity=n)
    ix.append_synth(SEED, 0, n)
    ix.set_mode(1)
    Q = O.synth_queries(SEED, 0, 4, d, n)
    for it in range(5):
        ix.search(Q[:1], 10)
        st = ix.stats()
        print(f"K1 {n}x{d} dtype={dt} B=1: main {st.ms_main_kernel*1e3:.1f} us total {st.ms_total*1e3:.1f} us  "
              f"{n*d*(2 if dt else 4)/st.ms_main_kernel/1e6:.0f} GB/s")
    f, _ = O.synth_corpus(SEED, 0, 200000, d)
    got = ix.search(Q[:1], 10)
    print("top1", got[0][0, :5], got[1][0, :5])
    ix.close()


def step_perf_metrics():
    """Batched search through the tensor-core path for each metric on the cfg4 dense shape (PROBE_DOCS x 768 bf16, B=1024,
    k=50): the Euclidean batch against the cosine batch (within 1.3x)."""
    n, d, B, k = int(os.environ.get("PROBE_DOCS", "10000000")), 768, 1024, 50
    Q = O.synth_queries(SEED, 0, B, d, n, corpus_bf16=True)
    for metric, name in ((0, "cosine"), (1, "euclidean"), (2, "dot")):
        ctx = api.Context(0)
        ix = api.DenseIndex(ctx, d, metric, 1, capacity=n)
        ix.append_synth(SEED, 0, n)
        ix.set_mode(2)
        for it in range(4):
            ix.search(Q, k)
            st = ix.stats()
        print(f"K2 {name}: {n}x{d} bf16 B={B} k={k}: call {st.ms_total:.3f} ms, tensor-core pass {st.ms_main_kernel:.3f} ms, "
              f"re-scoring width {st.rescore_width}, exact-scan fallbacks {st.n_guard_fallbacks}", flush=True)
        ix.close(); ctx.close()


def step_perf_gemm():
    ctx = api.Context(0)
    n, d, B = int(os.environ.get("PROBE_DOCS", "2000000")), int(os.environ.get("PROBE_DIM", "768")), int(os.environ.get("PROBE_B", "1024"))
    ix = api.DenseIndex(ctx, d, 0, 1, capacity=n)
    t0 = time.time()
    ix.append_synth(SEED, 0, n)
    print("synth gen s", time.time() - t0)
    ix.set_mode(2)
    Q = bf16_round(O.synth_queries(SEED, 0, B, d, n, corpus_bf16=True))
    for it in range(int(os.environ.get("PROBE_ITERS", "3"))):
        ix.search(Q, 50)
        st = ix.stats()
        fl = 2.0 * B * n * d
        print(f"K2 {n}x{d} bf16 B={B}: main {st.ms_main_kernel:.3f} ms total {st.ms_total:.3f} ms "
              f"{fl/st.ms_main_kernel/1e9:.0f} TFLOP/s fallbacks={st.n_guard_fallbacks} gap={st.max_fast_exact_gap:.2e}")
    ix.close()


def step_ab_pair():
    """alternates GEMM kernel variants in one process (same thermal / power state): (2-CTA?, TRR_GEMM_DEBUG bits)"""
    ctx = api.Context(0)
    n, d, B = int(os.environ.get("PROBE_DOCS", "10000000")), 768, 1024
    ix = api.DenseIndex(ctx, d, 0, 1, capacity=n)
    ix.append_synth(SEED, 0, n)
    ix.set_mode(2)
    Q = bf16_round(O.synth_queries(SEED, 0, B, d, n, corpus_bf16=True))
    variants = [tuple(int(x) for x in v.split(":")) for v in os.environ.get("PROBE_VARIANTS", "0:0,0:16,1:0,1:16").split(",")]
    acc = {v: [] for v in variants}
    ref = None
    for it in range(int(os.environ.get("PROBE_ITERS", "8")) * len(variants)):
        v = variants[it % len(variants)]
        os.environ["TRR_GEMM_PAIR"] = str(v[0])
        os.environ["TRR_GEMM_DEBUG"] = str(v[1])
        ids, sc, cnt = ix.search(Q, 50)
        if (v[1] & 7) == 0:
            if ref is None:
                ref = (ids.copy(), sc.copy())
            elif not (np.array_equal(ref[0], ids) and np.array_equal(ref[1], sc)):
                print("MISMATCH between variants", v)
        acc[v].append(ix.stats().ms_main_kernel)
    for v in variants:
        x = acc[v][2:]
        print(f"pair={v[0]} debug={v[1]}: main ms {' '.join('%.2f' % t for t in acc[v])}  mean(after 2) {sum(x)/len(x):.3f}")
    ix.close()


def step_perf_bm25():
    ctx = api.Context(0)
    n_docs, n_terms, B = 2_000_000, 200_000, 512
    cdf = O.zipf_cdf(n_terms)
    t0 = time.time()
    doc_off, toks = O.synth_doc_tokens(SEED, cdf, 0, n_docs)
    oix = O.BM25(n_terms=n_terms, doc_off=doc_off, tokens=toks)
    term_off, post_doc, post_tf, doc_len, df = oix.csr()
    print("host gen+build s", time.time() - t0, "postings", oix.n_postings)
    t0 = time.time()
    dev = api.Bm25Device(ctx, n_docs, term_off, post_doc, post_tf, doc_len, oix.avgdl, api.bm25_idf_host(n_docs, df))
    print("device build s", time.time() - t0)
    q_off, q_terms = O.synth_query_terms(SEED, cdf, 0, B)
    vol = int(np.diff(term_off)[q_terms].sum())
    for it in range(3):
        got = dev.search(q_terms, q_off, 100)
        st = dev.stats()
        print(f"K3 2M docs B={B}: main {st.ms_main_kernel:.3f} ms  postings/query {vol/B:.0f}  {8*vol/st.ms_main_kernel/1e6:.0f} GB/s "
              f"{B/st.ms_main_kernel*1e3:.0f} q/s")
    exp = oix.search_batch(q_terms[:q_off[8]], q_off[:9], 100)
    cmp_lists("bm25 2M check", (got[0][:8], got[1][:8], got[2][:8]), exp)
    dev.close()


def step_perf_bm25_big():
    """cfg4-shaped BM25 leg only (Zipf vocab 1M, B=1024, k=50) on PROBE_DOCS documents (default 4M)."""
    import ctypes as C
    from trueno_rag_b200 import _lib
    from trueno_rag_b200._lib import u32p, u64p
    L = _lib.load()
    ctx = api.Context(0)
    N, V, B, K = int(os.environ.get("PROBE_DOCS", "4000000")), 1_000_000, 1024, 50
    seed = 0x5EED0004
    cdf = O.zipf_cdf(V)
    df = np.zeros(V, np.uint32); dl = np.zeros(N, np.uint32); tot = C.c_uint64()
    api._check(L.trr_synth_bm25_count(seed, cdf.ctypes.data_as(u64p), V, 0, N, df.ctypes.data_as(u32p), dl.ctypes.data_as(u32p), C.byref(tot)))
    term_off = np.zeros(V + 1, np.uint64); np.cumsum(df, out=term_off[1:])
    P = int(term_off[-1])
    pd = np.zeros(P, np.uint32); ptf = np.zeros(P, np.uint32)
    api._check(L.trr_synth_bm25_fill(seed, cdf.ctypes.data_as(u64p), V, 0, N, term_off.ctypes.data_as(u64p), pd.ctypes.data_as(u32p), ptf.ctypes.data_as(u32p)))
    avgdl = float(np.float32(np.uint32(tot.value & 0xFFFFFFFF)) / np.float32(N))
    dev = api.Bm25Device(ctx, N, term_off, pd, ptf, dl, avgdl, api.bm25_idf_host(N, df))
    q_off, q_terms = O.synth_query_terms(seed, cdf, 0, B)
    vol = int(np.diff(term_off)[q_terms].sum())
    for it in range(4):
        got = dev.search(q_terms, q_off, K)
        st = dev.stats()
        print(f"K3 {N} docs B={B} k={K}: main {st.ms_main_kernel:.3f} ms  postings/query {vol/B:.0f}  {8*vol/st.ms_main_kernel/1e6:.0f} GB/s "
              f"{B/st.ms_main_kernel*1e3:.0f} q/s  fallbacks {st.n_guard_fallbacks}/{st.n_exact_fallbacks} launches {st.n_kernel_launches}", flush=True)
    print("checksum", int(got[0].astype(np.uint64).sum()), float(got[1].astype(np.float64).sum()), int(got[2].sum()))
    dev.close()


def step_corun():
    """Feasibility probe: tensor-core dense pass and BM25 kernel on two streams of one GPU (needs a build whose GEMM kernel
    and BM25 kernel each fit half of the shared memory: TRR_BUILD_DEFS='-DTRR_GEMM_STAGES=2 -DTRR_GEMM_LSTRIDE=17',
    TRR_BM25_RANGE_SHIFT=14 TRR_BM25_CTAS_PER_SM=1)."""
    import ctypes as C
    import torch
    from trueno_rag_b200 import _lib
    from trueno_rag_b200._lib import u32p, u64p
    L = _lib.load()
    ctxA, ctxB = api.Context(0), api.Context(0)
    N, D, V, B, K = int(os.environ.get("PROBE_DOCS", "4000000")), 768, 1_000_000, 1024, 50
    seed = 0x5EED0004
    dense = api.DenseIndex(ctxA, D, 0, 1, capacity=N)
    dense.append_synth(seed, 0, N)
    dense.set_mode(2)
    cdf = O.zipf_cdf(V)
    df = np.zeros(V, np.uint32); dl = np.zeros(N, np.uint32); tot = C.c_uint64()
    api._check(L.trr_synth_bm25_count(seed, cdf.ctypes.data_as(u64p), V, 0, N, df.ctypes.data_as(u32p), dl.ctypes.data_as(u32p), C.byref(tot)))
    term_off = np.zeros(V + 1, np.uint64); np.cumsum(df, out=term_off[1:])
    P = int(term_off[-1])
    pd = np.zeros(P, np.uint32); ptf = np.zeros(P, np.uint32)
    api._check(L.trr_synth_bm25_fill(seed, cdf.ctypes.data_as(u64p), V, 0, N, term_off.ctypes.data_as(u64p), pd.ctypes.data_as(u32p), ptf.ctypes.data_as(u32p)))
    avgdl = float(np.float32(np.uint32(tot.value & 0xFFFFFFFF)) / np.float32(N))
    bm = api.Bm25Device(ctxB, N, term_off, pd, ptf, dl, avgdl, api.bm25_idf_host(N, df))
    Q = bf16_round(O.synth_queries(seed, 0, B, D, N, corpus_bf16=True))
    q_off, q_terms = O.synth_query_terms(seed, cdf, 0, B)
    dev = torch.device("cuda", 0)
    d_q = torch.from_numpy(Q).to(dev)
    d_terms = torch.from_numpy(q_terms.view(np.int32).copy()).to(dev)
    d_off = torch.from_numpy(q_off.view(np.int32).copy()).to(dev)
    outs = [torch.zeros((B, K), dtype=torch.int32, device=dev), torch.zeros((B, K), dtype=torch.float32, device=dev),
            torch.zeros(B, dtype=torch.int32, device=dev)]
    outs2 = [torch.zeros((B, K), dtype=torch.int32, device=dev), torch.zeros((B, K), dtype=torch.float32, device=dev),
             torch.zeros(B, dtype=torch.int32, device=dev)]
    torch.cuda.synchronize()

    def run_dense():
        dense.search_device(d_q.data_ptr(), B, K, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr())

    def run_bm25():
        api._check(L.trr_bm25_search_device(bm.h, C.c_void_p(d_terms.data_ptr()), C.c_void_p(d_off.data_ptr()),
                                            q_off.ctypes.data_as(u32p), B, K, C.c_void_p(outs2[0].data_ptr()),
                                            C.c_void_p(outs2[1].data_ptr()), C.c_void_p(outs2[2].data_ptr())))

    def sync():
        api._check(L.trr_ctx_sync(ctxA.h)); api._check(L.trr_ctx_sync(ctxB.h))

    for name, fns in (("dense alone", [run_dense]), ("bm25 alone", [run_bm25]), ("both, two streams", [run_dense, run_bm25]),
                      ("both, bm25 first", [run_bm25, run_dense])):
        for f in fns:
            f()
        sync()
        t0 = time.perf_counter()
        for _ in range(5):
            for f in fns:
                f()
        sync()
        print(f"{name}: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms per batch", flush=True)
    ref = (outs[0].cpu().numpy().copy(), outs2[0].cpu().numpy().copy())
    print("checksums", int(ref[0].astype(np.int64).sum()), int(ref[1].astype(np.int64).sum()))


if __name__ == "__main__":
    globals()["step_" + sys.argv[1]]()
