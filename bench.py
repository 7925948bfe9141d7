#!/usr/bin/env python
"""bench.py — hybrid top-10 queries/s on the BASELINE.json workload (cfg4): 10M x 768 bf16 embeddings + Zipfian BM25
corpus (1M-term vocabulary), batch 1024, candidates_per_source 50, RRF k=60, top-10, corpus sharded by document over
the GPUs of one node (one process per GPU).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU algorithm (oracle port; the Rust reference cannot
                                              # be built in this image) on the host cores, bounded sample

One JSON line on stdout (rank 0).  A "step" = one pass of the hot path over one batch of 1024 synthetic queries:
shard-local dense top-C (tcgen05 GEMM + fused top-k + exact rescoring) and BM25 top-C, all-gather of the shard lists,
merge + fusion + top-k.  `value` times it with inputs resident in HBM; `e2e` times the same batch through the host-buffer
C-ABI calls (host->device query copy and device->host result copy inside the timed region).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0x5EED0004
METRIC = "hybrid top-10 queries/s at 10Mx768 (1/2/4/8 GPU); % HBM/tensor roofline"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--cands", type=int, default=50)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--cpu-sample-docs", type=int, default=400_000)
    ap.add_argument("--cpu-sample-queries", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bf16-queries", action="store_true", help="round the query embeddings to bf16 (default: full f32 queries, as an embedder emits them)")
    ap.add_argument("--verify", type=int, default=4, help="queries checked against the oracle after the timed region (0 = off)")
    return ap.parse_args()


def config(a, extra=None):
    c = {"workload": f"cfg4 hybrid dense+BM25 RRF k=60: {a.docs}x{a.dim} bf16, Zipf BM25 vocab {a.vocab}, "
                     f"batch {a.batch}, C={a.cands}, top-{a.k}",
         "docs": a.docs, "dim": a.dim, "batch": a.batch, "vocab": a.vocab, "candidates_per_source": a.cands, "k": a.k,
         "fusion": "RRF k=60", "exchange": "all-gather + merge of step i on a second stream, overlapped with the shard-local kernels of step i+1" if a.gpus > 1 else "none (one shard)",
         "queries": "bf16-rounded" if getattr(a, "bf16_queries", False) else "f32 (embedder output; the store is bf16)",
         "sharding": f"documents, contiguous ranges over {a.gpus} GPU(s)",
         "l2": "inputs (>=1.9 GB of embeddings per GPU) exceed the 126 MB L2; no flush needed"}
    if extra:
        c.update(extra)
    return c


def zipf_cdf(n_terms: int, clip: int = 90) -> np.ndarray:
    """u64 CDF of the clipped Zipf(s=1) over ranks clip+1 .. clip+n_terms; same table as the oracle's."""
    r = np.arange(clip + 1, clip + 1 + n_terms, dtype=np.float64)
    c = np.cumsum(1.0 / r)
    c /= c[-1]
    t = np.minimum(np.floor(c * 18446744073709551616.0), 18446744073709549568.0).astype(np.uint64)
    t[-1] = np.uint64(0xFFFFFFFFFFFFFFFF)
    return t


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- CPU legs
def oracle_hybrid_sample(a, n_docs, n_q, threads):
    """Builds the bounded CPU sample (first n_docs documents of the same synthetic corpus, first n_q queries) and returns
    a closure that runs one hybrid batch through the oracle with `threads` host threads."""
    from oracle import oracle as O
    rows_f32, rows_bf16 = O.synth_corpus(SEED, 0, n_docs, a.dim, bf16=True)
    q = O.synth_queries(SEED, 0, n_q, a.dim, a.docs, corpus_bf16=True)
    cdf = O.zipf_cdf(a.vocab)
    doc_off, toks = O.synth_doc_tokens(SEED, cdf, 0, n_docs)
    q_off, q_terms = O.synth_query_terms(SEED, cdf, 0, n_q)
    ix = O.BM25(n_terms=a.vocab, doc_off=doc_off, tokens=toks)

    def run():
        d = O.dense_search_batch(rows_bf16, q, a.cands, literal=True, threads=threads)   # full sort, as the reference does
        s = ix.search_batch(q_terms, q_off, a.cands, threads=threads)
        out = []
        for b in range(n_q):
            out.append(O.hybrid_assemble(O.RRF, 60.0, (d[0][b, :d[2][b]], d[1][b, :d[2][b]]),
                                         (s[0][b, :s[2][b]], s[1][b, :s[2][b]]), a.k))
        return out
    return run


def cpu_baseline(a):
    """Single-thread oracle port (what the scalar, single-threaded reference does) on a bounded sample."""
    n_docs, n_q = min(a.cpu_sample_docs, a.docs), a.cpu_sample_queries
    run = oracle_hybrid_sample(a, n_docs, n_q, threads=1)
    t0 = time.perf_counter()
    run()
    dt = time.perf_counter() - t0
    qps_sample = n_q / dt
    return {"value": qps_sample * n_docs / a.docs, "unit": "queries/s", "cores": 1, "kind": "port",
            "sample": f"oracle port of the reference algorithm, 1 thread, first {n_docs} docs x {n_q} queries: "
                      f"{qps_sample:.3f} q/s measured; value = that scaled linearly to {a.docs} docs (the path is O(N) per query)"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_docs = min(a.cpu_sample_docs, a.docs)
    n_q = max(threads, a.cpu_sample_queries)
    run = oracle_hybrid_sample(a, n_docs, n_q, threads=threads)
    for _ in range(a.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        run()
    dt = (time.perf_counter() - t0) / max(a.steps, 1)
    qps = n_q / dt * n_docs / a.docs
    sample = (f"oracle port of the reference's scalar CPU algorithm (the Rust reference cannot be compiled in this image), "
              f"{threads} host threads over queries; each step = first {n_docs} docs x {n_q} queries; value scaled linearly to "
              f"{a.docs} docs")
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config(a),
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- our arm
def run_ours(a):
    # torchrun sets OMP_NUM_THREADS=1; the host-side synthetic generators (OpenMP) should use this rank's share of cores
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(world_env, 1)))
    import torch
    import torch.distributed as dist
    from trueno_rag_b200 import api, shard, _lib
    from trueno_rag_b200._lib import f32p, u32p, u64p

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.load()
    ctx = api.Context(local_rank)
    stream = torch.cuda.current_stream()
    api._check(L.trr_ctx_set_stream(ctx.h, C.c_void_p(stream.cuda_stream)))

    B, D, Cn, K, V, N = a.batch, a.dim, a.cands, a.k, a.vocab, a.docs
    lo, hi = shard.shard_range(N, rank, world)
    n_loc = hi - lo
    t_setup = time.time()
    # ---- dense shard, generated on the device
    dense = api.DenseIndex(ctx, D, api.COSINE, api.BF16, capacity=n_loc, base=lo)
    dense.append_synth(SEED, lo, n_loc)
    dense.set_mode(api.MODE_GEMM)
    # ---- BM25 shard: host generation, GLOBAL statistics via all-reduce
    cdf = zipf_cdf(V)
    df_loc = np.zeros(V, np.uint32)
    doc_len = np.zeros(max(n_loc, 1), np.uint32)
    tot = C.c_uint64()
    api._check(L.trr_synth_bm25_count(SEED, cdf.ctypes.data_as(u64p), V, lo, hi, df_loc.ctypes.data_as(u32p),
                                      doc_len.ctypes.data_as(u32p), C.byref(tot)))
    df_glob = torch.from_numpy(df_loc.astype(np.int64)).to(dev)
    tot_glob = torch.tensor([tot.value], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(df_glob)
        dist.all_reduce(tot_glob)
    df_g = df_glob.cpu().numpy().astype(np.uint32)
    total_u32 = np.uint32(int(tot_glob.item()) & 0xFFFFFFFF)                      # u32 sum (src/index.rs:161)
    avgdl = float(np.float32(total_u32) / np.float32(N))
    idf = api.bm25_idf_host(N, df_g)
    term_off = np.zeros(V + 1, np.uint64)
    np.cumsum(df_loc, out=term_off[1:])
    P = int(term_off[-1])
    post_doc = np.zeros(max(P, 1), np.uint32)
    post_tf = np.zeros(max(P, 1), np.uint32)
    api._check(L.trr_synth_bm25_fill(SEED, cdf.ctypes.data_as(u64p), V, lo, hi, term_off.ctypes.data_as(u64p),
                                     post_doc.ctypes.data_as(u32p), post_tf.ctypes.data_as(u32p)))
    bm = api.Bm25Device(ctx, n_loc, term_off, post_doc, post_tf, doc_len[:n_loc], avgdl, idf, doc_base=lo)
    host_csr = (term_off, post_doc, post_tf, doc_len[:n_loc], df_g, avgdl) if (world == 1 and a.verify > 0) else None
    del post_doc, post_tf
    # ---- queries: f32 as an embedder emits them (the corpus is bf16); pinned host copies + device copies
    q_pin = torch.empty((B, D), dtype=torch.float32).pin_memory()
    q_np = q_pin.numpy()
    api._check(L.trr_synth_queries(SEED, 0, B, D, N, 1, 0, 1 if a.bf16_queries else 0, q_np.ctypes.data_as(f32p)))
    q_off = np.zeros(B + 1, np.uint32)
    api._check(L.trr_synth_query_terms(SEED, cdf.ctypes.data_as(u64p), V, 0, B, q_off.ctypes.data_as(u32p), None, 0))
    nt = int(q_off[-1])
    terms_pin = torch.empty(max(nt, 1), dtype=torch.int32).pin_memory()
    q_terms = terms_pin.numpy().view(np.uint32)
    api._check(L.trr_synth_query_terms(SEED, cdf.ctypes.data_as(u64p), V, 0, B, q_off.ctypes.data_as(u32p),
                                       q_terms.ctypes.data_as(u32p), nt))
    d_q = q_pin.to(dev)
    d_terms = terms_pin.to(dev)
    d_off = torch.from_numpy(q_off.view(np.int32).copy()).to(dev)
    rec_bytes = api.exchange_bytes(B, Cn)
    d_rec = torch.zeros(rec_bytes, dtype=torch.uint8, device=dev)
    d_gath = torch.zeros(world * rec_bytes, dtype=torch.uint8, device=dev)
    d_out = [torch.zeros((B, K), dtype=torch.int32, device=dev)] + \
            [torch.zeros((B, K), dtype=torch.float32, device=dev) for _ in range(3)] + \
            [torch.zeros(B, dtype=torch.int32, device=dev)]
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup
    postings_per_batch_local = int(np.diff(term_off)[q_terms[:nt]].sum())

    def gather():
        if world > 1:
            dist.all_gather_into_tensor(d_gath, d_rec)
            return d_gath
        return d_rec

    # Sharded runs overlap the exchange with compute: the shard-local kernels of step i+1 run on the compute stream while the
    # all-gather and the merge + fusion kernel of step i run on a second stream (exchange records and gather buffers are
    # double-buffered; events order record reuse).  This hides the all-gather latency and the skew between ranks.
    pipelined = world > 1 and not os.environ.get("TRR_BENCH_NO_PIPELINE")
    if pipelined:
        s_main = torch.cuda.current_stream()
        s_xchg = torch.cuda.Stream()
        ctx_x = api.Context(local_rank)
        api._check(L.trr_ctx_set_stream(ctx_x.h, C.c_void_p(s_xchg.cuda_stream)))
        recs = [d_rec, torch.zeros_like(d_rec)]
        gaths = [d_gath, torch.zeros_like(d_gath)]
        ev_local = [torch.cuda.Event(), torch.cuda.Event()]
        ev_gath = [torch.cuda.Event(), torch.cuda.Event()]
        for e in ev_gath:
            e.record(s_xchg)
        step_no = [0]

    def step_device():
        if not pipelined:
            api._check(L.trr_hybrid_local_device(dense.h, bm.h, C.c_void_p(d_q.data_ptr()), C.c_void_p(d_terms.data_ptr()),
                                                 C.c_void_p(d_off.data_ptr()), q_off.ctypes.data_as(u32p), B, Cn, 1, 1,
                                                 C.c_void_p(d_rec.data_ptr())))
            g = gather()
            api._check(L.trr_hybrid_merge_device(ctx.h, C.c_void_p(g.data_ptr()), world, B, Cn, api.RRF, 60.0, K,
                                                 *[C.c_void_p(t.data_ptr()) for t in d_out]))
            return
        p = step_no[0] & 1
        step_no[0] += 1
        s_main.wait_event(ev_gath[p])                      # the all-gather of step i-2 has consumed recs[p]
        api._check(L.trr_hybrid_local_device(dense.h, bm.h, C.c_void_p(d_q.data_ptr()), C.c_void_p(d_terms.data_ptr()),
                                             C.c_void_p(d_off.data_ptr()), q_off.ctypes.data_as(u32p), B, Cn, 1, 1,
                                             C.c_void_p(recs[p].data_ptr())))
        ev_local[p].record(s_main)
        with torch.cuda.stream(s_xchg):
            s_xchg.wait_event(ev_local[p])
            dist.all_gather_into_tensor(gaths[p], recs[p])
            ev_gath[p].record(s_xchg)
            api._check(L.trr_hybrid_merge_device(ctx_x.h, C.c_void_p(gaths[p].data_ptr()), world, B, Cn, api.RRF, 60.0, K,
                                                 *[C.c_void_p(t.data_ptr()) for t in d_out]))

    def step_host_buffers():
        """The blocking host-buffer calls (each copies in/out and synchronises): used for verification."""
        api.hybrid_local(dense, bm, q_np, q_terms[:nt], q_off, Cn, d_rec.data_ptr())       # H2D of queries inside
        g = gather()
        o = api.hybrid_merge(ctx, g.data_ptr(), world, B, Cn, api.RRF, 60.0, K)            # D2H of results inside
        return o

    # End-to-end serving step: every step copies its inputs from PINNED host memory to the device (asynchronously, on the
    # compute stream, in front of the kernels that read them), runs the device-resident step, and copies the step's results
    # back into pinned host buffers behind the merge kernel.  Nothing waits on the host inside the loop, so the copies of
    # one step overlap the kernels of its neighbours; the closing barrier of the timed region waits for the last copy.
    off_pin = torch.from_numpy(q_off.view(np.int32).copy()).pin_memory()
    out_pin = [torch.empty_like(t, device="cpu").pin_memory() for t in d_out]

    def step_e2e():
        d_q.copy_(q_pin, non_blocking=True)
        d_terms.copy_(terms_pin, non_blocking=True)
        d_off.copy_(off_pin, non_blocking=True)
        step_device()
        with torch.cuda.stream(s_xchg if pipelined else torch.cuda.current_stream()):
            for hp, dt in zip(out_pin, d_out):
                hp.copy_(dt, non_blocking=True)
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, collect=None):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            r = fn()
            if collect is not None:
                collect(r)
        if pipelined:
            torch.cuda.current_stream().wait_stream(s_xchg)   # the timed region ends when the last merge has finished
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    gemm_ms, bm25_ms, fallbacks = [], [], []

    def collect_stats(_):
        sd, sb = dense.stats(), bm.stats()
        gemm_ms.append(sd.ms_main_kernel); bm25_ms.append(sb.ms_main_kernel); fallbacks.append(sd.n_guard_fallbacks)

    for _ in range(a.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = C.c_uint64()
    L.trr_ctx_launch_count(ctx.h, C.byref(launches0))
    ms_dev = timed(step_device, a.steps)
    launches1 = C.c_uint64()
    L.trr_ctx_launch_count(ctx.h, C.byref(launches1))
    clocks = sampler.stop() if rank == 0 else None
    d_out_snapshot = [t.clone() for t in d_out]       # results of the last timed (device-resident) step
    # per-kernel device times: the library records CUDA events around its dominant kernels on the launching stream at every
    # step; what is read here (the stream is idle after the closing barrier) are the events of the LAST TIMED step
    sd_t, sb_t = dense.stats(), bm.stats()
    gemm_ms_timed, bm25_ms_timed = sd_t.ms_main_kernel, sb_t.ms_main_kernel
    # plus the mean over a few more (untimed) steps, each read back after its own synchronisation
    for _ in range(min(a.steps, 3)):
        step_device()
        collect_stats(None)
    # e2e (headline): the blocking host-buffer C-ABI calls, exactly what a host-language shim would call per batch
    for _ in range(max(1, a.warmup // 2)):
        step_host_buffers()
    last = []
    ms_e2e = timed(step_host_buffers, a.steps, collect=lambda r: last.append(r))
    # secondary: the same work with asynchronous pinned copies around the device-resident step (a serving loop)
    for _ in range(max(1, a.warmup // 2)):
        step_e2e()
    ms_e2e_async = timed(step_e2e, a.steps)
    e2e_out = [t.numpy().copy() for t in out_pin]     # what the last asynchronous end-to-end step delivered to the host
    launches = torch.tensor([launches1.value - launches0.value], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(launches)
    # device time of the merge + fusion kernel (K4), one un-pipelined step at a time on the compute stream.
    # Informational and single-GPU only: an extra collective here could hang a sharded run if one rank failed, and the
    # sharded runs overlap the all-gather with the next batch anyway (profiles/r01_scale.txt compares both schedules).
    exchange = None
    try:
        if world != 1:
            raise RuntimeError("measured on single-GPU runs only")
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t_gather = t_merge = 0.0
        reps = 5
        barrier()
        for _ in range(reps):
            api._check(L.trr_hybrid_local_device(dense.h, bm.h, C.c_void_p(d_q.data_ptr()), C.c_void_p(d_terms.data_ptr()),
                                                 C.c_void_p(d_off.data_ptr()), q_off.ctypes.data_as(u32p), B, Cn, 1, 1,
                                                 C.c_void_p(d_rec.data_ptr())))
            evs[0].record()
            g = gather()
            evs[1].record()
            api._check(L.trr_hybrid_merge_device(ctx.h, C.c_void_p(g.data_ptr()), world, B, Cn, api.RRF, 60.0, K,
                                                 *[C.c_void_p(t.data_ptr()) for t in d_out]))
            evs[2].record()
            torch.cuda.synchronize()
            t_gather += evs[0].elapsed_time(evs[1])
            t_merge += evs[1].elapsed_time(evs[2])
        exchange = {"all_gather_us": round(1e3 * t_gather / reps, 1), "merge_fuse_us": round(1e3 * t_merge / reps, 1),
                    "bytes_per_rank": int(d_rec.numel() * d_rec.element_size()),
                    "note": "un-pipelined, CUDA events on the compute stream; the timed steps overlap both with the next batch"}
        barrier()
    except Exception as ex:  # noqa: BLE001
        exchange = {"skipped": str(ex)[:120]}

    # ---- correctness spot check against the oracle (outside every timed region)
    verify = None
    if rank == 0 and a.verify > 0:
        verify = verify_full_size(a, api, dense, bm, q_np, q_terms[:nt], q_off, host_csr, last[-1], world)
        # the device-resident (and, when sharded, pipelined) step and the end-to-end step must have produced exactly what the
        # blocking host-buffer calls return
        e_ord, e_f, e_d, e_s, e_n = last[-1]
        for name, got in (("device_step_equals_host_buffer_call", [t.cpu().numpy() for t in d_out_snapshot]),
                          ("e2e_step_equals_host_buffer_call", e2e_out)):
            same = np.array_equal(got[4].view(np.uint32), e_n)
            for b in range(B):
                m = int(e_n[b])
                same = same and np.array_equal(got[0][b, :m].view(np.uint32), e_ord[b, :m]) and np.array_equal(got[1][b, :m], e_f[b, :m])
            verify[name] = bool(same)
            verify["consistent"] = bool(verify.get("consistent", True) and same)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained")
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)"
        if not peak_tf:
            peak_tf, peak_src = 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"
        peak_hbm = peaks.get("hbm_gbs") or 6650.0
        g_ms, b_ms = float(gemm_ms_timed), float(bm25_ms_timed)
        flops = 2.0 * B * n_loc * D
        tf = flops / g_ms / 1e9
        bm_gbs = 8.0 * postings_per_batch_local / b_ms / 1e6
        step_ms = ms_dev / a.steps
        e2e_ms = ms_e2e / a.steps
        line = {
            "metric": METRIC, "value": B / step_ms * 1e3, "unit": "queries/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": config(a, {"setup_s": round(setup_s, 1), "docs_per_gpu": n_loc}),
            "e2e": {"value": B / e2e_ms * 1e3, "unit": "queries/s", "ms_per_step": e2e_ms,
                    "how": "per step: the blocking host-buffer C-ABI calls trr_hybrid_local + (all-gather) + trr_hybrid_merge: H2D of queries / terms / offsets, kernels, D2H of the results, synchronised",
                    "async_pinned": {"value": B / (ms_e2e_async / a.steps) * 1e3, "unit": "queries/s",
                                     "how": "same bytes per step with asynchronous copies from / to pinned memory around the device-resident step (copies overlap neighbouring steps); results verified against the blocking calls"},
                    "h2d_bytes_per_step": int(B * D * 4 + nt * 4 + (B + 1) * 4), "d2h_bytes_per_step": int(B * K * 16 + B * 4)},
            "gpu_launches": int(launches.item()),
            "roofline": {"kernel": ("dense_gemm_topk_kernel (tcgen05 bf16 GEMM + fused top-k), rank 0 shard" if os.environ.get("TRR_GEMM_PAIR") == "0" or a.batch <= 128 else "dense_gemm_topk_pair_kernel (tcgen05 cta_group::2 bf16 GEMM + fused top-k), rank 0 shard"),
                         "bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                         "traffic": ncu_traffic("gemm", a), "traffic_source": "profiles/r*_gemm_ncu.txt (ncu --set full, same command)",
                         "peak_source": peak_src,
                         "algorithmic": f"2*B*N_shard*D = {flops:.4g} flop per launch / {g_ms:.3f} ms"},
            "kernels": {"gemm_ms": g_ms, "bm25_ms": b_ms, "timing": "CUDA events of the last timed step",
                        "gemm_ms_mean_of_extra_steps": float(np.mean(gemm_ms)), "bm25_ms_mean_of_extra_steps": float(np.mean(bm25_ms)),
                        "bm25": {"bound": "hbm", "achieved": bm_gbs, "peak": peak_hbm, "unit": "GB/s", "frac": bm_gbs / peak_hbm,
                                 "algorithmic": f"8 B x {postings_per_batch_local} postings per launch"},
                        "guard_fallbacks_per_batch": float(np.mean(fallbacks)), "exchange": exchange},
            "clocks": clocks, "verify": verify,
        }
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(a)
        print(json.dumps(line), flush=True)
    torch.cuda.synchronize()
    dense.close(); bm.close(); ctx.close()
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic(kind, a):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed `ncu --set full` capture of
    this same command (profiles/rNN_<kind>_ncu.txt, written by tools/make_profiles.py); None when the capture was taken
    at another size."""
    import glob
    import re
    if (a.docs, a.dim, a.batch, a.gpus) != (10_000_000, 768, 1024, 1):
        return None
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_{kind}_ncu.txt")))
    if not files:
        return None
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tot = 0.0
    for line in open(files[-1]):
        m = re.match(r"\s*dram__bytes_(read|write)\.sum \[(\w+)\] = ([0-9.eE+-]+)", line)
        if m:
            tot += float(m.group(3)) * mult.get(m.group(2), 1.0)
    return tot or None


def verify_full_size(a, api, dense, bm, q_np, q_terms, q_off, host_csr, outputs, world):
    """Parity at the FULL bench size, outside every timed region, on the first `--verify` queries:
      dense   the tensor-core path (K2 + exact re-scoring + proof) must return bit-identical ids and scores to the exact scan
              kernel K1 over the same 10M-row slab; K1's returned scores are recomputed by the CPU oracle from the stored
              rows (bit-exact), its list is in canonical order, and no row of a 20000-row random sample beats its C-th hit;
      sparse  the BM25 kernel vs the CPU oracle scoring the same host CSR with the same global statistics (single GPU only);
      fused   the e2e output vs the oracle's RRF fusion + take(k) of those two lists.
    With several GPUs only the internal-consistency properties are checked (ids unique, fused scores sorted and in range)."""
    from oracle import oracle as O
    o_ord, o_f, o_d, o_s, o_n = outputs
    n = int(min(a.verify, o_ord.shape[0]))
    res = {"checked_queries": n, "ids_unique_sorted_in_range": True}
    for b in range(n):
        m = int(o_n[b])
        ok = len(set(o_ord[b, :m].tolist())) == m and bool(np.all(np.diff(o_f[b, :m]) <= 0))
        ok &= bool(np.all((o_f[b, :m] > 0) & (o_f[b, :m] <= np.float32(2.0 / 61.0) + 1e-7)))
        res["ids_unique_sorted_in_range"] &= bool(ok)
    if world != 1:
        res["note"] = "multi-GPU run: full-size oracle parity is checked by the 1-GPU run and by tests/ (logical shards)"
        return res
    Cn, K = a.cands, a.k
    dense.set_mode(api.MODE_SCAN)
    s_ord, s_sc, s_n = dense.search(q_np[:n], Cn)
    dense.set_mode(api.MODE_GEMM)
    g_ord, g_sc, g_n = dense.search(q_np[:n], Cn)
    res["dense_gemm_equals_exact_scan"] = bool(np.array_equal(s_n, g_n) and np.array_equal(s_ord, g_ord) and
                                                np.array_equal(s_sc, g_sc))
    # the exact scan itself against the CPU oracle at full size: every returned score is recomputed by the oracle from
    # the stored rows (downloaded), and a random sample of 20000 other rows must not beat the C-th result
    rng = np.random.default_rng(1)
    ok = True
    for b in range(n):
        m = int(s_n[b])
        rows = dense.rows(s_ord[b, :m])
        f = (rows.astype(np.uint32) << 16).view(np.float32) if rows.dtype == np.uint16 else rows
        exp = np.array([O.cosine(q_np[b], f[i]) for i in range(m)], np.float32)
        ok = ok and np.array_equal(exp, s_sc[b, :m])
        ok = ok and bool(np.all((s_sc[b, :m - 1] > s_sc[b, 1:m]) | ((s_sc[b, :m - 1] == s_sc[b, 1:m]) & (s_ord[b, :m - 1] < s_ord[b, 1:m]))))
        samp = rng.integers(0, a.docs, 20000).astype(np.uint32)
        rs = dense.rows(samp)
        fs = (rs.astype(np.uint32) << 16).view(np.float32) if rs.dtype == np.uint16 else rs
        o_ids, o_sc, o_cnt = O.dense_search_batch(fs, q_np[b:b + 1], 1)
        best_ord, best_sc = int(samp[int(o_ids[0, 0])]), float(o_sc[0, 0])
        kth_sc, kth_ord = float(s_sc[b, m - 1]), int(s_ord[b, m - 1])
        in_list = best_ord in set(s_ord[b, :m].tolist())
        ok = ok and (in_list or best_sc < kth_sc or (best_sc == kth_sc and best_ord > kth_ord))
    res["exact_scan_scores_equal_oracle_and_dominate_sample"] = bool(ok)
    term_off, post_doc, post_tf, doc_len, df_g, avgdl = host_csr
    oix = O.BM25.from_csr(len(doc_len), a.vocab, term_off, post_doc, post_tf, doc_len, df_g, avgdl)
    qt, qo = q_terms[:int(q_off[n])], q_off[:n + 1]
    b_ord, b_sc, b_n = bm.search(qt, qo, Cn)
    e_ord, e_sc, e_n = oix.search_batch(qt, qo, Cn)
    ok = np.array_equal(b_n, e_n)
    for b in range(n):
        m = int(e_n[b])
        ok = ok and np.array_equal(b_ord[b, :m], e_ord[b, :m]) and np.array_equal(b_sc[b, :m], e_sc[b, :m])
    res["bm25_equals_oracle"] = bool(ok)
    ok = True
    for b in range(n):
        i, f, dd, ss = O.hybrid_assemble(O.RRF, 60.0, (s_ord[b, :s_n[b]], s_sc[b, :s_n[b]]),
                                         (e_ord[b, :e_n[b]], e_sc[b, :e_n[b]]), K)
        m = int(o_n[b])
        ok = ok and m == len(i) and np.array_equal(o_ord[b, :m], i) and np.array_equal(o_f[b, :m], f) and \
            np.array_equal(o_d[b, :m], dd, equal_nan=True) and np.array_equal(o_s[b, :m], ss, equal_nan=True)
    res["fused_equals_oracle"] = bool(ok)
    res["consistent"] = bool(res["ids_unique_sorted_in_range"] and res["dense_gemm_equals_exact_scan"] and
                             res["exact_scan_scores_equal_oracle_and_dominate_sample"] and
                             res["bm25_equals_oracle"] and res["fused_equals_oracle"])
    return res


if __name__ == "__main__":
    # stdout carries exactly one JSON line: everything else that writes to fd 1 (e.g. NCCL's version banner) goes to stderr
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _real_stdout
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
