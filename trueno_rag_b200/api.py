"""Python glue over the C ABI (tests and bench.py use it; the product is the shared library).

Two layers, both thin:
  * `Context`, `DenseIndex`, `Bm25Device`, `fuse`, `hybrid_search`, `hybrid_local/merge` bind include/trueno_rag_b200.h
    (ordinals and term ids, numpy arrays in and out) — what a Rust `-sys` crate would bind;
  * `VectorStore`, `BM25Index`, `FusionStrategy`, `HybridRetriever`, `Chunk`, `ChunkId` bind the C++ host mirror
    (include/trueno_rag.hpp through trueno_rag_host.h) and carry the reference's names and error behaviour.
No numeric work happens in Python.
"""
from __future__ import annotations

import ctypes as C
import uuid
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import HostId, Stats, f32p, u16p, u32p, u64p

COSINE, EUCLIDEAN, DOT = 0, 1, 2
F32, BF16 = 0, 1
RRF, LINEAR, CONVEX, DBSF, UNION, INTERSECTION = 0, 1, 2, 3, 4, 5
MODE_AUTO, MODE_SCAN, MODE_GEMM = 0, 1, 2


class TrrError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"[trr status {status}] {msg}")
        self.status = status


def _check(status: int):
    if status != 0:
        raise TrrError(status, _lib.load().trr_last_error().decode("utf-8", "replace"))


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def device_count() -> int:
    return int(_lib.load().trr_device_count())


class Context:
    def __init__(self, device: int = 0):
        self.L = _lib.load()
        self.h = C.c_void_p()
        _check(self.L.trr_ctx_create(device, C.byref(self.h)))
        self.device = device

    def close(self):
        if self.h:
            self.L.trr_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def sync(self):
        _check(self.L.trr_ctx_sync(self.h))

    @property
    def stream(self) -> int:
        s = C.c_void_p()
        _check(self.L.trr_ctx_stream(self.h, C.byref(s)))
        return int(s.value or 0)

    @property
    def sm_count(self) -> int:
        n = C.c_int()
        _check(self.L.trr_ctx_sm_count(self.h, C.byref(n)))
        return n.value

    def flush_l2(self, nbytes: int = 256 << 20):
        _check(self.L.trr_ctx_flush_l2(self.h, nbytes))


class DenseIndex:
    """trr_dense: the device side of VectorStore (reference src/index.rs:322-437)."""

    def __init__(self, ctx: Context, dim: int, metric: int = COSINE, dtype: int = F32, capacity: int = 0, base: int = 0):
        self.ctx, self.L, self.dim, self.metric, self.dtype = ctx, ctx.L, dim, metric, dtype
        self.h = C.c_void_p()
        _check(self.L.trr_dense_create(ctx.h, dim, metric, dtype, capacity, C.byref(self.h)))
        if base:
            _check(self.L.trr_dense_set_base(self.h, base))

    def close(self):
        if self.h:
            self.L.trr_dense_destroy(self.h)
            self.h = C.c_void_p()

    def save(self, path: str):
        """Snapshot of the device store (slab + tombstones + configuration) in a flat file."""
        _check(self.L.trr_dense_save(self.h, str(path).encode()))

    @classmethod
    def load(cls, ctx: Context, path: str) -> "DenseIndex":
        self = cls.__new__(cls)
        self.ctx, self.L = ctx, ctx.L
        self.h = C.c_void_p()
        _check(self.L.trr_dense_load(ctx.h, str(path).encode(), C.byref(self.h)))
        dim, metric, dtype, base = C.c_uint32(), C.c_int(), C.c_int(), C.c_uint32()
        _check(self.L.trr_dense_info(self.h, C.byref(dim), C.byref(metric), C.byref(dtype), C.byref(base)))
        self.dim, self.metric, self.dtype = dim.value, metric.value, dtype.value
        return self

    def append(self, rows: np.ndarray):
        if rows.dtype == np.uint16:
            rows = np.ascontiguousarray(rows)
            assert rows.ndim == 2 and rows.shape[1] == self.dim
            _check(self.L.trr_dense_append_bf16(self.h, _p(rows, u16p), rows.shape[0]))
        else:
            rows = np.ascontiguousarray(rows, dtype=np.float32)
            assert rows.ndim == 2 and rows.shape[1] == self.dim
            _check(self.L.trr_dense_append(self.h, _p(rows, f32p), rows.shape[0]))

    def append_device(self, dev_ptr: int, n: int):
        _check(self.L.trr_dense_append_device(self.h, C.c_void_p(dev_ptr), n))

    def append_synth(self, seed: int, first_row: int, n: int, dups: bool = False):
        _check(self.L.trr_dense_append_synth(self.h, seed, first_row, n, int(dups)))

    def remove(self, ordinal: int):
        _check(self.L.trr_dense_remove(self.h, ordinal))

    def freeze(self):
        _check(self.L.trr_dense_freeze(self.h))

    def set_mode(self, mode: int):
        _check(self.L.trr_dense_set_mode(self.h, mode))

    def __len__(self):
        n = C.c_uint64()
        _check(self.L.trr_dense_len(self.h, C.byref(n)))
        return n.value

    def search(self, q: np.ndarray, k: int):
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.shape[1] != self.dim:  # VectorStore::search dimension check (src/index.rs:387-392) lives in the host layer
            raise TrrError(_lib.TRR_ERR_DIM_MISMATCH, f"expected {self.dim}, got {q.shape[1]}")
        B = q.shape[0]
        kk = max(k, 1)
        ords = np.full((B, kk), 0xFFFFFFFF, np.uint32)
        scores = np.zeros((B, kk), np.float32)
        n = np.zeros(B, np.uint32)
        _check(self.L.trr_dense_search(self.h, _p(q, f32p), B, k, _p(ords, u32p), _p(scores, f32p), _p(n, u32p)))
        return ords, scores, n

    def search_device(self, d_q: int, B: int, k: int, d_ord: int, d_score: int, d_n: int):
        _check(self.L.trr_dense_search_device(self.h, C.c_void_p(d_q), B, k, C.c_void_p(d_ord), C.c_void_p(d_score),
                                              C.c_void_p(d_n)))

    def stats(self) -> Stats:
        s = Stats()
        _check(self.L.trr_dense_last_stats(self.h, C.byref(s)))
        return s

    def rows(self, ordinals) -> np.ndarray:
        """Stored rows (local ordinals) in the store's dtype: float32, or uint16 bf16 bit patterns."""
        o = np.ascontiguousarray(ordinals, dtype=np.uint32)
        out = np.zeros((o.size, self.dim), np.uint16 if self.dtype == BF16 else np.float32)
        if o.size:
            _check(self.L.trr_dense_copy_rows(self.h, _p(o, u32p), o.size, C.c_void_p(out.ctypes.data)))
        return out

    def norms(self, n: int) -> np.ndarray:
        out = np.zeros(n, np.float32)
        _check(self.L.trr_dense_copy_norms(self.h, _p(out, f32p), n))
        return out

    def debug_gemm_scores(self, q: np.ndarray, n_pad: int) -> np.ndarray:
        q = np.ascontiguousarray(q, dtype=np.float32)
        out = np.zeros((q.shape[0], n_pad), np.float32)
        _check(self.L.trr_debug_gemm_scores(self.h, _p(q, f32p), q.shape[0], _p(out, f32p), n_pad))
        return out


def bm25_idf_host(n_docs: int, df: np.ndarray) -> np.ndarray:
    """idf table with the platform logf (reference src/index.rs:147), as the host language would compute it."""
    libm = C.CDLL("libm.so.6")
    libm.logf.restype = C.c_float
    libm.logf.argtypes = [C.c_float]
    n = np.float32(n_docs)
    dff = df.astype(np.float32)
    x = (n - dff + np.float32(0.5)) / (dff + np.float32(0.5)) + np.float32(1.0)
    x = x.astype(np.float32)
    uniq, inv = np.unique(x, return_inverse=True)
    vals = np.array([libm.logf(float(v)) for v in uniq], dtype=np.float32)
    return vals[inv].astype(np.float32)


class Bm25Device:
    """trr_bm25: the device side of BM25Index (reference src/index.rs:30-280)."""

    def __init__(self, ctx: Context, n_docs: int, term_off, post_doc, post_tf, doc_len, avgdl: float, idf,
                 k1: float = 1.2, b: float = 0.75, doc_base: int = 0):
        self.ctx, self.L = ctx, ctx.L
        term_off = np.ascontiguousarray(term_off, dtype=np.uint64)
        self.n_terms = len(term_off) - 1
        post_doc = np.ascontiguousarray(post_doc, dtype=np.uint32)
        post_tf = np.ascontiguousarray(post_tf, dtype=np.uint32)
        doc_len = np.ascontiguousarray(doc_len, dtype=np.uint32)
        idf = np.ascontiguousarray(idf, dtype=np.float32)
        self.h = C.c_void_p()
        _check(self.L.trr_bm25_build(ctx.h, n_docs, self.n_terms, _p(term_off, u64p), _p(post_doc, u32p),
                                     _p(post_tf, u32p), _p(doc_len, u32p), avgdl, k1, b, _p(idf, f32p), doc_base,
                                     C.byref(self.h)))

    def close(self):
        if self.h:
            self.L.trr_bm25_destroy(self.h)
            self.h = C.c_void_p()

    def append(self, n_new_docs: int, term_off, post_doc, post_tf, doc_len, avgdl: float, idf, k1: float = 1.2, b: float = 0.75):
        """Appends documents (CSR with doc ids relative to the first new document); idf / avgdl are the new global statistics."""
        term_off = np.ascontiguousarray(term_off, dtype=np.uint64)
        post_doc = np.ascontiguousarray(post_doc, dtype=np.uint32)
        post_tf = np.ascontiguousarray(post_tf, dtype=np.uint32)
        doc_len = np.ascontiguousarray(doc_len, dtype=np.uint32)
        idf = np.ascontiguousarray(idf, dtype=np.float32)
        pd = post_doc if post_doc.size else np.zeros(1, np.uint32)
        pt = post_tf if post_tf.size else np.zeros(1, np.uint32)
        dl = doc_len if doc_len.size else np.zeros(1, np.uint32)
        _check(self.L.trr_bm25_append(self.h, n_new_docs, len(term_off) - 1, _p(term_off, u64p), _p(pd, u32p), _p(pt, u32p),
                                      _p(dl, u32p), avgdl, k1, b, _p(idf, f32p)))
        self.n_terms = len(term_off) - 1

    def remove(self, ordinals, avgdl: float, idf, k1: float = 1.2, b: float = 0.75) -> int:
        """Removes documents (local ordinals) in place; idf / avgdl are the new global statistics.  Returns the number of
        dead postings the index still holds."""
        o = np.ascontiguousarray(ordinals, dtype=np.uint32)
        idf = np.ascontiguousarray(idf, dtype=np.float32)
        dead = C.c_uint64()
        oo = o if o.size else np.zeros(1, np.uint32)
        _check(self.L.trr_bm25_remove(self.h, _p(oo, u32p), o.size, avgdl, k1, b, _p(idf, f32p), C.byref(dead)))
        return dead.value

    def save(self, path: str):
        """Snapshot of the device index (postings with impacts, skip table, per-term minimum impacts)."""
        _check(self.L.trr_bm25_save(self.h, str(path).encode()))

    @classmethod
    def load(cls, ctx: Context, path: str) -> "Bm25Device":
        self = cls.__new__(cls)
        self.ctx, self.L = ctx, ctx.L
        self.h = C.c_void_p()
        _check(self.L.trr_bm25_load(ctx.h, str(path).encode(), C.byref(self.h)))
        self.n_terms = None
        return self

    @property
    def n_postings(self) -> int:
        n = C.c_uint64()
        _check(self.L.trr_bm25_n_postings(self.h, C.byref(n)))
        return n.value

    def search(self, q_terms, q_off, k: int):
        q_terms = np.ascontiguousarray(q_terms, dtype=np.uint32)
        q_off = np.ascontiguousarray(q_off, dtype=np.uint32)
        B = len(q_off) - 1
        kk = max(k, 1)
        ords = np.full((B, kk), 0xFFFFFFFF, np.uint32)
        scores = np.zeros((B, kk), np.float32)
        n = np.zeros(B, np.uint32)
        qt = q_terms if q_terms.size else np.zeros(1, np.uint32)
        _check(self.L.trr_bm25_search(self.h, _p(qt, u32p), _p(q_off, u32p), B, k, _p(ords, u32p), _p(scores, f32p),
                                      _p(n, u32p)))
        return ords, scores, n

    def stats(self) -> Stats:
        s = Stats()
        _check(self.L.trr_bm25_last_stats(self.h, C.byref(s)))
        return s

    def impacts(self) -> np.ndarray:
        out = np.zeros(max(self.n_postings, 1), np.float32)
        _check(self.L.trr_bm25_copy_impacts(self.h, _p(out, f32p), self.n_postings))
        return out[:self.n_postings]


def _pad_lists(lists, C_):
    B = len(lists)
    ords = np.zeros((B, C_), np.uint32)
    sc = np.zeros((B, C_), np.float32)
    n = np.zeros(B, np.uint32)
    for b, (o, s) in enumerate(lists):
        n[b] = len(o)
        ords[b, :len(o)] = o
        sc[b, :len(o)] = s
    return ords, sc, n


def fuse(ctx: Context, strategy: int, param: float, dense_lists, sparse_lists, k_out: Optional[int] = None):
    """dense_lists / sparse_lists: per query (ordinals, scores).  Returns per query (ord, fused, dense, sparse)."""
    B = len(dense_lists)
    C_ = max([1] + [len(o) for o, _ in dense_lists] + [len(o) for o, _ in sparse_lists])
    d_ord, d_sc, d_n = _pad_lists(dense_lists, C_)
    s_ord, s_sc, s_n = _pad_lists(sparse_lists, C_)
    k_out = k_out or 2 * C_
    o_ord = np.zeros((B, k_out), np.uint32)
    o_f, o_d, o_s = (np.zeros((B, k_out), np.float32) for _ in range(3))
    o_n = np.zeros(B, np.uint32)
    _check(ctx.L.trr_fuse(ctx.h, strategy, param, _p(d_ord, u32p), _p(d_sc, f32p), _p(d_n, u32p), _p(s_ord, u32p),
                          _p(s_sc, f32p), _p(s_n, u32p), B, C_, k_out, _p(o_ord, u32p), _p(o_f, f32p), _p(o_d, f32p),
                          _p(o_s, f32p), _p(o_n, u32p)))
    return [(o_ord[b, :o_n[b]].copy(), o_f[b, :o_n[b]].copy(), o_d[b, :o_n[b]].copy(), o_s[b, :o_n[b]].copy())
            for b in range(B)]


def _hybrid_inputs(dense, q, q_terms, q_off, use_dense, use_sparse):
    """Shapes of a hybrid batch as the C ABI expects them; the dimension check of VectorStore::search
    (reference src/index.rs:387-392) lives in this host layer, as in DenseIndex.search."""
    q = np.ascontiguousarray(q, dtype=np.float32) if q is not None else None
    if q is not None and q.ndim == 1:
        q = q[None, :]
    if q is not None and use_dense and dense is not None and q.shape[1] != dense.dim:
        raise TrrError(_lib.TRR_ERR_DIM_MISMATCH, f"expected {dense.dim}, got {q.shape[1]}")
    q_off = np.ascontiguousarray(q_off, dtype=np.uint32) if q_off is not None else None
    q_terms = np.ascontiguousarray(q_terms, dtype=np.uint32) if q_terms is not None else None
    B = q.shape[0] if q is not None else len(q_off) - 1
    if q_off is not None and use_sparse:
        if len(q_off) != B + 1:
            raise TrrError(_lib.TRR_ERR_INVALID_ARG, f"q_off must have B + 1 = {B + 1} entries, got {len(q_off)}")
        if int(q_off[-1]) > (q_terms.size if q_terms is not None else 0):
            raise TrrError(_lib.TRR_ERR_INVALID_ARG, "q_off points past the end of q_terms")
    qt = q_terms if (q_terms is not None and q_terms.size) else np.zeros(1, np.uint32)
    return q, qt, q_off, B


def _hybrid_outputs(B, k):
    return (np.full((B, k), 0xFFFFFFFF, np.uint32), np.zeros((B, k), np.float32), np.zeros((B, k), np.float32),
            np.zeros((B, k), np.float32), np.zeros(B, np.uint32))


def hybrid_search(dense: Optional[DenseIndex], bm25: Optional[Bm25Device], q, q_terms, q_off, C_: int, strategy: int,
                  param: float, k: int, use_dense: bool = True, use_sparse: bool = True):
    """HybridRetriever::retrieve for B queries (reference src/retrieve.rs:175-220) in one C-ABI call."""
    L = (dense or bm25).L
    q, qt, q_off, B = _hybrid_inputs(dense, q, q_terms, q_off, use_dense, use_sparse)
    o_ord, o_f, o_d, o_s, o_n = _hybrid_outputs(B, k)
    _check(L.trr_hybrid_search(dense.h if dense else None, bm25.h if bm25 else None, _p(q, f32p), _p(qt, u32p),
                               _p(q_off, u32p), B, C_, strategy, param, k, int(use_dense), int(use_sparse),
                               _p(o_ord, u32p), _p(o_f, f32p), _p(o_d, f32p), _p(o_s, f32p), _p(o_n, u32p)))
    return o_ord, o_f, o_d, o_s, o_n


def exchange_bytes(B: int, C_: int) -> int:
    return int(_lib.load().trr_exchange_bytes(B, C_))


def hybrid_local(dense, bm25, q, q_terms, q_off, C_: int, d_exchange: int, use_dense=True, use_sparse=True):
    L = (dense or bm25).L
    q, qt, q_off, B = _hybrid_inputs(dense, q, q_terms, q_off, use_dense, use_sparse)
    _check(L.trr_hybrid_local(dense.h if dense else None, bm25.h if bm25 else None, _p(q, f32p), _p(qt, u32p),
                              _p(q_off, u32p), B, C_, int(use_dense), int(use_sparse), C.c_void_p(d_exchange)))


def hybrid_merge(ctx: Context, d_gathered: int, G: int, B: int, C_: int, strategy: int, param: float, k: int):
    o_ord, o_f, o_d, o_s, o_n = _hybrid_outputs(B, k)
    _check(ctx.L.trr_hybrid_merge(ctx.h, C.c_void_p(d_gathered), G, B, C_, strategy, param, k, _p(o_ord, u32p),
                                  _p(o_f, f32p), _p(o_d, f32p), _p(o_s, f32p), _p(o_n, u32p)))
    return o_ord, o_f, o_d, o_s, o_n


def rank_by_cosine(ctx: Context, query, embeddings):
    """The brute-force ranking of examples/nemotron_embeddings.rs:79-92 (cosine of the query against every document
    embedding, stable sort by similarity descending) through the exact scan kernel: returns (indices, similarities) of
    all documents, best first; equal similarities keep document order (the stable sort of the reference)."""
    emb = np.ascontiguousarray(embeddings, dtype=np.float32)
    if emb.ndim != 2 or emb.shape[0] == 0:
        return np.zeros(0, np.uint32), np.zeros(0, np.float32)
    if emb.shape[0] > 1024:
        raise TrrError(_lib.TRR_ERR_UNSUPPORTED, "rank_by_cosine ranks at most 1024 documents per call (k <= 1024)")
    ix = DenseIndex(ctx, emb.shape[1], COSINE, F32)
    try:
        ix.append(emb)
        ix.set_mode(MODE_SCAN)
        ords, scores, n = ix.search(np.ascontiguousarray(query, dtype=np.float32), emb.shape[0])
        return ords[0, :int(n[0])].copy(), scores[0, :int(n[0])].copy()
    finally:
        ix.close()


# ---------------------------------------------------------------------------------------------------
# sharded search: one process per GPU, the exchange inside the call
# ---------------------------------------------------------------------------------------------------
EXCHANGE_NCCL, EXCHANGE_PEER = 0, 1


def group_unique_id() -> bytes:
    """Rank 0: the 128-byte rendezvous id the other ranks need for Group(...)."""
    buf = C.create_string_buffer(128)
    _check(_lib.load().trr_group_unique_id(buf))
    return buf.raw


class Group:
    """The ranks (one process per GPU) that hold the shards of one corpus.  Owns the communicator and the exchange
    buffers; `search` is HybridRetriever::retrieve over the whole corpus in ONE call per rank
    (reference src/retrieve.rs:175-220)."""

    def __init__(self, ctx: Context, rank: int, world: int, unique_id: Optional[bytes] = None, exchange: int = EXCHANGE_NCCL,
                 max_record_bytes: int = 0):
        self.L, self.ctx, self.rank, self.world = ctx.L, ctx, rank, world
        self.h = C.c_void_p()
        idbuf = C.create_string_buffer(unique_id, 128) if unique_id else None
        _check(self.L.trr_group_create(ctx.h, idbuf, rank, world, exchange, max_record_bytes, C.byref(self.h)))

    def close(self):
        if self.h:
            self.L.trr_group_destroy(self.h)
            self.h = C.c_void_p()

    @property
    def exchange(self) -> int:
        e = C.c_int()
        _check(self.L.trr_group_info(self.h, None, None, C.byref(e)))
        return e.value

    def sync(self):
        _check(self.L.trr_group_sync(self.h))

    def allreduce_u64(self, values: np.ndarray, op_max: bool = False) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.uint64).copy()
        _check(self.L.trr_group_allreduce_u64(self.h, _p(v, u64p), v.size, int(op_max)))
        return v

    def search(self, dense, bm25, q, q_terms, q_off, C_: int, strategy: int, param: float, k: int, use_dense=True,
               use_sparse=True):
        q, qt, q_off, B = _hybrid_inputs(dense, q, q_terms, q_off, use_dense, use_sparse)
        o_ord, o_f, o_d, o_s, o_n = _hybrid_outputs(B, k)
        _check(self.L.trr_hybrid_search_sharded(self.h, dense.h if dense else None, bm25.h if bm25 else None, _p(q, f32p),
                                                _p(qt, u32p), _p(q_off, u32p), B, C_, strategy, param, k, int(use_dense),
                                                int(use_sparse), _p(o_ord, u32p), _p(o_f, f32p), _p(o_d, f32p), _p(o_s, f32p),
                                                _p(o_n, u32p)))
        return o_ord, o_f, o_d, o_s, o_n

    def search_async(self, dense, bm25, q_ptr: int, terms_ptr: int, off_ptr: int, B: int, C_: int, strategy: int, param: float,
                     k: int, out_ptrs, use_dense=True, use_sparse=True):
        """The host-buffer call without the final wait: raw addresses of (page-locked) host buffers, valid until sync()."""
        _check(self.L.trr_hybrid_search_sharded_async(self.h, dense.h if dense else None, bm25.h if bm25 else None,
                                                      C.c_void_p(q_ptr), C.c_void_p(terms_ptr), C.c_void_p(off_ptr), B, C_,
                                                      strategy, param, k, int(use_dense), int(use_sparse),
                                                      *[C.c_void_p(p) for p in out_ptrs]))

    def step_device(self, dense, bm25, d_q: int, d_terms: int, d_off: int, h_q_off: np.ndarray, B: int, C_: int,
                    strategy: int, param: float, k: int, d_out, use_dense=True, use_sparse=True):
        """Enqueues one sharded step on device buffers (d_out: five device pointers ord, fused, dense, sparse, n)."""
        _check(self.L.trr_hybrid_search_sharded_device(self.h, dense.h if dense else None, bm25.h if bm25 else None,
                                                       C.c_void_p(d_q), C.c_void_p(d_terms), C.c_void_p(d_off),
                                                       _p(h_q_off, u32p), B, C_, strategy, param, k, int(use_dense),
                                                       int(use_sparse), *[C.c_void_p(p) for p in d_out]))


# =====================================================================================================
# host mirror (reference names)
# =====================================================================================================
class Error(Exception):
    """Mirror of the reference's `Error` enum (src/error.rs:9-64) for the variants reachable from the path."""

    def __init__(self, kind: str, msg: str, expected: int = 0, actual: int = 0):
        super().__init__(f"{kind}: {msg}")
        self.kind, self.expected, self.actual = kind, expected, actual


_KINDS = {1: "InvalidConfig", 2: "DimensionMismatch", 3: "VectorStore", 6: "Unsupported", 7: "SerializationError"}


def _hcheck(status: int):
    if status != 0:
        L = _lib.load()
        raise Error(_KINDS.get(status, "VectorStore"), L.trrh_last_error().decode("utf-8", "replace"),
                    int(L.trrh_last_expected()), int(L.trrh_last_actual()))


@dataclass(frozen=True)
class ChunkId:
    value: uuid.UUID = field(default_factory=uuid.uuid4)

    @staticmethod
    def from_u128(n: int) -> "ChunkId":
        return ChunkId(uuid.UUID(int=n))

    def _c(self) -> HostId:
        return HostId(self.value.int >> 64, self.value.int & 0xFFFFFFFFFFFFFFFF)

    @staticmethod
    def _from_c(h: HostId) -> "ChunkId":
        return ChunkId(uuid.UUID(int=(int(h.hi) << 64) | int(h.lo)))


@dataclass
class Chunk:
    content: str
    id: ChunkId = field(default_factory=ChunkId)
    embedding: Optional[Sequence[float]] = None

    def set_embedding(self, e):
        self.embedding = e


class DistanceMetric:
    Cosine, Euclidean, DotProduct = COSINE, EUCLIDEAN, DOT


def _ids_out(n):
    return (HostId * max(n, 1))()


class VectorStore:
    """Reference `VectorStore` (src/index.rs:322-437) on the device."""

    def __init__(self, dimension: int = 384, metric: int = COSINE, dtype: int = F32, _h=None):
        self.L = _lib.load()
        self.dimension, self.metric = dimension, metric
        self.h = _h or C.c_void_p()
        if _h is None:
            _hcheck(self.L.trrh_store_new(dimension, metric, dtype, C.byref(self.h)))

    @staticmethod
    def with_dimension(dimension: int) -> "VectorStore":
        return VectorStore(dimension)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.trrh_store_free(self.h)
                self.h = None
        except Exception:
            pass

    def _release(self):
        h, self.h = self.h, None
        return h

    def insert(self, chunk: Chunk):
        emb = None if chunk.embedding is None else np.ascontiguousarray(chunk.embedding, dtype=np.float32)
        _hcheck(self.L.trrh_store_insert(self.h, chunk.id._c(), chunk.content.encode(), _p(emb, f32p),
                                         0 if emb is None else emb.size, int(emb is not None)))

    def insert_batch(self, chunks):
        for c in chunks:
            self.insert(c)

    def search(self, query_vector, k: int) -> List[Tuple[ChunkId, float]]:
        q = np.ascontiguousarray(query_vector, dtype=np.float32)
        cap = max(min(k, len(self)), 1)
        ids, sc, n = _ids_out(cap), np.zeros(cap, np.float32), C.c_uint32()
        _hcheck(self.L.trrh_store_search(self.h, _p(q, f32p), q.size, k, ids, _p(sc, f32p), C.byref(n)))
        return [(ChunkId._from_c(ids[i]), float(sc[i])) for i in range(n.value)]

    def get(self, chunk_id: ChunkId) -> Optional[str]:
        out = C.c_char_p()
        if self.L.trrh_store_get(self.h, chunk_id._c(), C.byref(out)):
            return out.value.decode()
        return None

    def remove(self, chunk_id: ChunkId) -> bool:
        r = self.L.trrh_store_remove(self.h, chunk_id._c())
        if r < 0:
            _hcheck(-r)
        return r == 1

    def __len__(self):
        return int(self.L.trrh_store_len(self.h))

    def is_empty(self):
        return len(self) == 0

    def set_mode(self, mode: int):
        _hcheck(self.L.trrh_store_set_mode(self.h, mode))

    def clone(self) -> "VectorStore":
        h = C.c_void_p()
        _hcheck(self.L.trrh_store_clone(self.h, C.byref(h)))
        return VectorStore(self.dimension, self.metric, _h=h)


class Compression:
    """Reference `Compression` (src/compressed.rs:13-31); `Lz4` is the default."""
    Lz4, Zstd = 0, 1

    @staticmethod
    def as_str(c: int) -> str:
        return "lz4" if c == Compression.Lz4 else "zstd"

    @staticmethod
    def default() -> int:
        return Compression.Lz4


def _take_bytes(L, out, n) -> bytes:
    try:
        return C.string_at(out, n.value) if n.value else b""
    finally:
        L.trrh_bytes_free(out)


def compress(data: bytes, compression: int = Compression.Lz4) -> bytes:
    """`Compression::compress` (src/compressed.rs:36-47)."""
    L = _lib.load()
    out, n = C.c_void_p(), C.c_uint64()
    _hcheck(L.trrh_compress(compression, bytes(data), len(data), C.byref(out), C.byref(n)))
    return _take_bytes(L, out, n)


def decompress(data: bytes, compression: int = Compression.Lz4) -> bytes:
    """`Compression::decompress` (src/compressed.rs:53-66)."""
    L = _lib.load()
    out, n = C.c_void_p(), C.c_uint64()
    _hcheck(L.trrh_decompress(compression, bytes(data), len(data), C.byref(out), C.byref(n)))
    return _take_bytes(L, out, n)


class BM25Index:
    """Reference `BM25Index` + `SparseIndex` impl (src/index.rs:30-280); scoring on the device."""

    def __init__(self, k1: float = 1.2, b: float = 0.75, _h=None):
        self.L = _lib.load()
        self.h = _h or C.c_void_p()
        if _h is None:
            _hcheck(self.L.trrh_bm25_new(k1, b, C.byref(self.h)))

    def to_bytes(self) -> bytes:
        """`bincode::serialize(&index)` — the reference's on-disk layout of `BM25Index`."""
        out, n = C.c_void_p(), C.c_uint64()
        _hcheck(self.L.trrh_bm25_to_bytes(self.h, -1, C.byref(out), C.byref(n)))
        return _take_bytes(self.L, out, n)

    @staticmethod
    def from_bytes(data: bytes) -> "BM25Index":
        h = C.c_void_p()
        _hcheck(_lib.load().trrh_bm25_from_bytes(bytes(data), len(data), -1, C.byref(h)))
        return BM25Index(_h=h)

    def to_compressed_bytes(self, compression: int = Compression.Lz4) -> bytes:
        """`BM25Index::to_compressed_bytes` (src/compressed.rs:92-94)."""
        out, n = C.c_void_p(), C.c_uint64()
        _hcheck(self.L.trrh_bm25_to_bytes(self.h, compression, C.byref(out), C.byref(n)))
        return _take_bytes(self.L, out, n)

    @staticmethod
    def from_compressed_bytes(data: bytes, compression: int = Compression.Lz4) -> "BM25Index":
        """`BM25Index::from_compressed_bytes` (src/compressed.rs:101-103)."""
        h = C.c_void_p()
        _hcheck(_lib.load().trrh_bm25_from_bytes(bytes(data), len(data), compression, C.byref(h)))
        return BM25Index(_h=h)

    @staticmethod
    def with_params(k1: float, b: float) -> "BM25Index":
        return BM25Index(k1, b)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.trrh_bm25_free(self.h)
                self.h = None
        except Exception:
            pass

    def _release(self):
        h, self.h = self.h, None
        return h

    @property
    def k1(self):
        return float(self.L.trrh_bm25_k1(self.h))

    @property
    def b(self):
        return float(self.L.trrh_bm25_b(self.h))

    @property
    def avg_doc_length(self):
        return float(self.L.trrh_bm25_avgdl(self.h))

    def tokenize(self, text: str) -> List[str]:
        raw = text.encode()
        cap = 4 * len(raw) + 16
        buf = C.create_string_buffer(cap)
        n = C.c_uint32()
        _hcheck(self.L.trrh_bm25_tokenize(self.h, raw, buf, cap, C.byref(n)))
        s = buf.raw[:n.value].decode()
        return s.split("\n") if s else []

    def add(self, chunk: Chunk):
        _hcheck(self.L.trrh_bm25_add(self.h, chunk.id._c(), chunk.content.encode()))

    def add_batch(self, chunks):
        for c in chunks:
            self.add(c)

    def contains_term(self, term: str) -> bool:
        return bool(self.L.trrh_bm25_contains_term(self.h, term.encode()))

    def search(self, query: str, k: int) -> List[Tuple[ChunkId, float]]:
        cap = max(k, 1)
        ids, sc, n = _ids_out(cap), np.zeros(cap, np.float32), C.c_uint32()
        _hcheck(self.L.trrh_bm25_search(self.h, query.encode(), k, ids, _p(sc, f32p), C.byref(n)))
        return [(ChunkId._from_c(ids[i]), float(sc[i])) for i in range(n.value)]

    def remove(self, chunk_id: ChunkId):
        _hcheck(self.L.trrh_bm25_remove(self.h, chunk_id._c()))

    def __len__(self):
        return int(self.L.trrh_bm25_len(self.h))

    def is_empty(self):
        return len(self) == 0


@dataclass
class PersistedChunk:
    """crates/trueno-rag-cli/src/main.rs:148-153"""
    content: str
    title: Optional[str] = None
    source: Optional[str] = None


class PersistedIndex:
    """The CLI's `index.json` (crates/trueno-rag-cli/src/main.rs:133-146): parsed by the library (serde_json::from_str,
    :437-439); `query` is `run_query`'s cosine scan + stable sort + truncate (:479-495) on the device."""

    def __init__(self, _h):
        self.L = _lib.load()
        self.h = _h

    @staticmethod
    def new(dimension: int, embedder_type: str = "", model_name: Optional[str] = None) -> "PersistedIndex":
        h = C.c_void_p()
        _hcheck(_lib.load().trrh_cli_index_new(dimension, embedder_type.encode(),
                                               None if model_name is None else model_name.encode(), C.byref(h)))
        return PersistedIndex(h)

    def push(self, chunk: "PersistedChunk", embedding) -> None:
        """one chunk and its embedding (`run_index`, crates/trueno-rag-cli/src/main.rs:380-414)"""
        e = np.ascontiguousarray(embedding, dtype=np.float32)
        _hcheck(self.L.trrh_cli_index_push(self.h, chunk.content.encode(), None if chunk.title is None else chunk.title.encode(),
                                           None if chunk.source is None else chunk.source.encode(), _p(e, f32p), e.size))

    def to_json(self) -> str:
        """`serde_json::to_string_pretty(&persisted)` (:423)"""
        out, n = C.c_void_p(), C.c_uint64()
        _hcheck(self.L.trrh_cli_index_to_json(self.h, C.byref(out), C.byref(n)))
        return _take_bytes(self.L, out, n).decode()

    @staticmethod
    def from_json(text) -> "PersistedIndex":
        raw = text.encode() if isinstance(text, str) else bytes(text)
        h = C.c_void_p()
        _hcheck(_lib.load().trrh_cli_index_from_json(raw, len(raw), C.byref(h)))
        return PersistedIndex(h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.trrh_cli_index_free(self.h)
                self.h = None
        except Exception:
            pass

    def __len__(self):
        return int(self.L.trrh_cli_index_len(self.h))

    @property
    def n_embeddings(self) -> int:
        return int(self.L.trrh_cli_index_n_embeddings(self.h))

    @property
    def dimension(self) -> int:
        return int(self.L.trrh_cli_index_dimension(self.h))

    @property
    def embedder_type(self) -> str:
        return self.L.trrh_cli_index_embedder_type(self.h).decode()

    @property
    def model_name(self) -> Optional[str]:
        v = self.L.trrh_cli_index_model_name(self.h)
        return None if v is None else v.decode()

    def chunk(self, i: int) -> PersistedChunk:
        content, n, title, source = C.c_void_p(), C.c_uint64(), C.c_char_p(), C.c_char_p()
        _hcheck(self.L.trrh_cli_index_chunk(self.h, i, C.byref(content), C.byref(n), C.byref(title), C.byref(source)))
        return PersistedChunk(C.string_at(content, n.value).decode(), None if title.value is None else title.value.decode(),
                              None if source.value is None else source.value.decode())

    def embedding(self, i: int) -> np.ndarray:
        data, n = f32p(), C.c_uint64()
        _hcheck(self.L.trrh_cli_index_embedding(self.h, i, C.byref(data), C.byref(n)))
        return np.ctypeslib.as_array(data, shape=(n.value,)).copy() if n.value else np.zeros(0, np.float32)

    def query(self, query_embedding, top_k: int) -> List[Tuple[int, float]]:
        q = np.ascontiguousarray(query_embedding, dtype=np.float32)
        cap = max(min(top_k, self.n_embeddings), 1)
        idx, sc, n = (C.c_uint64 * cap)(), np.zeros(cap, np.float32), C.c_uint64()
        _hcheck(self.L.trrh_cli_index_query(self.h, _p(q, f32p), q.size, top_k, idx, _p(sc, f32p), C.byref(n)))
        return [(int(idx[i]), float(sc[i])) for i in range(n.value)]


@dataclass
class FusionStrategy:
    """Reference `FusionStrategy` (src/fusion.rs:9-63)."""
    kind: int = RRF
    param: float = 60.0

    @staticmethod
    def RRF(k: float = 60.0):
        return FusionStrategy(RRF, k)

    @staticmethod
    def Linear(dense_weight: float):
        return FusionStrategy(LINEAR, dense_weight)

    @staticmethod
    def Convex(alpha: float):
        return FusionStrategy(CONVEX, alpha)

    @staticmethod
    def DBSF():
        return FusionStrategy(DBSF, 0.0)

    @staticmethod
    def Union():
        return FusionStrategy(UNION, 0.0)

    @staticmethod
    def Intersection():
        return FusionStrategy(INTERSECTION, 0.0)

    def fuse(self, dense_results, sparse_results) -> List[Tuple[ChunkId, float]]:
        L = _lib.load()
        nd, ns = len(dense_results), len(sparse_results)
        d_ids, s_ids = _ids_out(nd), _ids_out(ns)
        d_sc = np.array([s for _, s in dense_results] or [0.0], np.float32)
        s_sc = np.array([s for _, s in sparse_results] or [0.0], np.float32)
        for i, (cid, _) in enumerate(dense_results):
            d_ids[i] = cid._c()
        for i, (cid, _) in enumerate(sparse_results):
            s_ids[i] = cid._c()
        out_ids, out_sc, n = _ids_out(nd + ns), np.zeros(max(nd + ns, 1), np.float32), C.c_uint32()
        _hcheck(L.trrh_fuse(self.kind, self.param, d_ids, _p(d_sc, f32p), nd, s_ids, _p(s_sc, f32p), ns, out_ids,
                            _p(out_sc, f32p), C.byref(n)))
        return [(ChunkId._from_c(out_ids[i]), float(out_sc[i])) for i in range(n.value)]


@dataclass
class RetrievalResult:
    """Reference `RetrievalResult` (src/retrieve.rs:13-76)."""
    chunk_id: ChunkId
    content: Optional[str]
    dense_score: Optional[float] = None
    sparse_score: Optional[float] = None
    fused_score: Optional[float] = None
    rerank_score: Optional[float] = None

    def best_score(self) -> float:
        for s in (self.rerank_score, self.fused_score, self.dense_score, self.sparse_score):
            if s is not None:
                return s
        return 0.0


@dataclass
class HybridRetrieverConfig:
    candidates_per_source: int = 50
    fusion: FusionStrategy = field(default_factory=FusionStrategy)
    use_dense: bool = True
    use_sparse: bool = True


class HybridRetriever:
    """Reference `HybridRetriever<E>` (src/retrieve.rs:103-263).  `embedder` is any callable str -> vector."""

    def __init__(self, dense: VectorStore, sparse: BM25Index, embedder, config: Optional[HybridRetrieverConfig] = None):
        self.L = _lib.load()
        self.embedder = embedder
        self.config = config or HybridRetrieverConfig()
        self.dimension = dense.dimension
        self.h = C.c_void_p()
        c = self.config
        _hcheck(self.L.trrh_retriever_new(dense._release(), sparse._release(), c.candidates_per_source, c.fusion.kind,
                                          c.fusion.param, int(c.use_dense), int(c.use_sparse), C.byref(self.h)))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.trrh_retriever_free(self.h)
                self.h = None
        except Exception:
            pass

    def index(self, chunk: Chunk):
        emb = None if chunk.embedding is None else np.ascontiguousarray(chunk.embedding, dtype=np.float32)
        _hcheck(self.L.trrh_retriever_index(self.h, chunk.id._c(), chunk.content.encode(), _p(emb, f32p),
                                            0 if emb is None else emb.size, int(emb is not None)))

    def index_batch(self, chunks):
        for c in chunks:
            self.index(c)

    def _run(self, which: int, query: str, k: int) -> List[RetrievalResult]:
        q = np.ascontiguousarray(self.embedder(query), dtype=np.float32)
        cap = max(k, 1)
        ids = _ids_out(cap)
        f, d, s = (np.zeros(cap, np.float32) for _ in range(3))
        n = C.c_uint32()
        _hcheck(self.L.trrh_retriever_retrieve(self.h, which, query.encode(), _p(q, f32p), q.size, k, ids, _p(f, f32p),
                                               _p(d, f32p), _p(s, f32p), C.byref(n)))
        opt = lambda v: None if np.isnan(v) else float(v)
        return [RetrievalResult(ChunkId._from_c(ids[i]), None, opt(d[i]), opt(s[i]), opt(f[i])) for i in range(n.value)]

    def retrieve(self, query: str, k: int):
        return self._run(0, query, k)

    def retrieve_dense(self, query: str, k: int):
        return self._run(1, query, k)

    def retrieve_sparse(self, query: str, k: int):
        return self._run(2, query, k)

    def __len__(self):
        return int(self.L.trrh_retriever_len(self.h))

    def is_empty(self):
        return len(self) == 0
