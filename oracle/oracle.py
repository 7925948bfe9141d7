"""ctypes binding of the CPU oracle (oracle/trr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  Nothing under trueno_rag_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtrr_oracle.so")

COSINE, EUCLIDEAN, DOT = 0, 1, 2
RRF, LINEAR, CONVEX, DBSF, UNION, INTERSECTION = 0, 1, 2, 3, 4, 5


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "trr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _proto(_lib)
    return _lib


u8p, u16p, u32p, u64p, f32p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64, C.c_float))


def _proto(L):
    L.orc_dot.restype = C.c_float
    L.orc_dot.argtypes = [f32p, f32p, C.c_uint32]
    L.orc_cosine.restype = C.c_float
    L.orc_cosine.argtypes = [f32p, f32p, C.c_uint32]
    L.orc_euclidean.restype = C.c_float
    L.orc_euclidean.argtypes = [f32p, f32p, C.c_uint32]
    L.orc_dense_score.restype = C.c_float
    L.orc_dense_score.argtypes = [C.c_int, f32p, f32p, C.c_uint32]
    L.orc_dense_search.restype = C.c_uint32
    L.orc_dense_search.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_uint32, u8p, f32p, C.c_uint32,
                                   C.c_int, u32p, f32p]
    L.orc_dense_search_batch.restype = None
    L.orc_dense_search_batch.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_uint32, u8p, f32p, C.c_uint32,
                                         C.c_uint32, C.c_int, C.c_int, u32p, f32p, u32p]
    L.orc_dense_search_par.restype = C.c_uint32
    L.orc_dense_search_par.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_uint64, C.c_uint32, f32p, C.c_uint32,
                                       C.c_int, u32p, f32p]
    L.orc_bm25_build.restype = C.c_void_p
    L.orc_bm25_build.argtypes = [u64p, u32p, C.c_uint32, C.c_uint32, C.c_float, C.c_float]
    L.orc_bm25_from_csr.restype = C.c_void_p
    L.orc_bm25_from_csr.argtypes = [C.c_uint32, C.c_uint32, u64p, u32p, u32p, u32p, u32p, C.c_float, C.c_float, C.c_float]
    L.orc_bm25_free.restype = None
    L.orc_bm25_free.argtypes = [C.c_void_p]
    L.orc_bm25_set_stat_docs.restype = None
    L.orc_bm25_set_stat_docs.argtypes = [C.c_void_p, C.c_uint32]
    L.orc_bm25_n_postings.restype = C.c_uint64
    L.orc_bm25_n_postings.argtypes = [C.c_void_p]
    L.orc_bm25_avgdl.restype = C.c_float
    L.orc_bm25_avgdl.argtypes = [C.c_void_p]
    for name, rt in (("term_off", u64p), ("post_doc", u32p), ("post_tf", u32p), ("doc_len", u32p), ("df", u32p)):
        f = getattr(L, "orc_bm25_" + name)
        f.restype = rt
        f.argtypes = [C.c_void_p]
    L.orc_bm25_idf.restype = C.c_float
    L.orc_bm25_idf.argtypes = [C.c_uint32, C.c_uint32]
    L.orc_bm25_score_term.restype = C.c_float
    L.orc_bm25_score_term.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float]
    for name in ("orc_bm25_search_literal", "orc_bm25_search"):
        f = getattr(L, name)
        f.restype = C.c_uint32
        f.argtypes = [C.c_void_p, u32p, C.c_uint32, C.c_uint32, u32p, f32p]
    L.orc_bm25_search_batch.restype = None
    L.orc_bm25_search_batch.argtypes = [C.c_void_p, u32p, u32p, C.c_uint32, C.c_uint32, C.c_int, u32p, f32p, u32p]
    L.orc_min_max.restype = None
    L.orc_min_max.argtypes = [f32p, C.c_uint32, f32p]
    L.orc_z_score.restype = None
    L.orc_z_score.argtypes = [f32p, C.c_uint32, f32p]
    L.orc_fuse.restype = C.c_uint32
    L.orc_fuse.argtypes = [C.c_int, C.c_float, u32p, f32p, C.c_uint32, u32p, f32p, C.c_uint32, u32p, f32p]
    L.orc_hybrid_assemble.restype = C.c_uint32
    L.orc_hybrid_assemble.argtypes = [C.c_int, C.c_float, u32p, f32p, C.c_uint32, u32p, f32p, C.c_uint32, C.c_uint32,
                                      u8p, u32p, f32p, f32p, f32p]
    L.orc_synth_corpus_rows.restype = None
    L.orc_synth_corpus_rows.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int, f32p, u16p]
    L.orc_synth_queries.restype = None
    L.orc_synth_queries.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.c_int, f32p]
    L.orc_synth_doc_tokens.restype = C.c_uint64
    L.orc_synth_doc_tokens.argtypes = [C.c_uint64, u64p, C.c_uint32, C.c_uint64, C.c_uint64, u64p, u32p]
    L.orc_synth_query_terms.restype = C.c_uint64
    L.orc_synth_query_terms.argtypes = [C.c_uint64, u64p, C.c_uint32, C.c_uint64, C.c_uint64, u32p, u32p]


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


# ---------------------------------------------------------------------------------------------- dense
def cosine(a, b):
    a, b = _f32(a), _f32(b)
    if len(a) != len(b):  # reference src/embed.rs:313-315 (public copy): 0.0 on length mismatch
        return 0.0
    return float(lib().orc_cosine(_p(a, f32p), _p(b, f32p), len(a)))


def dot(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_dot(_p(a, f32p), _p(b, f32p), len(a)))


def euclidean(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_euclidean(_p(a, f32p), _p(b, f32p), len(a)))


def dense_search(rows, q, k, metric=COSINE, alive=None, literal=True):
    """rows: (n,d) float32, or (n,d) uint16 holding bf16 bits.  Returns (ords, scores)."""
    is_bf16 = rows.dtype == np.uint16
    rows = np.ascontiguousarray(rows)
    n, d = rows.shape
    q = _f32(q)
    assert q.shape == (d,)
    al = np.ascontiguousarray(alive, dtype=np.uint8) if alive is not None else None
    out_o = np.zeros(max(k, 1), np.uint32)
    out_s = np.zeros(max(k, 1), np.float32)
    cnt = lib().orc_dense_search(metric, rows.ctypes.data, int(is_bf16), n, d, _p(al, u8p), _p(q, f32p), k,
                                 int(literal), _p(out_o, u32p), _p(out_s, f32p))
    return out_o[:cnt].copy(), out_s[:cnt].copy()


def dense_search_batch(rows, Q, k, metric=COSINE, alive=None, literal=False, threads=None):
    is_bf16 = rows.dtype == np.uint16
    rows = np.ascontiguousarray(rows)
    n, d = rows.shape
    Q = _f32(Q)
    B = Q.shape[0]
    al = np.ascontiguousarray(alive, dtype=np.uint8) if alive is not None else None
    out_o = np.zeros((B, max(k, 1)), np.uint32)
    out_s = np.zeros((B, max(k, 1)), np.float32)
    out_n = np.zeros(B, np.uint32)
    lib().orc_dense_search_batch(metric, rows.ctypes.data, int(is_bf16), n, d, _p(al, u8p), _p(Q, f32p), B, k,
                                 int(literal), threads or os.cpu_count() or 1, _p(out_o, u32p), _p(out_s, f32p),
                                 _p(out_n, u32p))
    return out_o, out_s, out_n


def dense_search_par(rows, q, k, metric=COSINE, threads=None):
    is_bf16 = rows.dtype == np.uint16
    rows = np.ascontiguousarray(rows)
    n, d = rows.shape
    q = _f32(q)
    out_o = np.zeros(max(k, 1), np.uint32)
    out_s = np.zeros(max(k, 1), np.float32)
    cnt = lib().orc_dense_search_par(metric, rows.ctypes.data, int(is_bf16), n, d, _p(q, f32p), k,
                                     threads or os.cpu_count() or 1, _p(out_o, u32p), _p(out_s, f32p))
    return out_o[:cnt].copy(), out_s[:cnt].copy()


# ---------------------------------------------------------------------------------------------- bm25
class BM25:
    """Oracle BM25 index over tokenised documents (term ids)."""

    def __init__(self, docs_tokens=None, n_terms=None, k1=1.2, b=0.75, doc_off=None, tokens=None):
        if docs_tokens is not None:
            lens = np.array([len(t) for t in docs_tokens], dtype=np.uint64)
            doc_off = np.zeros(len(docs_tokens) + 1, np.uint64)
            np.cumsum(lens, out=doc_off[1:])
            tokens = np.array([x for t in docs_tokens for x in t], dtype=np.uint32)
        self.doc_off = np.ascontiguousarray(doc_off, dtype=np.uint64)
        self.tokens = np.ascontiguousarray(tokens, dtype=np.uint32)
        if self.tokens.size == 0:
            self.tokens = np.zeros(1, np.uint32)
        self.n_docs = len(self.doc_off) - 1
        self.n_terms = int(n_terms)
        self.k1, self.b = k1, b
        self.h = lib().orc_bm25_build(_p(self.doc_off, u64p), _p(self.tokens, u32p), self.n_docs, self.n_terms, k1, b)

    @classmethod
    def from_csr(cls, n_docs, n_terms, term_off, post_doc, post_tf, doc_len, df, avgdl, k1=1.2, b=0.75):
        """Oracle index over caller-provided CSR arrays (kept alive by this object, not copied)."""
        self = cls.__new__(cls)
        self.n_docs, self.n_terms, self.k1, self.b = int(n_docs), int(n_terms), k1, b
        self._keep = [np.ascontiguousarray(term_off, np.uint64), np.ascontiguousarray(post_doc, np.uint32),
                      np.ascontiguousarray(post_tf, np.uint32), np.ascontiguousarray(doc_len, np.uint32),
                      np.ascontiguousarray(df, np.uint32)]
        t, pd, ptf, dl, dfa = self._keep
        self.h = lib().orc_bm25_from_csr(self.n_docs, self.n_terms, _p(t, u64p), _p(pd, u32p), _p(ptf, u32p), _p(dl, u32p),
                                         _p(dfa, u32p), C.c_float(avgdl), C.c_float(k1), C.c_float(b))
        return self

    def set_stat_docs(self, n_docs_global: int):
        """This index is one shard: N of the idf formula is the global document count (df / avgdl already are global)."""
        lib().orc_bm25_set_stat_docs(self.h, int(n_docs_global))

    def __del__(self):
        try:
            if self.h:
                lib().orc_bm25_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def n_postings(self):
        return int(lib().orc_bm25_n_postings(self.h))

    @property
    def avgdl(self):
        return float(lib().orc_bm25_avgdl(self.h))

    def csr(self):
        L = lib()
        P = self.n_postings
        term_off = np.ctypeslib.as_array(L.orc_bm25_term_off(self.h), (self.n_terms + 1,)).copy()
        post_doc = np.ctypeslib.as_array(L.orc_bm25_post_doc(self.h), (max(P, 1),))[:P].copy()
        post_tf = np.ctypeslib.as_array(L.orc_bm25_post_tf(self.h), (max(P, 1),))[:P].copy()
        doc_len = np.ctypeslib.as_array(L.orc_bm25_doc_len(self.h), (max(self.n_docs, 1),))[:self.n_docs].copy()
        df = np.ctypeslib.as_array(L.orc_bm25_df(self.h), (max(self.n_terms, 1),))[:self.n_terms].copy()
        return term_off, post_doc, post_tf, doc_len, df

    def search(self, q_terms, k, literal=False):
        q = _u32(q_terms)
        out_o = np.zeros(max(k, 1), np.uint32)
        out_s = np.zeros(max(k, 1), np.float32)
        fn = lib().orc_bm25_search_literal if literal else lib().orc_bm25_search
        qq = q if q.size else np.zeros(1, np.uint32)
        cnt = fn(self.h, _p(qq, u32p), q.size, k, _p(out_o, u32p), _p(out_s, f32p))
        return out_o[:cnt].copy(), out_s[:cnt].copy()

    def search_batch(self, q_terms, q_off, k, threads=None):
        q_terms, q_off = _u32(q_terms), _u32(q_off)
        B = len(q_off) - 1
        out_o = np.zeros((B, max(k, 1)), np.uint32)
        out_s = np.zeros((B, max(k, 1)), np.float32)
        out_n = np.zeros(B, np.uint32)
        lib().orc_bm25_search_batch(self.h, _p(q_terms, u32p), _p(q_off, u32p), B, k, threads or os.cpu_count() or 1,
                                    _p(out_o, u32p), _p(out_s, f32p), _p(out_n, u32p))
        return out_o, out_s, out_n


def bm25_idf(n_docs, df):
    return float(lib().orc_bm25_idf(n_docs, df))


def bm25_score_term(tf, df, n_docs, doc_len, avgdl, k1=1.2, b=0.75):
    return float(lib().orc_bm25_score_term(tf, df, n_docs, doc_len, avgdl, k1, b))


# ---------------------------------------------------------------------------------------------- fusion
def min_max(s):
    s = _f32(s)
    out = np.zeros_like(s)
    if s.size:
        lib().orc_min_max(_p(s, f32p), s.size, _p(out, f32p))
    return out


def z_score(s):
    s = _f32(s)
    out = np.zeros_like(s)
    if s.size:
        lib().orc_z_score(_p(s, f32p), s.size, _p(out, f32p))
    return out


def _lists(ids, sc):
    ids, sc = _u32(ids), _f32(sc)
    assert ids.shape == sc.shape
    n = ids.size
    if n == 0:
        ids, sc = np.zeros(1, np.uint32), np.zeros(1, np.float32)
    return ids, sc, n


def fuse(strategy, param, dense, sparse):
    """dense/sparse: (ids, scores) pairs.  Returns (ids, scores) of the fused list."""
    d_id, d_sc, nd = _lists(*dense)
    s_id, s_sc, ns = _lists(*sparse)
    out_i = np.zeros(max(nd + ns, 1), np.uint32)
    out_s = np.zeros(max(nd + ns, 1), np.float32)
    n = lib().orc_fuse(strategy, param, _p(d_id, u32p), _p(d_sc, f32p), nd, _p(s_id, u32p), _p(s_sc, f32p), ns,
                       _p(out_i, u32p), _p(out_s, f32p))
    return out_i[:n].copy(), out_s[:n].copy()


def hybrid_assemble(strategy, param, dense, sparse, k, alive=None):
    d_id, d_sc, nd = _lists(*dense)
    s_id, s_sc, ns = _lists(*sparse)
    al = np.ascontiguousarray(alive, dtype=np.uint8) if alive is not None else None
    kk = max(k, 1)
    out_i = np.zeros(kk, np.uint32)
    out_f, out_d, out_s = (np.zeros(kk, np.float32) for _ in range(3))
    n = lib().orc_hybrid_assemble(strategy, param, _p(d_id, u32p), _p(d_sc, f32p), nd, _p(s_id, u32p), _p(s_sc, f32p),
                                  ns, k, _p(al, u8p), _p(out_i, u32p), _p(out_f, f32p), _p(out_d, f32p),
                                  _p(out_s, f32p))
    return out_i[:n].copy(), out_f[:n].copy(), out_d[:n].copy(), out_s[:n].copy()


# ---------------------------------------------------------------------------------------------- synthetic inputs
def zipf_cdf(n_terms: int, clip: int = 90) -> np.ndarray:
    """u64 CDF table of the clipped Zipf(s=1) over ranks clip+1 .. clip+n_terms."""
    r = np.arange(clip + 1, clip + 1 + n_terms, dtype=np.float64)
    c = np.cumsum(1.0 / r)
    c /= c[-1]
    t = np.minimum(np.floor(c * 18446744073709551616.0), 18446744073709549568.0).astype(np.uint64)
    t[-1] = np.uint64(0xFFFFFFFFFFFFFFFF)
    return t


def synth_corpus(seed, row0, n, d, bf16=False, dups=False):
    """Returns float32 rows (values already bf16-rounded if bf16) and, if bf16, the raw uint16 bits."""
    f = np.zeros((n, d), np.float32)
    b = np.zeros((n, d), np.uint16) if bf16 else None
    lib().orc_synth_corpus_rows(seed, row0, n, d, int(bf16), int(dups), _p(f, f32p), _p(b, u16p))
    return f, b


def synth_queries(seed, q0, n, d, n_corpus, corpus_bf16=False, dups=False):
    out = np.zeros((n, d), np.float32)
    lib().orc_synth_queries(seed, q0, n, d, n_corpus, int(corpus_bf16), int(dups), _p(out, f32p))
    return out


def synth_doc_tokens(seed, cdf, doc0, n):
    cdf = np.ascontiguousarray(cdf, dtype=np.uint64)
    doc_off = np.zeros(n + 1, np.uint64)
    total = lib().orc_synth_doc_tokens(seed, _p(cdf, u64p), len(cdf), doc0, n, _p(doc_off, u64p), None)
    toks = np.zeros(max(int(total), 1), np.uint32)
    lib().orc_synth_doc_tokens(seed, _p(cdf, u64p), len(cdf), doc0, n, _p(doc_off, u64p), _p(toks, u32p))
    return doc_off, toks[:int(total)]


def synth_query_terms(seed, cdf, q0, n):
    cdf = np.ascontiguousarray(cdf, dtype=np.uint64)
    q_off = np.zeros(n + 1, np.uint32)
    total = lib().orc_synth_query_terms(seed, _p(cdf, u64p), len(cdf), q0, n, _p(q_off, u32p), None)
    terms = np.zeros(max(int(total), 1), np.uint32)
    lib().orc_synth_query_terms(seed, _p(cdf, u64p), len(cdf), q0, n, _p(q_off, u32p), _p(terms, u32p))
    return q_off, terms[:int(total)]
