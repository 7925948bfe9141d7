"""ctypes loader of libtrueno_rag_b200.so (the C ABI of include/trueno_rag_b200.h and trueno_rag_host.h).

The library is the product; this module is only the Python-side binding used by tests and bench.py.
It never falls back to a CPU implementation: if the shared object is missing the import fails, and if no
CUDA device is present every compute entry point fails with TRR_ERR_NO_DEVICE.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libtrueno_rag_b200.so")

TRR_OK, TRR_ERR_INVALID_ARG, TRR_ERR_DIM_MISMATCH, TRR_ERR_CUDA, TRR_ERR_OOM = 0, 1, 2, 3, 4
TRR_ERR_NOT_FROZEN, TRR_ERR_UNSUPPORTED, TRR_ERR_NO_DEVICE = 5, 6, 7

u8p, u16p, u32p, u64p, f32p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint16, C.c_uint32, C.c_uint64, C.c_float))
vp, vpp = C.c_void_p, C.POINTER(C.c_void_p)


class Stats(C.Structure):
    _fields_ = [("mode_used", C.c_uint32), ("n_queries", C.c_uint32), ("n_guard_fallbacks", C.c_uint32),
                ("n_kernel_launches", C.c_uint32), ("ms_total", C.c_float), ("ms_main_kernel", C.c_float),
                ("max_fast_exact_gap", C.c_float), ("eps_bound", C.c_float), ("n_exact_fallbacks", C.c_uint32),
                ("rescore_width", C.c_uint32)]


class HostId(C.Structure):
    _fields_ = [("hi", C.c_uint64), ("lo", C.c_uint64)]


# name -> (restype, argtypes); every symbol declared in include/trueno_rag_b200.h
TRR_PROTOS = {
    "trr_last_error": (C.c_char_p, []),
    "trr_version": (C.c_int, []),
    "trr_device_count": (C.c_int, []),
    "trr_ctx_create": (C.c_int, [C.c_int, vpp]),
    "trr_ctx_destroy": (C.c_int, [vp]),
    "trr_ctx_sync": (C.c_int, [vp]),
    "trr_ctx_stream": (C.c_int, [vp, vpp]),
    "trr_ctx_sm_count": (C.c_int, [vp, C.POINTER(C.c_int)]),
    "trr_ctx_flush_l2": (C.c_int, [vp, C.c_size_t]),
    "trr_ctx_set_stream": (C.c_int, [vp, vp]),
    "trr_ctx_launch_count": (C.c_int, [vp, u64p]),
    "trr_hybrid_local_device": (C.c_int, [vp, vp, vp, vp, vp, u32p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, vp]),
    "trr_hybrid_merge_device": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.c_uint32, vp,
                                          vp, vp, vp, vp]),
    "trr_synth_queries": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                    f32p]),
    "trr_synth_query_terms": (C.c_int, [C.c_uint64, u64p, C.c_uint32, C.c_uint64, C.c_uint64, u32p, u32p, C.c_uint64]),
    "trr_synth_bm25_count": (C.c_int, [C.c_uint64, u64p, C.c_uint32, C.c_uint64, C.c_uint64, u32p, u32p, u64p]),
    "trr_synth_bm25_fill": (C.c_int, [C.c_uint64, u64p, C.c_uint32, C.c_uint64, C.c_uint64, u64p, u32p, u32p]),
    "trr_dense_create": (C.c_int, [vp, C.c_uint32, C.c_int, C.c_int, C.c_uint64, vpp]),
    "trr_dense_destroy": (C.c_int, [vp]),
    "trr_dense_set_base": (C.c_int, [vp, C.c_uint32]),
    "trr_dense_append": (C.c_int, [vp, f32p, C.c_uint64]),
    "trr_dense_append_bf16": (C.c_int, [vp, u16p, C.c_uint64]),
    "trr_dense_append_device": (C.c_int, [vp, vp, C.c_uint64]),
    "trr_dense_append_synth": (C.c_int, [vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]),
    "trr_dense_remove": (C.c_int, [vp, C.c_uint32]),
    "trr_dense_freeze": (C.c_int, [vp]),
    "trr_dense_len": (C.c_int, [vp, u64p]),
    "trr_dense_set_mode": (C.c_int, [vp, C.c_int]),
    "trr_dense_search": (C.c_int, [vp, f32p, C.c_uint32, C.c_uint32, u32p, f32p, u32p]),
    "trr_dense_search_device": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp]),
    "trr_dense_last_stats": (C.c_int, [vp, C.POINTER(Stats)]),
    "trr_dense_copy_norms": (C.c_int, [vp, f32p, C.c_uint64]),
    "trr_dense_copy_rows": (C.c_int, [vp, u32p, C.c_uint64, vp]),
    "trr_dense_info": (C.c_int, [vp, u32p, C.POINTER(C.c_int), C.POINTER(C.c_int), u32p]),
    "trr_dense_save": (C.c_int, [vp, C.c_char_p]),
    "trr_dense_load": (C.c_int, [vp, C.c_char_p, vpp]),
    "trr_bm25_append": (C.c_int, [vp, C.c_uint32, C.c_uint32, u64p, u32p, u32p, u32p, C.c_float, C.c_float, C.c_float, f32p]),
    "trr_bm25_remove": (C.c_int, [vp, u32p, C.c_uint32, C.c_float, C.c_float, C.c_float, f32p, u64p]),
    "trr_bm25_save": (C.c_int, [vp, C.c_char_p]),
    "trr_bm25_load": (C.c_int, [vp, C.c_char_p, vpp]),
    "trr_bm25_build": (C.c_int, [vp, C.c_uint32, C.c_uint32, u64p, u32p, u32p, u32p, C.c_float, C.c_float, C.c_float,
                                 f32p, C.c_uint32, vpp]),
    "trr_bm25_destroy": (C.c_int, [vp]),
    "trr_bm25_n_postings": (C.c_int, [vp, u64p]),
    "trr_bm25_search": (C.c_int, [vp, u32p, u32p, C.c_uint32, C.c_uint32, u32p, f32p, u32p]),
    "trr_bm25_search_device": (C.c_int, [vp, vp, vp, u32p, C.c_uint32, C.c_uint32, vp, vp, vp]),
    "trr_bm25_last_stats": (C.c_int, [vp, C.POINTER(Stats)]),
    "trr_bm25_copy_impacts": (C.c_int, [vp, f32p, C.c_uint64]),
    "trr_fuse": (C.c_int, [vp, C.c_int, C.c_float, u32p, f32p, u32p, u32p, f32p, u32p, C.c_uint32, C.c_uint32,
                           C.c_uint32, u32p, f32p, f32p, f32p, u32p]),
    "trr_hybrid_search": (C.c_int, [vp, vp, f32p, u32p, u32p, C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.c_uint32,
                                    C.c_int, C.c_int, u32p, f32p, f32p, f32p, u32p]),
    "trr_exchange_bytes": (C.c_size_t, [C.c_uint32, C.c_uint32]),
    "trr_hybrid_local": (C.c_int, [vp, vp, f32p, u32p, u32p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, vp]),
    "trr_hybrid_merge": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.c_uint32, u32p,
                                   f32p, f32p, f32p, u32p]),
    "trr_group_unique_id": (C.c_int, [vp]),
    "trr_group_create": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_size_t, vpp]),
    "trr_group_destroy": (C.c_int, [vp]),
    "trr_group_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "trr_group_sync": (C.c_int, [vp]),
    "trr_group_allreduce_u64": (C.c_int, [vp, u64p, C.c_size_t, C.c_int]),
    "trr_hybrid_search_sharded": (C.c_int, [vp, vp, vp, f32p, u32p, u32p, C.c_uint32, C.c_uint32, C.c_int, C.c_float,
                                            C.c_uint32, C.c_int, C.c_int, u32p, f32p, f32p, f32p, u32p]),
    "trr_hybrid_search_sharded_async": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_uint32, C.c_uint32, C.c_int, C.c_float,
                                                  C.c_uint32, C.c_int, C.c_int, vp, vp, vp, vp, vp]),
    "trr_hybrid_search_sharded_device": (C.c_int, [vp, vp, vp, vp, vp, vp, u32p, C.c_uint32, C.c_uint32, C.c_int, C.c_float,
                                                   C.c_uint32, C.c_int, C.c_int, vp, vp, vp, vp, vp]),
}

idp = C.POINTER(HostId)
TRRH_PROTOS = {
    "trrh_last_error": (C.c_char_p, []),
    "trrh_last_expected": (C.c_uint64, []),
    "trrh_last_actual": (C.c_uint64, []),
    "trrh_store_new": (C.c_int, [C.c_uint32, C.c_int, C.c_int, vpp]),
    "trrh_store_free": (None, [vp]),
    "trrh_store_insert": (C.c_int, [vp, HostId, C.c_char_p, f32p, C.c_uint32, C.c_int]),
    "trrh_store_search": (C.c_int, [vp, f32p, C.c_uint32, C.c_uint32, idp, f32p, u32p]),
    "trrh_store_get": (C.c_int, [vp, HostId, C.POINTER(C.c_char_p)]),
    "trrh_store_remove": (C.c_int, [vp, HostId]),
    "trrh_store_len": (C.c_uint64, [vp]),
    "trrh_store_set_mode": (C.c_int, [vp, C.c_int]),
    "trrh_store_clone": (C.c_int, [vp, vpp]),
    "trrh_bm25_new": (C.c_int, [C.c_float, C.c_float, vpp]),
    "trrh_bm25_free": (None, [vp]),
    "trrh_bm25_tokenize": (C.c_int, [vp, C.c_char_p, C.c_char_p, C.c_uint32, u32p]),
    "trrh_bm25_add": (C.c_int, [vp, HostId, C.c_char_p]),
    "trrh_bm25_search": (C.c_int, [vp, C.c_char_p, C.c_uint32, idp, f32p, u32p]),
    "trrh_bm25_remove": (C.c_int, [vp, HostId]),
    "trrh_bm25_len": (C.c_uint64, [vp]),
    "trrh_bm25_avgdl": (C.c_float, [vp]),
    "trrh_bm25_k1": (C.c_float, [vp]),
    "trrh_bm25_b": (C.c_float, [vp]),
    "trrh_bm25_contains_term": (C.c_int, [vp, C.c_char_p]),
    "trrh_compress": (C.c_int, [C.c_int, C.c_char_p, C.c_uint64, C.POINTER(vp), C.POINTER(C.c_uint64)]),
    "trrh_decompress": (C.c_int, [C.c_int, C.c_char_p, C.c_uint64, C.POINTER(vp), C.POINTER(C.c_uint64)]),
    "trrh_bytes_free": (None, [vp]),
    "trrh_bm25_to_bytes": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_uint64)]),
    "trrh_bm25_from_bytes": (C.c_int, [C.c_char_p, C.c_uint64, C.c_int, vpp]),
    "trrh_cli_index_from_json": (C.c_int, [C.c_char_p, C.c_uint64, vpp]),
    "trrh_cli_index_free": (None, [vp]),
    "trrh_cli_index_new": (C.c_int, [C.c_uint64, C.c_char_p, C.c_char_p, vpp]),
    "trrh_cli_index_push": (C.c_int, [vp, C.c_char_p, C.c_char_p, C.c_char_p, f32p, C.c_uint64]),
    "trrh_cli_index_to_json": (C.c_int, [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]),
    "trrh_cli_index_len": (C.c_uint64, [vp]),
    "trrh_cli_index_n_embeddings": (C.c_uint64, [vp]),
    "trrh_cli_index_dimension": (C.c_uint64, [vp]),
    "trrh_cli_index_embedder_type": (C.c_char_p, [vp]),
    "trrh_cli_index_model_name": (C.c_char_p, [vp]),
    "trrh_cli_index_chunk": (C.c_int, [vp, C.c_uint64, C.POINTER(vp), C.POINTER(C.c_uint64), C.POINTER(C.c_char_p),
                                       C.POINTER(C.c_char_p)]),
    "trrh_cli_index_embedding": (C.c_int, [vp, C.c_uint64, C.POINTER(f32p), C.POINTER(C.c_uint64)]),
    "trrh_cli_index_query": (C.c_int, [vp, f32p, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64), f32p,
                                       C.POINTER(C.c_uint64)]),
    "trrh_fuse": (C.c_int, [C.c_int, C.c_float, idp, f32p, C.c_uint32, idp, f32p, C.c_uint32, idp, f32p, u32p]),
    "trrh_retriever_new": (C.c_int, [vp, vp, C.c_uint32, C.c_int, C.c_float, C.c_int, C.c_int, vpp]),
    "trrh_retriever_free": (None, [vp]),
    "trrh_retriever_index": (C.c_int, [vp, HostId, C.c_char_p, f32p, C.c_uint32, C.c_int]),
    "trrh_retriever_retrieve": (C.c_int, [vp, C.c_int, C.c_char_p, f32p, C.c_uint32, C.c_uint32, idp, f32p, f32p, f32p,
                                          u32p]),
    "trrh_retriever_len": (C.c_uint64, [vp]),
}

DEBUG_PROTOS = {
    "trr_debug_gemm_scores": (C.c_int, [vp, f32p, C.c_uint32, f32p, C.c_uint32]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the shared object; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python -m trueno_rag_b200.build` (needs nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for the retrieval kernels.")
    lib = C.CDLL(SO_PATH)
    for table in (TRR_PROTOS, TRRH_PROTOS, DEBUG_PROTOS):
        for name, (res, args) in table.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
    _lib = lib
    return lib
