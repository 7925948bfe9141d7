"""trueno_rag_b200 — B200-native retrieval hot path for trueno-rag (dense scan / tensor-core batch scoring,
BM25 posting scoring, score fusion + top-k) behind the reference's VectorStore / BM25Index / FusionStrategy /
HybridRetriever surface.  The product is libtrueno_rag_b200.so (C ABI in include/); this package only binds it."""
from . import _lib  # noqa: F401  (does not load the shared object until first use)

__all__ = ["_lib", "api", "shard", "build"]
