"""Document sharding across GPUs (one process per GPU).

The corpus is split into contiguous ordinal ranges so that (rank, local ordinal) order equals global ordinal
order and the canonical tie-break survives sharding.  Each rank produces a shard-local top-C per source in an
*exchange record*; the records are all-gathered and every rank runs the merge + fusion kernel on the result.
This module holds the host-side plumbing only: range planning, the record layout, the collective call.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_docs: int, rank: int, world: int):
    """Contiguous range [lo, hi) owned by `rank`; sizes differ by at most one and earlier ranks are larger."""
    base, rem = divmod(n_docs, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def exchange_words(B: int, C: int) -> int:
    """Size of one exchange record in 4-byte words: u32 ord[2][B][C]; f32 score[2][B][C]; u32 n[2][B]."""
    return 4 * B * C + 2 * B


def unpack_exchange(words: np.ndarray, B: int, C: int):
    """Views into one record (numpy uint32 array of exchange_words entries): ord, score, n with leading axis = source."""
    w = np.ascontiguousarray(words, dtype=np.uint32)
    assert w.size == exchange_words(B, C)
    bc = B * C
    ords = w[:2 * bc].reshape(2, B, C)
    scores = w[2 * bc:4 * bc].view(np.float32).reshape(2, B, C)
    n = w[4 * bc:].reshape(2, B)
    return ords, scores, n


def pack_exchange(dense, sparse, B: int, C: int) -> np.ndarray:
    """Builds a record from per-source (ord[B][C], score[B][C], n[B]) triples (host arrays)."""
    w = np.zeros(exchange_words(B, C), np.uint32)
    ords, scores, n = unpack_exchange(w, B, C)
    for s, src in enumerate((dense, sparse)):
        if src is None:
            continue
        o, sc, cnt = src
        ords[s, :, :] = o
        scores[s, :, :] = sc
        n[s, :] = cnt
    return w


def all_gather_records(local_record, world: int):
    """All-gathers one exchange record per rank.  `local_record` is a torch tensor (uint8/int32, on the device for
    NCCL, on the host for gloo); returns a tensor holding `world` records back to back."""
    import torch
    import torch.distributed as dist
    flat = local_record.contiguous().view(-1)
    out = torch.empty(world * flat.numel(), dtype=flat.dtype, device=flat.device)
    if world == 1 or not dist.is_initialized():
        out.copy_(flat)
    else:
        dist.all_gather_into_tensor(out, flat)
    return out.view((world,) + tuple(local_record.shape))
