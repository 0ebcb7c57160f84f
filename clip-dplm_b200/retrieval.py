"""Retrieval on the same similarity sweep: top-k most similar library rows per query, never forming [n_q, n_lib].

Reference anchors: the evaluation tail of run1/full.py -- ``logits.argmax(dim=1)`` (:152, :138-139) and the
``F.cosine_similarity(a.unsqueeze(1), b.unsqueeze(0), dim=2)`` broadcast (:157, an [N,N,d] intermediate) -- and
BASELINE.json config 5 (1M-entry protein library x 16k TF queries, top-10, library sharded over 8 GPUs).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .engine import default_engine


def _gather_cols(x, group):
    world = dist.get_world_size(group)
    outs = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(outs, x.contiguous(), group=group)
    return torch.cat(outs, dim=1)


def topk_similarity(queries, library, k: int = 10, *, group=None, library_offset=None, engine=None):
    """-> (scores [n_q,k] f32 descending, indices [n_q,k] i64 into the GLOBAL library).

    queries  [n_q,d]  the same on every rank;  library [n_lib_local,d]  this rank's row shard (the whole library
    without ``group``).  ``library_offset``: global index of the shard's first row (default: rank * n_lib_local, i.e.
    equal shards).  Cosine similarity of L2-normalised rows (old/clip.py:63-64 normalisation), bf16 tensor-core sweep."""
    engine = engine or default_engine()
    if queries.dim() != 2 or library.dim() != 2 or queries.shape[1] != library.shape[1]:
        raise ValueError(f"expected [n_q,d] and [n_lib,d], got {tuple(queries.shape)} and {tuple(library.shape)}")
    q, _ = engine.stage(queries.detach().contiguous(), torch.bfloat16)
    lib, _ = engine.stage(library.detach().contiguous(), torch.bfloat16)
    rq, _ = engine.normalize(q)
    rl, _ = engine.normalize(lib)
    if group is not None and library_offset is None:
        library_offset = dist.get_rank(group) * library.shape[0]
    scores, idx = engine.topk(q, lib, rq, rl, k, col_offset=int(library_offset or 0))
    if group is not None and dist.get_world_size(group) > 1:
        all_s, all_i = _gather_cols(scores, group), _gather_cols(idx, group)       # [n_q, world * k]
        all_s = torch.where(all_i < 0, torch.full_like(all_s, float("-inf")), all_s)
        scores, pos = torch.topk(all_s, k, dim=1)
        idx = torch.gather(all_i, 1, pos)
    return scores, idx


def top1_accuracy(a, b, *, engine=None):
    """Fraction of rows whose most similar row of ``b`` is their own positive: ``(logits.argmax(dim=1) == arange).mean()``
    of run1/full.py:152-153 without the logits."""
    _, idx = topk_similarity(a, b, 1, engine=engine)
    return (idx[:, 0] == torch.arange(a.shape[0], device=idx.device)).float().mean()
