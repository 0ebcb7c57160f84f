"""`fused_clip_loss`: the one public op of the hot path, as a torch.autograd.Function.

It replaces, in one call, the tail every reference model/loss shares
(F.normalize x2 -> exp(logit_scale) -> matmul * scale -> cross_entropy on S and S^T -> /2, and its
autograd backward; old/clip.py:63-67, current/rna_clip_codes.ipynb:1948-1953, run1/full.py:88-100,
old/clip_opt.py:130-151, tong/utils/losses.py:4-19).  Saved for backward: O(N d) only -- the N x N
logits never exist in HBM.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import step as _step
from .engine import default_engine


def _scale_value(logit_scale, scale_is_log: bool, clamp_max: Optional[float]):
    """Host value of s and whether the clamp is active (d s / d logit_scale = 0 there)."""
    t = float(logit_scale.detach()) if torch.is_tensor(logit_scale) else float(logit_scale)   # one host read of a 0-d parameter per step (the reference reads loss.item())
    s = math.exp(t) if scale_is_log else t
    clamped = False
    if clamp_max is not None and s > clamp_max:
        s, clamped = float(clamp_max), True
    return t, s, clamped


class _FusedClipLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, logit_scale, extra, symmetric, scale_is_log, clamp_max, group, compute_dtype, flags, engine,
                want_stats=True):
        need_grad = a.requires_grad or b.requires_grad or (torch.is_tensor(logit_scale) and logit_scale.requires_grad)
        info = {}

        def scale_now():   # called by the step after the scale-independent kernels and the all-gather are enqueued
            info["v"] = _scale_value(logit_scale, scale_is_log, clamp_max)
            return info["v"][1]

        loss, st = _step.contrastive_forward(engine, a.detach().contiguous(), b.detach().contiguous(), scale_now,
                                             symmetric=symmetric, extra=extra, group=group,
                                             compute_dtype=compute_dtype, flags=flags, need_grad=need_grad)
        t, s, clamped = info["v"]
        ctx.st, ctx.engine = st, engine
        ctx.scale_info = (s, clamped, scale_is_log)
        ctx.ls_meta = (logit_scale.dtype, logit_scale.device) if torch.is_tensor(logit_scale) else None
        if want_stats:
            row_lse = engine.combine_lse(st.row_m, st.row_l)
            col_lse = engine.combine_lse(st.col_m, st.col_l)
        else:   # two launches nobody reads
            row_lse = col_lse = st.diag
        ctx.mark_non_differentiable(row_lse, col_lse, st.diag)
        return loss.reshape(()), row_lse, col_lse, st.diag

    @staticmethod
    def backward(ctx, g_loss, _g1, _g2, _g3):
        st, engine = ctx.st, ctx.engine
        s, clamped, scale_is_log = ctx.scale_info
        g = g_loss.reshape(1).to(torch.float32).contiguous()
        da, db, ds = _step.contrastive_backward(engine, st, grad_scale=g)
        d_ls = None
        if ctx.ls_meta is not None and ctx.needs_input_grad[2]:
            if clamped:
                d_ls = torch.zeros((), dtype=ctx.ls_meta[0], device=ctx.ls_meta[1])
            else:
                # sum G.S = dL/dt for s = exp(t);  dL/ds = sum G.S / s for a raw scale
                coef = 1.0 if scale_is_log else 1.0 / s
                d_ls = (ds * g * coef).reshape(()).to(device=ctx.ls_meta[1], dtype=ctx.ls_meta[0])
        ctx.st = None
        return da, db, d_ls, None, None, None, None, None, None, None, None, None


def fused_clip_loss(a, b, logit_scale, *, symmetric: bool = True, scale_is_log: bool = True,
                    clamp_max: Optional[float] = None, extra_cols=None, extra_normalized: bool = True,
                    group=None, compute_dtype: Optional[torch.dtype] = None, return_stats: bool = False,
                    engine=None):
    """Fused CLIP / InfoNCE loss of two [N,d] embedding batches (un-normalised projection-head outputs).

    logit_scale   0-d tensor (learnable parameter) or float; ``s = exp(logit_scale)`` (scale_is_log) or
                  ``s = logit_scale`` (tong's ``1 / temperature``), optionally clamped (old/clip_opt.py:100).
    symmetric     (CE(S) + CE(S^T)) / 2 (rna_clip_codes.ipynb:1953) or CE(S) only (run1/full.py:133).
    extra_cols    [M,d] additional negative columns (hard-negative cache old/clip_opt.py:118-121, memory
                  queue tong/utils/losses.py:10-11), used as stored, no gradient.  ``extra_normalized=False``
                  (rows of arbitrary norm) routes to the exact kernels.
    group         torch.distributed process group: rows are this rank's shard of a global batch; negatives
                  are global, gradients are exact (reduce-scattered), the returned loss is the global mean.
    compute_dtype torch.bfloat16 (tcgen05 tensor-core kernels) or torch.float32 (exact check mode).
                  Default: bf16 for bf16/fp16 inputs or under autocast, else fp32 (the reference's numerics).
    """
    engine = engine or default_engine()
    if a.dtype == torch.float16:
        a = a.to(torch.bfloat16)
    if b.dtype == torch.float16:
        b = b.to(torch.bfloat16)
    if compute_dtype is None:
        low = a.dtype == torch.bfloat16 or b.dtype == torch.bfloat16 or torch.is_autocast_enabled()
        compute_dtype = torch.bfloat16 if low else torch.float32
    flags = 0
    if extra_cols is not None and not extra_normalized:
        flags |= _step.FLAG_FORCE_EXACT
    loss, row_lse, col_lse, diag = _FusedClipLoss.apply(a, b, logit_scale, extra_cols, symmetric, scale_is_log,
                                                        clamp_max, group, compute_dtype, flags, engine, bool(return_stats))
    if return_stats:
        return loss, {"row_lse": row_lse, "col_lse": col_lse, "diag": diag}
    return loss
