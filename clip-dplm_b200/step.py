"""Host-side orchestration of one contrastive step (engine-agnostic).

Single GPU:   normalise -> forward statistics -> loss      |  backward (two sides) -> normalise-bwd
Row-sharded global batch over a process group (replaces the autograd-blind
``dist.all_gather`` + ``torch.cat`` of old/clip_opt.py:102-112 / run1/full.py:77-84):

    rank p owns rows [p*n, (p+1)*n) of both modalities
    fwd:  all-gather normalised B  ->  local rows x all columns  ->  all-reduce column (max, sumexp)
          -> all-reduce the scalar loss
    bwd:  dA complete locally;  dB partial [N,d] -> reduce-scatter -> local normalise-bwd

The engine (``engine.CudaEngine`` in production) provides the kernels; tests drive the same code
with a CPU stand-in over gloo to cover the sharding/offset/collective logic without a GPU.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

FLAG_FORCE_EXACT = 1


# ------------------------------------------------------------------------------------------------
# collectives (NCCL on GPUs; gloo in the CPU tests, which lacks some tensor collectives)
# ------------------------------------------------------------------------------------------------
def _all_gather_rows(x, group):
    world = dist.get_world_size(group)
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    if dist.get_backend(group) == "gloo" and x.dtype == torch.bfloat16:
        tmp = torch.empty(out.shape, dtype=torch.float32, device=x.device)
        dist.all_gather_into_tensor(tmp, x.float().contiguous(), group=group)
        return tmp.to(x.dtype)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def _reduce_scatter_rows(x, group, async_op=False):
    """-> (out [n/world, ...], work or None).  With ``async_op`` the NCCL kernel runs on the communicator's own stream
    behind everything already enqueued; the caller keeps launching compute and calls ``work.wait()`` (a stream wait,
    not a host wait) before consuming ``out``."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = x.shape[0] // world
    if dist.get_backend(group) == "gloo":
        y = x.clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y[rank * n:(rank + 1) * n].contiguous(), None
    out = torch.empty((n,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    work = dist.reduce_scatter_tensor(out, x.contiguous(), op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return out, (work if async_op else None)


@dataclass
class StepState:
    a: torch.Tensor                 # caller's embeddings (for the normalise backward)
    b: torch.Tensor
    a_c: torch.Tensor               # raw rows in the compute dtype
    a_c_t: Optional[torch.Tensor]
    y: torch.Tensor                 # all columns: gathered B (+ extra negatives), raw, compute dtype
    y_t: Optional[torch.Tensor]
    rinv_a: torch.Tensor
    rinv_b: torch.Tensor
    rinv_y: torch.Tensor
    row_m: torch.Tensor             # log-sum-exp kept as (shift, sum) pairs: r = row_m + log(row_l)
    row_l: torch.Tensor
    col_m: torch.Tensor             # globally combined column statistics
    col_l: torch.Tensor
    diag: torch.Tensor
    scale: float
    scale_dev: Optional[torch.Tensor]   # device scalar s (the kernels read it; `scale` is then only the host's hint)
    symmetric: bool
    n_local: int
    n_global: int
    diag_offset: int
    flags: int
    group: object


def contrastive_forward(engine, a, b, scale, *, symmetric=True, extra=None, group=None,
                        compute_dtype=torch.bfloat16, flags=0, need_grad=True, scale_dev=None):
    """Returns (loss [1] f32 -- the GLOBAL mean loss, identical on every rank --, StepState).

    ``scale`` is s as a float, or a zero-argument callable returning it: the callable is invoked only after the
    scale-independent work (row norms, the all-gather) has been enqueued, so that a host read of the logit_scale
    parameter overlaps with it instead of leaving the GPU idle."""
    if a.dim() != 2 or b.dim() != 2 or a.shape != b.shape:
        raise ValueError(f"expected two [N,d] embedding matrices of equal shape, got {tuple(a.shape)} and {tuple(b.shape)}")
    n_local = a.shape[0]
    world = dist.get_world_size(group) if group is not None else 1
    rank = dist.get_rank(group) if group is not None else 0
    n_global = n_local * world
    diag_offset = rank * n_local

    rinv_a, _ = engine.normalize(a)
    rinv_b, _ = engine.normalize(b)
    a_c, _ = engine.stage(a, compute_dtype)
    b_c, _ = engine.stage(b, compute_dtype)
    y, rinv_y = b_c, rinv_b
    if world > 1:
        y = _all_gather_rows(b_c, group)
        if b.dtype == compute_dtype:
            rinv_y, _ = engine.normalize(y)   # same kernel on the same rows as on their owner: identical values, one collective less
        else:   # norms were taken on the caller's (wider) rows before staging: ship them
            rinv_y = _all_gather_rows(rinv_b, group)
    if callable(scale):
        scale = scale()
    tc = engine.uses_tensor_cores(compute_dtype, a.shape[1], scale, flags)
    want_t = need_grad and tc and engine.needs_transposed(compute_dtype, a.shape[1], scale, flags)
    a_c_t = b_c_t = None
    if want_t:
        _, a_c_t = engine.stage(a_c, compute_dtype, want_t=True)
        if world == 1 and extra is None:
            _, b_c_t = engine.stage(b_c, compute_dtype, want_t=True)
    y_t = b_c_t
    if extra is not None:   # used as stored (old/clip_opt.py:118-121, tong/utils/losses.py:10-11): rinv = 1
        y = torch.cat([y, extra.detach().to(compute_dtype)], dim=0).contiguous()
        rinv_y = torch.cat([rinv_y, torch.ones(extra.shape[0], dtype=rinv_y.dtype, device=rinv_y.device)])
    if want_t and y_t is None:
        _, y_t = engine.stage(y, compute_dtype, want_t=True)

    kw = {"scale_dev": scale_dev} if scale_dev is not None else {}
    row_m, row_l, col_m, col_l, diag = engine.forward(a_c, y, rinv_a, rinv_y, diag_offset, scale, flags, **kw)
    if world > 1:
        fixed = getattr(engine, "fixed_shift", None)
        if fixed is not None and fixed(compute_dtype, a.shape[1], scale, flags):
            # tensor-core kernels: every partial sum already shares the fixed shift col_m == s on every rank
            dist.all_reduce(col_l, op=dist.ReduceOp.SUM, group=group)
        else:
            m_max = col_m.clone()
            dist.all_reduce(m_max, op=dist.ReduceOp.MAX, group=group)
            col_l = col_l * torch.exp(col_m - m_max)
            dist.all_reduce(col_l, op=dist.ReduceOp.SUM, group=group)
            col_m = m_max
    if y.shape[0] > n_global:
        col_l[n_global:] = float("inf")   # extra negatives carry no positives: no column loss, no column soft-max
    loss = engine.loss(row_m, row_l, col_m, col_l, diag, diag_offset, n_global, symmetric)
    if world > 1:
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)
    st = StepState(a, b, a_c, a_c_t, y, y_t, rinv_a, rinv_b, rinv_y, row_m, row_l, col_m, col_l, diag, scale,
                   scale_dev, symmetric, n_local, n_global, diag_offset, flags, group)
    return loss, st


def contrastive_backward(engine, st: StepState, grad_scale=None, grad_dtype_a=None, grad_dtype_b=None):
    """Returns (dA [n,d], dB [n,d], d_scale_sum [1]) for upstream gradient 1; ``grad_scale`` is an optional
    device scalar multiplied into dA/dB inside the normalise-backward kernel (d_scale_sum is left
    unscaled -- the caller multiplies that single element)."""
    n_glob = st.n_global
    world = dist.get_world_size(st.group) if st.group is not None else 1
    coef = 1.0 / ((2.0 if st.symmetric else 1.0) * n_glob)
    row_w = engine.softmax_weights(st.row_l, coef)
    if st.symmetric:
        col_m, col_w = st.col_m, engine.softmax_weights(st.col_l, coef)
    else:
        col_m, col_w = st.col_m, torch.zeros_like(st.col_l)
    diag_w = 1.0 / n_glob
    kw = {"scale_dev": st.scale_dev} if st.scale_dev is not None else {}

    # side B first: the positive-carrying columns as rows x local rows as columns -> partial dB_hat [N,d]; its
    # reduce-scatter over NVLink then runs behind side A's contraction instead of after it
    db_part, _ = engine.backward(st.y[:n_glob], st.a_c, st.a_c_t, st.rinv_y[:n_glob].contiguous(), st.rinv_a,
                                 -st.diag_offset, st.scale, col_m[:n_glob].contiguous(), col_w[:n_glob].contiguous(),
                                 st.row_m, row_w, diag_w, 1.0, st.flags, want_dscale=False, **kw)
    rs_work = None
    if world > 1:
        db_hat, rs_work = _reduce_scatter_rows(db_part, st.group, async_op=True)
    else:
        db_hat = db_part
    # side A: local rows x all columns -> dA_hat (complete) and sum G.S over the local row block
    da_hat, ds = engine.backward(st.a_c, st.y, st.y_t, st.rinv_a, st.rinv_y, st.diag_offset, st.scale, st.row_m, row_w,
                                 col_m, col_w, diag_w, 1.0, st.flags, want_dscale=True, **kw)
    ds_work = None
    if world > 1:   # the scalar's all-reduce (which also absorbs the ranks' skew) runs behind the normalise backward
        nccl = dist.get_backend(st.group) != "gloo"
        ds_work = dist.all_reduce(ds, op=dist.ReduceOp.SUM, group=st.group, async_op=nccl)
    da = engine.normalize_backward(st.a, st.rinv_a, da_hat, grad_dtype_a or st.a.dtype, grad_scale)
    if rs_work is not None:
        rs_work.wait()
    db = engine.normalize_backward(st.b, st.rinv_b, db_hat, grad_dtype_b or st.b.dtype, grad_scale)
    if ds_work is not None and hasattr(ds_work, "wait"):
        ds_work.wait()
    return da, db, ds
