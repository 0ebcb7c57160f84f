"""CPU oracle for the CLIP / InfoNCE hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product path
(``clip_dplm_b200``) never does; it fails loudly when the CUDA library is missing.

Parity status: the reference's own tests pin nothing on this path (its
``tong/tests`` scripts hold no assertions), so the oracle is pinned the other
way round: ``oracle/gen_golden.py`` imports the reference's modules from
``/root/reference`` in the build container, asserts that every function below
reproduces them bit-for-bit (same torch ops in the same order), and commits the
resulting vectors under ``tests/golden/``.  The arithmetic itself lives in an
un-pinned third-party dependency (PyTorch ATen: normalize / mm / log_softmax /
nll_loss; the reference ships no requirements file) -- the engine used here is
this image's torch 2.11.0 CPU.

Every function cites the reference lines it restates (paths relative to the
reference checkout).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

LOGIT_SCALE_INIT = math.log(1.0 / 0.07)  # run1/configuration_hybrid_clip.py:100, rna_clip_codes.ipynb:1933


def effective_scale(logit_scale: torch.Tensor, scale_is_log: bool = True, clamp_max=None) -> torch.Tensor:
    """s = exp(t) (old/clip.py:66) optionally .clamp(max=100) (old/clip_opt.py:100);
    tong divides by a fixed temperature instead (tong/utils/losses.py:14) -> s = 1/temperature."""
    s = logit_scale.exp() if scale_is_log else logit_scale
    if clamp_max is not None:
        s = s.clamp(max=clamp_max)
    return s


def ref_logits(a, b, logit_scale, scale_is_log=True, clamp_max=None):
    """old/clip.py:63-67 (== :100-104, run1/full.py:47-50, rna_clip_codes.ipynb:1948-1951)."""
    a_hat = F.normalize(a, dim=-1)
    b_hat = F.normalize(b, dim=-1)
    s = effective_scale(logit_scale, scale_is_log, clamp_max)
    return torch.matmul(a_hat, b_hat.t()) * s, a_hat, b_hat


def ref_loss(a, b, logit_scale, *, symmetric=True, scale_is_log=True, clamp_max=None, extra_cols=None):
    """Loss of the hot path, op for op.

    symmetric=True, extra_cols=None : rna_clip_codes.ipynb:1948-1953 and each pair of
                                      tf_clip_codes (1).ipynb:13146-13165.
    symmetric=False                 : run1/full.py:132-133, old/ablation.py:16 (one-directional CE).
    extra_cols (already normalised) : old/clip_opt.py:115-121 + :136-149  ->  (CE([S | S_c]) + CE(S^T)) / 2.
    symmetric=False + extra_cols    : tong/utils/losses.py:4-19 (queue rows are appended to the
                                      columns AFTER y is normalised (:10-11), i.e. used as stored).
    """
    sim, a_hat, b_hat = ref_logits(a, b, logit_scale, scale_is_log, clamp_max)
    n = sim.size(0)
    labels = torch.arange(n, device=sim.device)
    rows = sim
    if extra_cols is not None:
        e = extra_cols.detach()
        s = effective_scale(logit_scale, scale_is_log, clamp_max)
        rows = torch.cat([sim, torch.matmul(a_hat, e.t()) * s], dim=1)
    loss_rows = F.cross_entropy(rows, labels)
    if not symmetric:
        return loss_rows
    loss_cols = F.cross_entropy(sim.t(), labels)
    return (loss_rows + loss_cols) / 2


def ref_step(a, b, logit_scale, **kw):
    """Forward + backward of the hot path (rna_clip_codes.ipynb:2074 ``loss.backward()``).

    Returns dict(loss, d_a, d_b, d_logit_scale) as float64/float32 tensors matching ``a``'s dtype.
    """
    a = a.detach().clone().requires_grad_(True)
    b = b.detach().clone().requires_grad_(True)
    t = torch.as_tensor(logit_scale, dtype=a.dtype).detach().clone().requires_grad_(True)
    loss = ref_loss(a, b, t, **kw)
    loss.backward()
    return {"loss": loss.detach(), "d_a": a.grad, "d_b": b.grad,
            "d_logit_scale": t.grad if t.grad is not None else torch.zeros_like(t)}


# ----------------------------------------------------------------------------------------------
# Closed-form restatement (numpy float64).  Used to check the *intermediate* quantities the CUDA
# path exposes through the C-ABI (rinv, row/col LSE, diagonal) and the gradient formulas the
# backward kernel implements.
# ----------------------------------------------------------------------------------------------

def closed_form(a: np.ndarray, b: np.ndarray, scale: float, *, symmetric=True, n_pos_cols=None,
                row_offset=0, n_global=None, eps=1e-12):
    """float64 closed form.  ``a``: [n_rows,d], ``b``: [n_cols,d] raw embeddings, ``scale`` = s.

    Columns >= n_pos_cols (hard-negative cache / queue) carry no positives and take part in the
    row soft-max only.  ``row_offset`` is the global index of local row 0 (multi-GPU row shard),
    ``n_global`` the global batch the mean is taken over.
    """
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    n_rows, n_cols = a.shape[0], b.shape[0]
    n_pos = n_cols if n_pos_cols is None else n_pos_cols
    n_glob = n_rows if n_global is None else n_global
    na = np.maximum(np.linalg.norm(a, axis=1), eps)
    nb = np.maximum(np.linalg.norm(b, axis=1), eps)
    ah, bh = a / na[:, None], b / nb[:, None]
    S = scale * (ah @ bh.T)
    m_r = S.max(axis=1)
    row_lse = m_r + np.log(np.exp(S - m_r[:, None]).sum(axis=1))
    Sp = S[:, :n_pos]
    m_c = Sp.max(axis=0)
    col_lse = m_c + np.log(np.exp(Sp - m_c[None, :]).sum(axis=0))   # over the LOCAL rows only
    idx = np.arange(n_rows)
    diag = S[idx, idx + row_offset]
    P_r = np.exp(S - row_lse[:, None])
    G = np.zeros_like(S)
    if symmetric:
        P_c = np.exp(Sp - col_lse[None, :])
        G += P_r / (2 * n_glob)
        G[:, :n_pos] += P_c / (2 * n_glob)
        G[idx, idx + row_offset] -= 1.0 / n_glob
        loss = ((row_lse - diag).sum() + (col_lse[row_offset:row_offset + n_rows] - diag).sum()) / (2 * n_glob)
    else:
        G += P_r / n_glob
        G[idx, idx + row_offset] -= 1.0 / n_glob
        loss = (row_lse - diag).sum() / n_glob
    d_ah = scale * (G @ bh)
    d_bh = scale * (G.T @ ah)
    d_scale_sum = float((G * S).sum())          # = dL/dt when s = exp(t) un-clamped; dL/ds = this / s
    d_a = (d_ah - ah * (ah * d_ah).sum(axis=1, keepdims=True)) / na[:, None]
    d_b = (d_bh - bh * (bh * d_bh).sum(axis=1, keepdims=True)) / nb[:, None]
    return {"rinv_a": 1.0 / na, "rinv_b": 1.0 / nb, "a_hat": ah, "b_hat": bh, "row_lse": row_lse,
            "col_lse": col_lse, "diag": diag, "loss": float(loss), "d_a_hat": d_ah, "d_b_hat": d_bh,
            "d_a": d_a, "d_b": d_b, "d_scale_sum": d_scale_sum}


# ----------------------------------------------------------------------------------------------
# Synthetic inputs shared by tests and bench.
# ----------------------------------------------------------------------------------------------

def make_inputs(n, d, seed=1234, correlated=True, dtype=torch.float32, round_bf16=True, n_cols=None, mix=0.5):
    """a = randn; b = mix*a + (1-mix)*randn (positives on the diagonal) or b = randn; values are rounded
    to bf16 once so the bf16 CUDA path and the fp32/fp64 reference see identical inputs."""
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(n, d, generator=g)
    m = n if n_cols is None else n_cols
    if correlated:
        nb = torch.randn(m, d, generator=g)
        b = nb.clone()
        k = min(n, m)
        b[:k] = mix * a[:k] + (1.0 - mix) * nb[:k]
    else:
        b = torch.randn(m, d, generator=g)
    if round_bf16:
        a = a.to(torch.bfloat16).to(torch.float32)
        b = b.to(torch.bfloat16).to(torch.float32)
    return a.to(dtype), b.to(dtype)


def retrieval_topk(queries, library, k=10):
    """Normalised similarity + top-k (run1/full.py:142-160: argmax at :152, cosine_similarity at :157)."""
    q = F.normalize(queries, dim=-1)
    l = F.normalize(library, dim=-1)
    sim = torch.matmul(q, l.t())
    return torch.topk(sim, k, dim=1)


def ref_topk(q, lib, k):
    """Top-k retrieval the way the reference's evaluation forms it (run1/full.py:157 cosine similarity of every pair,
    :152 argmax over it) -- as a dense normalised matmul (identical values; the reference's broadcast builds an
    [n_q, n_lib, d] intermediate).  -> (scores [n_q,k], indices [n_q,k], sim [n_q,n_lib])."""
    import torch.nn.functional as F
    sim = F.normalize(q, dim=-1) @ F.normalize(lib, dim=-1).t()
    scores, idx = torch.topk(sim, min(k, lib.shape[0]), dim=1)
    return scores, idx, sim
