"""Sampled-row CPU oracle for shapes whose N x N logits do not fit a host (N = 4096 ... 262144).  TEST INFRASTRUCTURE ONLY
(same rule as ``oracle/ref_step.py``: imported by ``tests/``, ``bench.py``'s parity / cpu_baseline legs and tools only).

What the reference computes (old/clip.py:63-67, current/rna_clip_codes.ipynb:1948-1953, ``loss.backward()`` :2074) is
restated here in the form ``oracle.ref_step.closed_form`` pins against the reference, evaluated blockwise:

  * one pass over row chunks of the logits (fp32 GEMM per chunk, soft-max sums accumulated in float64 with a TRUE running
    maximum -- not the CUDA path's fixed shift) gives every row's and every column's log-sum-exp, the diagonal, the loss
    and ``sum G.S`` (the logit-scale gradient);
  * for a sample of row indices the gradient rows ``d_a[i]`` / ``d_b[j]`` are then formed in float64 against ALL N
    columns / rows: ``G_ij = (exp(S_ij - r_i) + exp(S_ij - c_j)) / 2N - delta_ij / N``, ``d a_hat_i = s sum_j G_ij b_hat_j``,
    followed by the normalise backward.

``test_cpu.py::test_sampled_oracle_matches_closed_form`` checks this module against ``closed_form`` (itself checked against
autograd of the reference's op sequence) on sizes where the dense form fits.
"""
from __future__ import annotations

import numpy as np
import torch

EPS = 1e-12  # F.normalize's clamp (old/clip.py:63-64)


def _normalized(x: torch.Tensor):
    x64 = x.detach().to(torch.float64)
    nrm = x64.norm(dim=1).clamp_min(EPS)
    return x64 / nrm[:, None], nrm


def full_statistics(a, b, scale: float, *, chunk: int = 2048, symmetric: bool = True):
    """a, b: [N,d] CPU tensors (the values the CUDA path saw), scale = s.  One blockwise pass over the N x N logits.

    -> dict(row_lse [N], col_lse [N], diag [N], loss, d_scale_sum) in float64 (numpy)."""
    ah, _ = _normalized(a)
    bh, _ = _normalized(b)
    n = ah.shape[0]
    ah32, bh32 = ah.float(), bh.float()
    row_lse = torch.empty(n, dtype=torch.float64)
    row_es = torch.zeros(n, dtype=torch.float64)          # sum_j P_ij S_ij
    col_m = torch.full((n,), -float("inf"), dtype=torch.float64)
    col_l = torch.zeros(n, dtype=torch.float64)
    col_t = torch.zeros(n, dtype=torch.float64)           # sum_i exp(S_ij - col_m_j) S_ij
    diag = (ah * bh).sum(dim=1) * scale
    for i0 in range(0, n, chunk):
        i1 = min(n, i0 + chunk)
        s_blk = torch.matmul(ah32[i0:i1], bh32.t()) * float(scale)            # [c, N] fp32
        m = s_blk.max(dim=1).values
        e = (s_blk - m[:, None]).exp_()
        l = e.sum(dim=1, dtype=torch.float64)
        row_lse[i0:i1] = m.double() + l.log()
        row_es[i0:i1] = (e * s_blk).sum(dim=1, dtype=torch.float64) / l
        cm = s_blk.max(dim=0).values.double()
        new_m = torch.maximum(col_m, cm)
        resc = (col_m - new_m).exp()
        resc[torch.isinf(col_m)] = 0.0
        e = (s_blk - new_m.float()[None, :]).exp_()
        col_l = col_l * resc + e.sum(dim=0, dtype=torch.float64)
        col_t = col_t * resc + (e * s_blk).sum(dim=0, dtype=torch.float64)
        col_m = new_m
    col_lse = col_m + col_l.log()
    if symmetric:
        loss = ((row_lse - diag).sum() + (col_lse - diag).sum()) / (2 * n)
        dss = (row_es.sum() + (col_t / col_l).sum()) / (2 * n) - diag.sum() / n
    else:
        loss = (row_lse - diag).sum() / n
        dss = row_es.sum() / n - diag.sum() / n
    return {"row_lse": row_lse.numpy(), "col_lse": col_lse.numpy(), "diag": diag.numpy(), "loss": float(loss),
            "d_scale_sum": float(dss)}


def _side_rows(x, y, scale, idx, lse_own, lse_other_all, symmetric, own_is_row_softmax):
    """Gradient rows of `x` at `idx` against all rows of `y` in float64.  lse_own[i]: the LSE over the swept index of the
    sampled rows; lse_other_all[j]: the LSE of every swept index's own soft-max (the other direction)."""
    xh, xn = _normalized(x)
    yh, _ = _normalized(y)
    n = yh.shape[0]
    idx_t = torch.as_tensor(np.asarray(idx), dtype=torch.long)
    s_blk = torch.matmul(xh[idx_t], yh.t()) * float(scale)                    # [k, N] float64
    own = torch.as_tensor(lse_own)[idx_t]
    other = torch.as_tensor(lse_other_all)
    if symmetric:
        g = ((s_blk - own[:, None]).exp() + (s_blk - other[None, :]).exp()) / (2 * n)
    elif own_is_row_softmax:
        g = (s_blk - own[:, None]).exp() / n
    else:   # one-directional loss: the swept index carries the soft-max
        g = (s_blk - other[None, :]).exp() / n
    g[torch.arange(len(idx_t)), idx_t] -= 1.0 / n
    d_hat = float(scale) * torch.matmul(g, yh)
    xs = xh[idx_t]
    d_x = (d_hat - xs * (xs * d_hat).sum(dim=1, keepdim=True)) / xn[idx_t][:, None]
    return d_x.numpy(), s_blk


def sampled_reference(a, b, scale: float, rows_a, rows_b, *, chunk: int = 2048, symmetric: bool = True, stats=None):
    """Everything the parity checks compare, for a sample of rows of both modalities.

    -> dict(loss, d_scale_sum, row_lse, col_lse, diag (full vectors), d_a [len(rows_a), d], d_b [len(rows_b), d])."""
    st = stats if stats is not None else full_statistics(a, b, scale, chunk=chunk, symmetric=symmetric)
    d_a, s_rows = _side_rows(a, b, scale, rows_a, st["row_lse"], st["col_lse"], symmetric, True)
    d_b, _ = _side_rows(b, a, scale, rows_b, st["col_lse"], st["row_lse"], symmetric, False)
    # self-check: the float64 LSE of the sampled rows against the blockwise fp32 pass
    m = s_rows.max(dim=1).values
    lse64 = (m + (s_rows - m[:, None]).exp().sum(dim=1).log()).numpy()
    dev = float(np.abs(lse64 - st["row_lse"][np.asarray(rows_a)]).max())
    if dev > 1e-4:
        raise AssertionError(f"sampled oracle: blockwise and float64 row LSE differ by {dev}")
    out = dict(st)
    out.update({"d_a": d_a, "d_b": d_b, "rows_a": np.asarray(rows_a), "rows_b": np.asarray(rows_b)})
    return out


def compare(ref, loss, d_a_rows, d_b_rows, d_scale_sum=None, *, loss_tol=1e-3, grad_tol=2e-2):
    """Relative errors of a step's outputs (its loss, its gradient rows at ref['rows_a'] / ref['rows_b']) -> dict with 'ok'."""
    def rel(x, r):
        x = np.asarray(x, np.float64)
        return float(np.linalg.norm(x - r) / max(np.linalg.norm(r), 1e-300))
    out = {"rows": int(len(ref["rows_a"])), "loss_rel": abs(float(loss) - ref["loss"]) / abs(ref["loss"]),
           "dA_rel": rel(d_a_rows, ref["d_a"]), "dB_rel": rel(d_b_rows, ref["d_b"]),
           "loss_tol": loss_tol, "grad_tol": grad_tol}
    ok = out["loss_rel"] <= loss_tol and out["dA_rel"] <= grad_tol and out["dB_rel"] <= grad_tol
    if d_scale_sum is not None:
        out["dscale_rel"] = abs(float(d_scale_sum) - ref["d_scale_sum"]) / max(abs(ref["d_scale_sum"]), 1e-300)
        ok = ok and out["dscale_rel"] <= grad_tol
    out["ok"] = bool(ok)
    return out
