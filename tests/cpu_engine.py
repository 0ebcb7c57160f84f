"""CPU stand-in for clip_dplm_b200.engine.CudaEngine -- TEST INFRASTRUCTURE ONLY.

Implements the stage contract of include/clipnce.h with dense torch ops (it materialises the logits,
which the product never does) so that the host-side orchestration in clip_dplm_b200/step.py --
sharding offsets, LSE combination, collectives, gradient routing -- can be exercised on CPU, including
world_size-2 gloo runs.  The product path never imports this file.
"""
from __future__ import annotations

import torch

EPS = 1e-12


class TorchCpuEngine:
    name = "cpu-test"

    def __init__(self, dtype=torch.float64):
        self.dt = dtype

    def uses_tensor_cores(self, dtype, d, scale, flags=0):
        return True   # so that step.py exercises the transposed-operand plumbing too

    def needs_transposed(self, dtype, d, scale, flags=0):
        return True

    def normalize(self, x, want_hat=None):
        xf = x.to(self.dt)
        denom = xf.norm(dim=1).clamp_min(EPS)
        xh = (xf / denom[:, None]).to(want_hat) if want_hat is not None else None
        return 1.0 / denom, xh

    def stage(self, x, c_dtype, want_t=False):
        xc = x.to(self.dt)
        xt = None
        if want_t:
            n = x.shape[0]
            ld = (n + 63) // 64 * 64
            xt = torch.zeros(x.shape[1], ld, dtype=self.dt)
            xt[:, :n] = xc.t()
        return xc, xt

    def topk(self, q, lib, rinv_q, rinv_lib, k, col_offset=0):
        sim = (q.to(self.dt) * rinv_q[:, None]) @ (lib.to(self.dt) * rinv_lib[:, None]).t()
        kk = min(k, lib.shape[0])
        s, i = torch.topk(sim, kk, dim=1)
        if kk < k:
            s = torch.cat([s, torch.full((s.shape[0], k - kk), float("-inf"), dtype=s.dtype)], dim=1)
            i = torch.cat([i - col_offset, torch.full((i.shape[0], k - kk), -1 - col_offset, dtype=i.dtype)], dim=1)
            return s.float(), i + col_offset
        return s.float(), i + col_offset

    def _logits(self, x, y, rinv_x, rinv_y, scale):
        return scale * ((x * rinv_x[:, None]) @ (y * rinv_y[:, None]).t())

    def forward(self, x, y, rinv_x, rinv_y, diag_offset, scale, flags=0, scale_dev=None):
        if scale_dev is not None:
            scale = float(scale_dev)
        S = self._logits(x, y, rinv_x, rinv_y, scale)
        n = x.shape[0]
        row_m = S.max(dim=1).values
        row_l = torch.exp(S - row_m[:, None]).sum(dim=1)
        col_m = S.max(dim=0).values
        col_l = torch.exp(S - col_m[None, :]).sum(dim=0)
        idx = torch.arange(n)
        diag = S[idx, idx + diag_offset]
        return row_m, row_l, col_m, col_l, diag

    def backward(self, x, y, y_t, rinv_x, rinv_y, diag_offset, scale, row_m, row_w, col_m, col_w, diag_w, grad_out,
                 flags=0, want_dscale=True, scale_dev=None):
        if scale_dev is not None:
            scale = float(scale_dev)
        if y_t is not None:   # the transposed operand must be the same matrix
            assert torch.equal(y_t[:, :y.shape[0]].t(), y)
        S = self._logits(x, y, rinv_x, rinv_y, scale)
        G = torch.exp(S - row_m[:, None]) * row_w[:, None]
        if col_w is not None:
            G = G + torch.exp(S - col_m[None, :]) * col_w[None, :]
        n_rows, n_cols = S.shape
        i = torch.arange(n_rows)
        j = i + diag_offset
        ok = (j >= 0) & (j < n_cols)
        G[i[ok], j[ok]] -= diag_w
        dx = grad_out * scale * (G @ (y * rinv_y[:, None]))
        ds = (grad_out * (G * S).sum()).reshape(1) if want_dscale else None
        return dx, ds

    def backward_dx(self, x, y, y_t, rinv_x, rinv_y, diag_offset, scale, row_m, row_w, col_m, col_w, diag_w, x_orig,
                    out_dtype, grad_scale=None, flags=0, want_dscale=True, scale_dev=None):
        dx_hat, ds = self.backward(x, y, y_t, rinv_x, rinv_y, diag_offset, scale, row_m, row_w, col_m, col_w, diag_w, 1.0,
                                   flags, want_dscale=want_dscale, scale_dev=scale_dev)
        return self.normalize_backward(x_orig, rinv_x, dx_hat, out_dtype, grad_scale), ds

    def softmax_weights(self, l, coef):
        return coef / l

    def combine_lse(self, m, l):
        return m + torch.log(l)

    def normalize_backward(self, x, rinv, dx_hat, out_dtype, grad_scale=None):
        xh = x.to(self.dt) * rinv[:, None]
        g = dx_hat
        dx = (g - xh * (xh * g).sum(dim=1, keepdim=True)) * rinv[:, None]
        if grad_scale is not None:
            dx = dx * grad_scale.to(self.dt)
        return dx.to(out_dtype)

    def loss(self, row_m, row_l, col_m, col_l, diag, diag_offset, n_global, symmetric):
        row_lse, col_lse = row_m + torch.log(row_l), col_m + torch.log(col_l)
        n = row_lse.numel()
        tot = (row_lse - diag).sum()
        if symmetric:
            tot = tot + (col_lse[diag_offset:diag_offset + n] - diag).sum()
        return (tot / ((2 if symmetric else 1) * n_global)).reshape(1)
