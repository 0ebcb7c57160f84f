"""Projection-head tail fused in front of the loss.

``fused_linear_layernorm(h, linear, layernorm)`` replaces the last two layers of the reference's projection heads
(``nn.Linear(hidden, p)`` -> ``nn.LayerNorm(p)``: old/clip.py:26-33, old/clip_opt.py:16-44,
current/rna_clip_codes.ipynb:1901-1909) AND the ``F.normalize`` that follows them in every model (old/clip.py:63-64) by
one tcgen05 kernel (csrc/kernels_head.cuh, ``clipnce_head_tail``): it returns the bf16 rows the contrastive kernels read
together with their 1/norm, which ``fused_clip_loss(..., rinv_a=, rinv_b=)`` takes as is.  The [N, p] fp32 Linear output,
the LayerNorm output and the normalised copy never exist in HBM.

Backward: one row kernel (LayerNorm backward, ``clipnce_head_tail_backward``) and the two plain GEMMs dW = dz^T h,
dh = dz W through the library (cuBLAS), which is what they are.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .engine import default_engine


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def tail_is_served(hidden: int, width: int) -> bool:
    return hidden % 64 == 0 and hidden >= 64 and width in (128, 256, 384, 512)


class _LinearLayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, weight, bias, gamma, beta, eps):
        lib = default_engine().lib
        if not h.is_cuda:
            raise RuntimeError("clip_dplm_b200: fused_linear_layernorm needs CUDA tensors (there is no CPU path)")
        n, k = h.shape
        p = weight.shape[0]
        hb = h.detach().to(torch.bfloat16).contiguous()
        wb = weight.detach().to(torch.bfloat16).contiguous()
        g32 = gamma.detach().float().contiguous()
        b32 = beta.detach().float().contiguous()
        bias32 = bias.detach().float().contiguous() if bias is not None else None
        e = torch.empty((n, p), dtype=torch.bfloat16, device=h.device)
        zhat = torch.empty((n, p), dtype=torch.bfloat16, device=h.device)
        rstd = torch.empty(n, dtype=torch.float32, device=h.device)
        rinv = torch.empty(n, dtype=torch.float32, device=h.device)
        with torch.cuda.device(h.device):
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(lib.clipnce_head_tail(_p(hb), _p(wb), _p(bias32), _p(g32), _p(b32), n, k, p, float(eps), _p(e), _p(zhat),
                                             _p(rstd), _p(rinv), st), "head_tail")
        ctx.save_for_backward(hb, wb, zhat, rstd, g32)
        ctx.meta = (h.dtype, weight.dtype, bias is not None, bias.dtype if bias is not None else None, gamma.dtype, beta.dtype)
        ctx.mark_non_differentiable(rinv)
        return e, rinv

    @staticmethod
    def backward(ctx, de, _g_rinv):
        hb, wb, zhat, rstd, g32 = ctx.saved_tensors
        h_dt, w_dt, has_bias, b_dt, g_dt, be_dt = ctx.meta
        lib = default_engine().lib
        n, p = zhat.shape
        de = de.contiguous()
        if de.dtype not in (torch.bfloat16, torch.float32):
            de = de.float()
        dz = torch.empty((n, p), dtype=torch.bfloat16, device=de.device)
        with torch.cuda.device(de.device):
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(lib.clipnce_head_tail_backward(_p(de), _lib.BF16 if de.dtype == torch.bfloat16 else _lib.F32, _p(zhat),
                                                      _p(rstd), _p(g32), n, p, _p(dz), st), "head_tail_backward")
        def32 = de.float()
        d_gamma = (def32 * zhat.float()).sum(dim=0).to(g_dt)
        d_beta = def32.sum(dim=0).to(be_dt)
        d_w = torch.matmul(dz.t(), hb).to(w_dt)        # [p, k]
        d_h = torch.matmul(dz, wb).to(h_dt)            # [n, k]
        d_b = dz.float().sum(dim=0).to(b_dt) if has_bias else None
        return d_h, d_w, d_b, d_gamma, d_beta, None


def fused_linear_layernorm(h, linear: torch.nn.Linear, layernorm: torch.nn.LayerNorm):
    """-> (e [N,p] bf16 = layernorm(linear(h)), rinv [N] f32 = 1 / max(|e|, 1e-12)) for a [N,hidden] CUDA tensor."""
    if h.dim() != 2:
        raise ValueError(f"expected a [N, hidden] tensor, got {tuple(h.shape)}")
    if not tail_is_served(linear.in_features, linear.out_features):
        raise RuntimeError(f"clip_dplm_b200: head tail {linear.in_features} -> {linear.out_features} is not served "
                           "(hidden % 64 == 0, width in {128, 256, 384, 512})")
    return _LinearLayerNorm.apply(h, linear.weight, linear.bias, layernorm.weight, layernorm.bias, layernorm.eps)
