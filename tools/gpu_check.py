"""Stage-by-stage GPU bring-up checks (developer tool, not a test).  Each check runs in its own
process so a faulting kernel cannot poison the others:

    python tools/gpu_check.py            # run all checks, print a table
    python tools/gpu_check.py tc_fwd     # one check in this process
"""
from __future__ import annotations

import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(x, ref):
    import torch
    x, ref = x.double().cpu(), torch.as_tensor(ref).double().cpu()
    return float((x - ref).norm() / ref.norm().clamp_min(1e-300))


def setup(n, d, n_cols=None, scale=1 / 0.07, corr=True, seed=1234):
    import numpy as np
    import torch
    from oracle import ref_step as O
    a, b = O.make_inputs(n, d, seed=seed, correlated=corr, n_cols=n_cols)
    cf = O.closed_form(a.numpy(), b.numpy(), scale)
    return a, b, cf


def check_aux():
    import torch
    from clip_dplm_b200.engine import CudaEngine
    eng = CudaEngine()
    a, b, cf = setup(300, 192)
    for dt in (torch.bfloat16, torch.float32):
        x = a.cuda().to(dt)
        rinv, xh = eng.normalize(x, want_hat=dt)
        xc, xt = eng.stage(a.cuda(), dt, want_t=True)
        torch.cuda.synchronize()
        print(f"normalize {dt}: xhat rel {rel(xh, cf['a_hat']):.2e} rinv rel {rel(rinv, cf['rinv_a']):.2e} "
              f"stage ok {bool(torch.equal(xt[:, :300].t().contiguous(), xc))} conv ok {bool(torch.equal(xc, x))}")


def check_exact(dtype_name="f32"):
    import torch
    from clip_dplm_b200.engine import CudaEngine
    eng = CudaEngine()
    dt = torch.float32 if dtype_name == "f32" else torch.bfloat16
    for (n, d, scale) in [(256, 128, 14.29), (333, 192, 14.29), (200, 64, 100.0)]:
        a, b, cf = setup(n, d, scale=scale)
        ah, _ = eng.stage(a.cuda(), dt)
        bh, _ = eng.stage(b.cuda(), dt)
        ra, _ = eng.normalize(ah)
        rb, _ = eng.normalize(bh)
        row_lse, col_m, col_l, diag = eng.forward(ah, bh, ra, rb, 0, scale, flags=1)
        col_lse = eng.combine_lse(col_m, col_l)
        loss = eng.loss(row_lse, col_lse, diag, 0, n, True)
        torch.cuda.synchronize()
        print(f"exact[{dtype_name}] n={n} d={d} s={scale}: row_lse {rel(row_lse, cf['row_lse']):.2e} col_lse "
              f"{rel(col_lse, cf['col_lse']):.2e} diag {rel(diag, cf['diag']):.2e} loss {float(loss):.6f} vs {cf['loss']:.6f}")
        lc = -math.log(2 * n)
        lu, lv = eng.log_weights(row_lse, lc), eng.log_weights(col_lse, lc)
        da, ds = eng.backward(ah, bh, None, ra, rb, 0, scale, lu, lv, 1.0 / n, 1.0, flags=1)
        db, _ = eng.backward(bh, ah, None, rb, ra, 0, scale, lv, lu, 1.0 / n, 1.0, flags=1, want_dscale=False)
        torch.cuda.synchronize()
        print(f"    d_a_hat {rel(da, cf['d_a_hat']):.2e} d_b_hat {rel(db, cf['d_b_hat']):.2e} dscale {float(ds):.6e} vs "
              f"{cf['d_scale_sum']:.6e}")


def _tc_case(eng, n, d, scale, n_cols=None, corr=True, bwd=True):
    import torch
    a, b, cf = setup(n, d, n_cols=n_cols, scale=scale, corr=corr)
    m = b.shape[0]
    ah, aht = eng.stage(a.cuda().bfloat16(), torch.bfloat16, want_t=True)
    bh, bht = eng.stage(b.cuda().bfloat16(), torch.bfloat16, want_t=True)
    ra, _ = eng.normalize(ah)
    rb, _ = eng.normalize(bh)
    assert eng.uses_tensor_cores(torch.bfloat16, d, scale)
    row_lse, col_m, col_l, diag = eng.forward(ah, bh, ra, rb, 0, scale)
    col_lse = eng.combine_lse(col_m, col_l)
    torch.cuda.synchronize()
    msg = (f"tc n={n} m={m} d={d} s={scale:.2f}: row_lse {rel(row_lse, cf['row_lse']):.2e} col_lse "
           f"{rel(col_lse, cf['col_lse']):.2e} diag {rel(diag, cf['diag']):.2e}")
    if n == m:
        loss = eng.loss(row_lse, col_lse, diag, 0, n, True)
        torch.cuda.synchronize()
        msg += f" loss {float(loss):.6f} vs {cf['loss']:.6f}"
    print(msg, flush=True)
    if bwd and n == m:
        lc = -math.log(2 * n)
        lu = eng.log_weights(row_lse, lc)
        lv = eng.log_weights(col_lse, lc)
        da, ds = eng.backward(ah, bh, bht, ra, rb, 0, scale, lu, lv, 1.0 / n, 1.0)
        torch.cuda.synchronize()
        print(f"    d_a_hat {rel(da, cf['d_a_hat']):.2e} dscale {float(ds):.6e} vs {cf['d_scale_sum']:.6e}", flush=True)
        db, _ = eng.backward(bh, ah, aht, rb, ra, 0, scale, lv, lu, 1.0 / n, 1.0, want_dscale=False)
        torch.cuda.synchronize()
        print(f"    d_b_hat {rel(db, cf['d_b_hat']):.2e}", flush=True)


def check_tc_fwd():
    from clip_dplm_b200.engine import CudaEngine
    eng = CudaEngine()
    for (n, d) in [(128, 64), (128, 128), (256, 512), (333, 192), (1000, 512), (192, 768)]:
        _tc_case(eng, n, d, 1 / 0.07, bwd=False)
    _tc_case(eng, 200, 128, 10.0, n_cols=333, bwd=False)


def check_tc_bwd():
    from clip_dplm_b200.engine import CudaEngine
    eng = CudaEngine()
    for (n, d) in [(128, 64), (128, 128), (256, 512), (333, 192), (1000, 512), (192, 768)]:
        _tc_case(eng, n, d, 1 / 0.07, bwd=True)
    _tc_case(eng, 512, 256, 10.0, corr=False, bwd=True)


def check_tc_big():
    import torch
    from clip_dplm_b200.engine import CudaEngine
    eng = CudaEngine()
    n, d, scale = 20000, 512, 1 / 0.07      # exercises BLOCK_I=128 forward (n >= 128*148) and many tiles
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(n, d, device="cuda", generator=g).bfloat16()
    b = (0.5 * a.float() + 0.5 * torch.randn(n, d, device="cuda", generator=g)).bfloat16()
    ah, aht = eng.stage(a, torch.bfloat16, want_t=True)
    bh, bht = eng.stage(b, torch.bfloat16, want_t=True)
    ra, _ = eng.normalize(ah)
    rb, _ = eng.normalize(bh)
    for _ in range(2):
        t0 = time.time()
        row_lse, col_m, col_l, diag = eng.forward(ah, bh, ra, rb, 0, scale)
        torch.cuda.synchronize()
        print(f"    fwd wall {1e3 * (time.time() - t0):.2f} ms", flush=True)
    col_lse = eng.combine_lse(col_m, col_l)
    torch.cuda.synchronize()
    an, bn = ah.float() * ra[:, None], bh.float() * rb[:, None]
    S = scale * (an @ bn.t())
    r_ref, c_ref = torch.logsumexp(S, 1), torch.logsumexp(S, 0)
    print(f"tc big n={n}: row_lse {rel(row_lse, r_ref):.2e} col_lse {rel(col_lse, c_ref):.2e} diag {rel(diag, S.diagonal()):.2e}",
          flush=True)
    lc = -math.log(2 * n)
    lu, lv = lc - r_ref, lc - c_ref
    G = torch.exp(S + lu[:, None]) + torch.exp(S + lv[None, :])
    G.diagonal().sub_(1.0 / n)
    da_ref = scale * (G @ bn)
    db_ref = scale * (G.t() @ an)
    ds_ref = float((G * S).sum())
    del G, S
    for _ in range(2):
        t0 = time.time()
        da, ds = eng.backward(ah, bh, bht, ra, rb, 0, scale, lu.contiguous(), lv.contiguous(), 1.0 / n, 1.0)
        db, _ = eng.backward(bh, ah, aht, rb, ra, 0, scale, lv.contiguous(), lu.contiguous(), 1.0 / n, 1.0, want_dscale=False)
        torch.cuda.synchronize()
        print(f"    bwd wall {1e3 * (time.time() - t0):.2f} ms", flush=True)
    print(f"    d_a_hat {rel(da, da_ref):.2e} d_b_hat {rel(db, db_ref):.2e} dscale {float(ds):.5e} vs {ds_ref:.5e}", flush=True)


def check_e2e():
    import torch
    from clip_dplm_b200 import fused_clip_loss
    from oracle import ref_step as O
    for (n, d, kw) in [(256, 512, {}), (333, 192, {}), (256, 128, {"symmetric": False}),
                       (256, 128, {"clamp_max": 100.0, "ls": 5.0})]:
        kw = dict(kw)
        ls = kw.pop("ls", O.LOGIT_SCALE_INIT)
        a, b = O.make_inputs(n, d)
        ref = O.ref_step(a.double(), b.double(), ls, **kw)
        for cd in (torch.bfloat16, torch.float32):
            ac = a.cuda().to(cd).requires_grad_(True)
            bc = b.cuda().to(cd).requires_grad_(True)
            t = torch.tensor(ls, device="cuda", requires_grad=True)
            loss = fused_clip_loss(ac, bc, t, compute_dtype=cd, **kw)
            loss.backward()
            torch.cuda.synchronize()
            print(f"e2e n={n} d={d} {kw} {cd}: loss {float(loss):.6f} ref {float(ref['loss']):.6f} dA {rel(ac.grad, ref['d_a']):.2e} "
                  f"dB {rel(bc.grad, ref['d_b']):.2e} dt {float(t.grad):.5e} ref {float(ref['d_logit_scale']):.5e}", flush=True)


CHECKS = {"aux": check_aux, "exact_f32": lambda: check_exact("f32"), "exact_bf16": lambda: check_exact("bf16"),
          "tc_fwd": check_tc_fwd, "tc_bwd": check_tc_bwd, "tc_big": check_tc_big, "e2e": check_e2e}

if __name__ == "__main__":
    if len(sys.argv) > 1:
        for name in sys.argv[1:]:
            CHECKS[name]()
        sys.exit(0)
    for name in CHECKS:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True, timeout=300)
            out, code = r.stdout + r.stderr[-3000:], r.returncode
        except subprocess.TimeoutExpired as e:
            out, code = f"TIMEOUT\n{e.stdout}\n{e.stderr}", -9
        print(f"===== {name}: exit {code} ({time.time() - t0:.1f}s)\n{out}", flush=True)
