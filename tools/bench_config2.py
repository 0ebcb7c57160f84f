"""BASELINE.json config 2: protein-TF CLIP step with frozen ESM-2 650M-width inputs (1280 -> 512 projection heads), global
batch 4096, bf16, one B200.  Heads AND loss inside the timed region (forward + backward), CUDA events.  At this size the
device work is ~0.1 ms, so every variant is replayed as ONE CUDA graph (`graph_ms`; eager launches are host-bound and are
reported beside it as `eager_launch_ms`):

    fused   ProjectionHead with the fused tail (Linear -> LayerNorm -> row norm in one tcgen05 kernel) + fused_clip_loss
    mixed   torch heads (fuse_tail off)                                                                + fused_clip_loss
    eager   torch heads + the reference's loss lines (old/clip.py:63-67, rna_clip_codes.ipynb:1952-1953), logits materialised

    python tools/bench_config2.py [--n 4096] [--width 1280] [--p 512] [--steps 50]
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from clip_dplm_b200 import fused_clip_loss  # noqa: E402
from clip_dplm_b200 import modules as M  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--p", type=int, default=512)
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    heads = [M.ProjectionHead(args.width, args.p, hidden_dim=2 * args.p).to(dev).bfloat16() for _ in range(2)]
    for h in heads:
        for m in h.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    ls = torch.nn.Parameter(torch.tensor(math.log(1 / 0.07), device=dev))
    xa = torch.randn(args.n, args.width, device=dev).bfloat16()      # frozen-LM embeddings (ESM-2 650M width)
    xb = (0.5 * xa.float() + 0.5 * torch.randn(args.n, args.width, device=dev)).bfloat16()
    params = [p for h in heads for p in h.parameters()] + [ls]

    def step(kind):
        for p in params:
            p.grad = None
        for h in heads:
            h.fuse_tail, h.fuse_tail_min_rows = kind == "fused", 1
        ea, eb = heads[0](xa), heads[1](xb)
        if kind == "eager":
            a, b = F.normalize(ea, dim=-1), F.normalize(eb, dim=-1)
            sim = torch.matmul(a, b.t()) * ls.exp()
            lab = torch.arange(args.n, device=dev)
            loss = (F.cross_entropy(sim, lab) + F.cross_entropy(sim.t(), lab)) / 2
        else:
            loss = fused_clip_loss(ea, eb, ls, rinv_a=getattr(ea, "_clipnce_rinv", None), rinv_b=getattr(eb, "_clipnce_rinv", None))
        loss.backward()
        return loss

    def timed(fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    out = {}
    for kind in ("fused", "mixed", "eager"):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(5):
                loss = step(kind)
        torch.cuda.current_stream().wait_stream(side)
        ms_eager = timed(lambda: step(kind))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = step(kind)
        for _ in range(3):
            graph.replay()
        ms = timed(graph.replay)
        out[kind] = {"graph_ms": ms, "pairs_per_s": args.n / (ms * 1e-3), "eager_launch_ms": ms_eager, "loss": float(loss.detach())}
    line = {"config": f"BASELINE config 2: heads {args.width}->{2 * args.p}->{args.p} x2 + symmetric InfoNCE, global batch {args.n}, bf16, 1 B200, "
                      "forward + backward, one CUDA graph per step", "n": args.n, "steps": args.steps, **out,
            "fused_over_eager": out["eager"]["graph_ms"] / out["fused"]["graph_ms"],
            "fused_over_mixed": out["mixed"]["graph_ms"] / out["fused"]["graph_ms"]}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
