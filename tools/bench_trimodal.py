"""Tri-modal model's loss step (tf_clip_codes (1).ipynb:13146-13165): three symmetric InfoNCE pairs over three embeddings
sharing one logit_scale, forward + backward, bf16, one B200.  Every variant is replayed as ONE CUDA graph:

    grouped    modules.trimodal_contrastive_losses through the grouped launch (clipnce_group_*: one forward sweep, one
               backward sweep over all six sides, one finishing pass)
    three      the same function with the grouped launch switched off: three fused_clip_loss pair steps sharing row norms
               (the previous implementation)
    eager      the reference's lines in torch (bf16 autocast), three materialised [N,N] logits matrices

    python tools/bench_trimodal.py [--n 4096] [--d 512] [--steps 50]
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

from clip_dplm_b200 import functional as Fn  # noqa: E402
from clip_dplm_b200 import modules as M  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--d", type=int, default=512)
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    base = torch.randn(args.n, args.d, device=dev)
    embs = [(m * base + (1 - m) * torch.randn(args.n, args.d, device=dev)).bfloat16().requires_grad_(True) for m in (1.0, 0.5, 0.3)]
    ls = torch.nn.Parameter(torch.tensor(math.log(1 / 0.07), device=dev))

    def step(kind):
        for e in embs:
            e.grad = None
        ls.grad = None
        if kind == "eager":
            c, p, q = (F.normalize(e, dim=-1) for e in embs)
            s = ls.exp()
            lab = torch.arange(args.n, device=dev)
            loss = 0
            for x, y in ((c, p), (c, q), (p, q)):
                sim = torch.matmul(x, y.t()) * s
                loss = loss + (F.cross_entropy(sim, lab) + F.cross_entropy(sim.t(), lab)) / 2
        else:
            Fn.GROUP_MAX_ROWS = (1 << 30) if kind == "grouped" else 0
            loss = M.trimodal_contrastive_losses(*embs, ls)["loss"]
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
    for kind in ("grouped", "three", "eager"):
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
        out[kind] = {"graph_ms": ms, "pairs_per_s": 3 * args.n / (ms * 1e-3), "eager_launch_ms": ms_eager,
                     "loss": float(loss.detach()), "grad_norm_cell": float(embs[0].grad.float().norm())}
    Fn.GROUP_MAX_ROWS = 8192
    line = {"config": f"tri-modal loss step: 3 symmetric InfoNCE pairs over 3 embeddings [{args.n}, {args.d}] bf16, one logit_scale, "
                      "forward + backward, 1 B200, one CUDA graph per step", "n": args.n, "d": args.d, "steps": args.steps, **out,
            "grouped_over_three": out["three"]["graph_ms"] / out["grouped"]["graph_ms"],
            "grouped_over_eager": out["eager"]["graph_ms"] / out["grouped"]["graph_ms"]}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
