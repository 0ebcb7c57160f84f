"""Retrieval benchmark (BASELINE.json config 5): 16k TF queries x a 1M-entry protein library, cosine top-10.

    python tools/bench_retrieval.py [--queries 16384] [--library 1000000] [--d 512] [--k 10] [--gpus N]

One process per GPU (torchrun for N > 1): the library is sharded by rows, every rank scans its shard with the tcgen05
pair kernel (running top-k per row in the epilogue), the [n_q, k] candidates are all-gathered and merged.  Prints one
JSON line: queries x library pairs scored per second, the sweep's share of the measured bf16 peak, and -- on one GPU --
the time torch needs for the same answer by materialising the similarity matrix in chunks (matmul + topk).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=16384)
    ap.add_argument("--library", type=int, default=1000000)
    ap.add_argument("--d", type=int, default=512)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    group = None
    if world > 1:
        dist.init_process_group("nccl")
        group = dist.group.WORLD
    from clip_dplm_b200.retrieval import topk_similarity
    dev = torch.device("cuda", local)
    n_lib = args.library // world
    g = torch.Generator(device=dev).manual_seed(7)
    q = torch.randn(args.queries, args.d, device=dev, generator=g).bfloat16()        # same queries on every rank
    g2 = torch.Generator(device=dev).manual_seed(100 + rank)
    lib = torch.randn(n_lib, args.d, device=dev, generator=g2).bfloat16()

    def run():
        return topk_similarity(q, lib, args.k, group=group)

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        s, i = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    ref_ms = None
    if world == 1:
        # what torch does for the same answer: chunked matmul + topk (the full [16k, 1M] fp32 matrix is 65 GB)
        qn = torch.nn.functional.normalize(q.float(), dim=-1).bfloat16()
        ln = torch.nn.functional.normalize(lib.float(), dim=-1).bfloat16()

        def ref():
            best_s = torch.full((args.queries, args.k), -2.0, device=dev)
            best_i = torch.zeros((args.queries, args.k), dtype=torch.int64, device=dev)
            for c0 in range(0, n_lib, 65536):
                sim = (qn @ ln[c0:c0 + 65536].t()).float()
                cs, ci = torch.topk(sim, args.k, dim=1)
                alls, alli = torch.cat([best_s, cs], 1), torch.cat([best_i, ci + c0], 1)
                best_s, pos = torch.topk(alls, args.k, dim=1)
                best_i = torch.gather(alli, 1, pos)
            return best_s, best_i

        ref()
        torch.cuda.synchronize()
        e0.record()
        rs, ri = ref()
        e1.record()
        torch.cuda.synchronize()
        ref_ms = e0.elapsed_time(e1)
        agree = float((ri == i).float().mean())
    if rank == 0:
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
            peak = float(peaks["bf16_tflops"])
        except Exception:
            peak = 1590.0
        flops = 2.0 * args.queries * args.library * args.d
        line = {"metric": "retrieval_pairs_scored_per_sec", "value": args.queries * args.library / (ms * 1e-3), "unit": "pairs/s",
                "n_gpus": world, "ms": ms, "config": {"workload": f"cosine top-{args.k}, {args.queries} queries x {args.library}-row library, d={args.d}, bf16",
                                                        "library_rows_per_gpu": n_lib},
                "roofline": {"bound": "tensor", "achieved": flops / world / (ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                             "frac": flops / world / (ms * 1e-3) / 1e12 / peak}}
        if ref_ms is not None:
            line["torch_chunked_matmul_topk_ms"] = ref_ms
            line["index_agreement_with_torch_bf16_normalised"] = agree
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
