"""Several bench.py configurations in ONE launch (one process group, one CUDA context per rank).

A torchrun start costs ~20 s per process before the first kernel (imports, NCCL init, symmetric-memory rendezvous), and GPU
time on a multi-GPU box is charged per GPU: comparing two exchange implementations on 8 GPUs as two bench.py runs pays that
start-up sixteen times.  This tool pays it once:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep.py \
        --configs 65536x512:link 65536x512:nccl 262144x512:link 32768x768:link --steps 20

Every configuration is `bench.run_ours` unchanged: same timing rules, the same JSON line on rank 0's stdout, in the order
of --configs (stderr carries a "# sweep: <config>" marker before each).  Not run on a GPU yet (written after the round's
GPU budget was spent); it is a measurement convenience, not part of the product path.
Format of a configuration:  <global batch>x<d>[:<comm>]   with comm in {auto, link, nccl}.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", nargs="+", required=True)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    import bench
    from clip_dplm_b200 import exchange

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        for cfg in args.configs:
            shape, _, comm = cfg.partition(":")
            n, d = (int(v) for v in shape.lower().split("x"))
            comm = comm or "auto"
            if comm not in ("auto", "link", "nccl"):
                raise SystemExit(f"sweep: unknown comm '{comm}' in '{cfg}'")
            if comm == "auto":
                os.environ.pop("CLIPNCE_COMM", None)
            else:
                os.environ["CLIPNCE_COMM"] = comm
            exchange.reset()     # forget the previous configuration's buffers and its link/nccl decision
            ns = argparse.Namespace(gpus=world, steps=args.steps, warmup=args.warmup, impl="ours", n=n, d=d, ref_rows=1024,
                                    no_cpu_baseline=True, no_graph=args.no_graph, timeline=None, trace=False, comm=comm)
            if rank == 0:
                print(f"# sweep: {cfg}", file=sys.stderr, flush=True)
            bench.run_ours(ns, rank, local_rank, world)
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
