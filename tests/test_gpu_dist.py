"""Row-sharded global batch over NCCL (needs >= 2 GPUs; skipped on a single-GPU box): every rank's loss and
gradients must equal the single-process reference on the concatenated batch."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_step as O

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, d, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from clip_dplm_b200 import fused_clip_loss
        a, b = O.make_inputs(n, d, seed=33)
        nl = n // world
        ac = a[rank * nl:(rank + 1) * nl].cuda().bfloat16().requires_grad_(True)
        bc = b[rank * nl:(rank + 1) * nl].cuda().bfloat16().requires_grad_(True)
        t = torch.tensor(O.LOGIT_SCALE_INIT, device="cuda", requires_grad=True)
        loss = fused_clip_loss(ac, bc, t, group=dist.group.WORLD)
        loss.backward()
        torch.cuda.synchronize()
        q.put((rank, float(loss.detach()), ac.grad.float().cpu().numpy(), bc.grad.float().cpu().numpy(), float(t.grad)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_row_sharded_matches_single_process():
    world, n, d = 2, 1024, 256
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=300) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    a, b = O.make_inputs(n, d, seed=33)
    ref = O.ref_step(a.double(), b.double(), O.LOGIT_SCALE_INIT)
    nl = n // world
    rel = lambda x, r: float((torch.as_tensor(x).double() - r.double()).norm() / r.double().norm())
    for rank, loss, da, db, dt in out:
        assert abs(loss - float(ref["loss"])) <= 1e-3 * abs(float(ref["loss"]))
        assert rel(da, ref["d_a"][rank * nl:(rank + 1) * nl]) <= 2e-2
        assert rel(db, ref["d_b"][rank * nl:(rank + 1) * nl]) <= 2e-2
        assert abs(dt - float(ref["d_logit_scale"])) <= 2e-2 * abs(float(ref["d_logit_scale"]))


def _worker_topk_and_graph(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from clip_dplm_b200.graph import GraphedClipStep
        from clip_dplm_b200.retrieval import topk_similarity
        # (1) library row-sharded retrieval, candidates merged over NCCL
        g = torch.Generator().manual_seed(9)
        qs, lib = torch.randn(200, 128, generator=g).bfloat16(), torch.randn(3000, 128, generator=g).bfloat16()
        nl = lib.shape[0] // world
        s, i = topk_similarity(qs.cuda(), lib[rank * nl:(rank + 1) * nl].cuda(), 10, group=dist.group.WORLD)
        # (2) the row-sharded training step replayed as one CUDA graph (collectives captured)
        n, d = 1024, 256
        a, b = O.make_inputs(n, d, seed=33)
        ml = n // world
        step = GraphedClipStep(ml, d, group=dist.group.WORLD)
        out = None
        for ls in (O.LOGIT_SCALE_INIT, 2.2):
            loss, da, db, dt = step(a[rank * ml:(rank + 1) * ml].cuda().bfloat16(), b[rank * ml:(rank + 1) * ml].cuda().bfloat16(), ls)
            torch.cuda.synchronize()
            out = (float(loss), da.float().cpu().numpy(), db.float().cpu().numpy(), float(dt))
        q.put((rank, s.cpu().numpy(), i.cpu().numpy(), out))
        step.close()                # a live graph holding captured NCCL kernels would stall the communicator teardown
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_sharded_retrieval_and_graphed_step():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 27700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_topk_and_graph, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
        for p in procs:
            p.join(timeout=45)
            assert p.exitcode == 0
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    g = torch.Generator().manual_seed(9)
    qs, lib = torch.randn(200, 128, generator=g).bfloat16(), torch.randn(3000, 128, generator=g).bfloat16()
    s_ref, i_ref, sim = O.ref_topk(qs.double(), lib.double(), 10)
    a, b = O.make_inputs(1024, 256, seed=33)
    ref = O.ref_step(a.double(), b.double(), 2.2)
    rel = lambda x, r: float((torch.as_tensor(x).double() - r.double()).norm() / r.double().norm())
    for rank, s, i, (loss, da, db, dt) in out:
        assert torch.allclose(torch.from_numpy(s).double(), s_ref, atol=2e-5, rtol=0)
        assert torch.allclose(torch.gather(sim, 1, torch.from_numpy(i)), s_ref, atol=2e-5, rtol=0)
        assert abs(loss - float(ref["loss"])) <= 1e-3 * abs(float(ref["loss"]))
        assert rel(da, ref["d_a"][rank * 512:(rank + 1) * 512]) <= 2e-2 and rel(db, ref["d_b"][rank * 512:(rank + 1) * 512]) <= 2e-2
        assert abs(dt - float(ref["d_logit_scale"])) <= 2e-2 * abs(float(ref["d_logit_scale"]))
