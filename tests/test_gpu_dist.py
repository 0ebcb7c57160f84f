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
