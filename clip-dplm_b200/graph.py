"""`GraphedClipStep`: the fused loss forward + backward of one fixed shape captured in a CUDA graph.

One training step of the hot path is ~25 launches (the three contraction kernels, the HBM-bound helpers around them
and, row-sharded, the exchange kernels and copy-engine copies of exchange.py).  At the headline size on one GPU their
launch gaps are ~1 % of the step; on 8 GPUs, where the same work takes 2.4 ms, launch latency and Python would become a
fifth of it.  Because the logit scale is read on the device (``scale_dev``, include/clipnce.h) and the exchange keeps its
barrier epochs on the device, nothing in the step needs the host, so the whole step -- exchange included -- replays as
one graph:

    step = GraphedClipStep(n_local, d, group=group)        # captures once (after eager warm-up steps)
    loss, d_a, d_b, d_logit_scale = step(a, b, logit_scale)

``a``, ``b`` ([n_local, d], the step's dtype) and ``logit_scale`` (0-d) are copied into the graph's static inputs; the
returned tensors are the graph's static outputs (valid until the next call).  The kernel family (fixed shift vs true
running maxima, include/clipnce.h) is chosen from the logit scale seen at capture time; every ``family_check_every``
replays the step compares it with the family the current scale selects (a non-blocking read, functional.scale_family) and
re-captures by itself when exp(logit_scale) has crossed 40 -- a frozen fixed-shift graph would underflow beyond 43.
"""
from __future__ import annotations

from typing import Optional

import torch

from .engine import default_engine
from .functional import fused_clip_loss, scale_family


class GraphedClipStep:
    def __init__(self, n_local: int, d: int, *, dtype=torch.bfloat16, device=None, group=None, symmetric: bool = True,
                 scale_is_log: bool = True, clamp_max: Optional[float] = None, logit_scale_init: float = 2.6592,
                 engine=None, warmup: int = 3, split: bool = False):
        self.device = torch.device(device if device is not None else torch.cuda.current_device())
        self.group, self.engine = group, engine or default_engine()
        self.kw = dict(symmetric=symmetric, scale_is_log=scale_is_log, clamp_max=clamp_max, group=group, engine=self.engine)
        self.a = torch.zeros(n_local, d, dtype=dtype, device=self.device)
        self.b = torch.zeros(n_local, d, dtype=dtype, device=self.device)
        # a persistent leaf: the host-side scale hint (functional._ScaleHint) is keyed by the tensor object
        self.logit_scale = torch.full((), float(logit_scale_init), dtype=torch.float32, device=self.device,
                                      requires_grad=True)
        self.warmup = warmup
        # split: forward and backward are captured as TWO graphs replayed back to back, with an event recorded between
        # them -- `HostFedClipStep` starts the next batch's H2D behind it, so that the transfer overlaps the backward sweep
        # instead of the push / barrier / forward phase (measured on 4 GPUs: a barrier that takes 6-15 us takes 65 us while
        # a 32 MB H2D is in flight)
        self.split = bool(split)
        self.graph_b = None
        self.after_forward = torch.cuda.Event() if self.split else None
        self.graph = None
        self._primed = False
        self.family_check_every = 16 if group is None else 64   # with a group the check is a (tiny, synchronising) all-reduce
        self._family = None
        self._replays = 0
        self._fam_args = (scale_is_log, clamp_max, dtype, d)

    def _eager(self):
        a = self.a.detach().requires_grad_(True)
        b = self.b.detach().requires_grad_(True)
        t = self.logit_scale
        loss = fused_clip_loss(a, b, t, **self.kw)
        d_a, d_b, d_t = torch.autograd.grad(loss, (a, b, t))
        return loss.detach(), d_a, d_b, d_t

    def recapture(self):
        """(Re)build the graph from the current contents of the static inputs."""
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):   # allocator, kernel attributes, NCCL communicators, the scale hint
                self._eager()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self._family = self._current_family()
        self.graph = torch.cuda.CUDAGraph()
        if not self.split:
            with torch.cuda.graph(self.graph):
                self.out = self._eager()
            return self
        self.graph_b = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            a = self.a.detach().requires_grad_(True)
            b = self.b.detach().requires_grad_(True)
            loss = fused_clip_loss(a, b, self.logit_scale, **self.kw)
        with torch.cuda.graph(self.graph_b, pool=self.graph.pool()):   # what the forward saved lives in the shared pool
            d_a, d_b, d_t = torch.autograd.grad(loss, (a, b, self.logit_scale))
        self.out = (loss.detach(), d_a, d_b, d_t)
        return self

    def _current_family(self):
        sil, cm, dt, d = self._fam_args
        fam = scale_family(self.logit_scale, sil, cm, dt, d, 0, self.engine)
        if self.group is not None:
            # ranks read their (possibly stale) copies of s at different moments: agree on one answer, or some ranks would
            # re-capture -- eager steps with barriers in them -- while others replay
            import torch.distributed as dist
            t = torch.tensor([fam], dtype=torch.int32, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            fam = int(t.item())
        return fam

    def close(self):
        """Drop the captured graph.  Call before ``dist.destroy_process_group()``: a live graph that holds captured NCCL
        kernels (the collectives baseline) stalls the communicator's teardown."""
        if self.graph is not None:
            torch.cuda.synchronize(self.device)
            if self.graph_b is not None:
                self.graph_b.reset()
                self.graph_b = None
            self.graph.reset()
            self.graph = None

    def replay(self):
        """Replay on whatever the static inputs ``self.a``, ``self.b``, ``self.logit_scale`` hold (callers that write their
        embeddings straight into them -- e.g. an H2D copy -- skip the device-to-device copies of ``__call__``)."""
        if self.graph is None:
            self.recapture()
        self._replays += 1
        if self.family_check_every and self._replays % self.family_check_every == 0 and self._current_family() != self._family:
            self.recapture()      # the scale left the captured kernel family's range (same decision on every rank's schedule)
        self.graph.replay()
        if self.split:
            self.after_forward.record(torch.cuda.current_stream(self.device))
            self.graph_b.replay()
        return self.out

    def __call__(self, a, b, logit_scale):
        self.a.copy_(a, non_blocking=True)
        self.b.copy_(b, non_blocking=True)
        with torch.no_grad():
            if torch.is_tensor(logit_scale):
                self.logit_scale.copy_(logit_scale.detach(), non_blocking=True)
            else:
                self.logit_scale.fill_(float(logit_scale))
        if self.graph is None:
            # capture with REAL inputs in the static buffers: the row norms etc. of the warm-up steps are then finite
            self.recapture()
        return self.replay()


class HostFedClipStep:
    """A `GraphedClipStep` fed from pinned HOST buffers with the copies pipelined: while the graph of step k runs, the
    embeddings of step k+1 cross PCIe on a copy stream into the other of two staging buffers.

        feeder = HostFedClipStep(n_local, d, group=group)
        for a_host, b_host in batches:                  # pinned [n_local, d] tensors
            loss, d_a, d_b, d_t = feeder.step(a_host, b_host)   # results of THIS batch (device tensors)

    `step` enqueues this batch's H2D (unless `prefetch` already did), waits for it on the compute stream, copies the
    staging buffers into the graph's static inputs (device to device) and replays the graph; call
    `prefetch(next_a_host, next_b_host)` right after `step` to overlap the next batch's transfer with this step.
    """

    def __init__(self, n_local: Optional[int] = None, d: Optional[int] = None, *, inner: Optional[GraphedClipStep] = None,
                 **kw):
        if inner is None and kw.get("group") is not None:
            # row-sharded over 4+ ranks: keep the H2D away from the exchange phase (GraphedClipStep.split); with fewer
            # ranks the per-rank transfer is large and is better hidden beside the forward (measured)
            import torch.distributed as dist
            kw.setdefault("split", dist.get_world_size(kw["group"]) >= 4)
        self.inner = inner if inner is not None else GraphedClipStep(n_local, d, **kw)
        dev = self.inner.device
        self.stage = [(torch.empty_like(self.inner.a), torch.empty_like(self.inner.b)) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]     # H2D into stage[i] done
        self.free = [torch.cuda.Event(), torch.cuda.Event()]      # stage[i] consumed by the compute stream
        self.slot = 0
        self.prefetched = False
        for e in self.free:
            e.record()

    @property
    def logit_scale(self):
        return self.inner.logit_scale

    def prefetch(self, a_host, b_host):
        """Start the H2D of the NEXT batch (returns immediately)."""
        i = self.slot
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[i])
            if getattr(self.inner, "split", False) and self.inner._replays > 0:
                # not before the forward of the step just launched has finished: the copy then runs beside the backward
                self.copy_stream.wait_event(self.inner.after_forward)
            self.stage[i][0].copy_(a_host, non_blocking=True)
            self.stage[i][1].copy_(b_host, non_blocking=True)
            self.ready[i].record(self.copy_stream)
        self.prefetched = True

    def step(self, a_host=None, b_host=None):
        if not self.prefetched:
            self.prefetch(a_host, b_host)
        i = self.slot
        cur = torch.cuda.current_stream(self.inner.device)
        cur.wait_event(self.ready[i])
        self.inner.a.copy_(self.stage[i][0], non_blocking=True)
        self.inner.b.copy_(self.stage[i][1], non_blocking=True)
        self.free[i].record(cur)
        self.slot, self.prefetched = 1 - i, False
        return self.inner.replay()
