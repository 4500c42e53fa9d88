"""Data-parallel plumbing: one process per GPU, batch sharded by rank, gradients all-reduced over NCCL.

Replaces the reference's per-call NN.DataParallel wrappers (audiogan.py:314, :379-410, :492-508: parameter
broadcast on every module call, LSTMs on GPU 0 only) with a bucketed all-reduce of each net's flat gradient
buffer between backward and the fused optimizer step.  The gradients of one network are views of a
single flat buffer (plan.pack_backward), so a bucket is a contiguous slice and nothing is copied.
"""
import os

import torch
import torch.distributed as dist


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_batch(global_batch, rank, world):
    """Equal per-rank share of the minibatch (mean of per-rank means == global mean)."""
    if global_batch % world:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    return rank * per, per


def _flat_base(grads):
    """The common flat buffer the gradients are views of, when they tile it in order; else None."""
    base = getattr(grads[0], "_base", None)
    if base is None or base.dim() != 1:
        return None
    for g in grads:
        if getattr(g, "_base", None) is not base or not g.is_contiguous():
            return None
    return base


class GradSync:
    """Callable handed to d_update / g_update: sums the gradients over ranks in `nbuckets` asynchronous
    all-reduces and returns the scale (1/world) the fused optimizer applies while it reads them."""

    def __init__(self, nbuckets=4, group=None):
        self.nbuckets, self.group = nbuckets, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.bytes_last = 0
        self._pending = {}           # id(plan) -> ([async works], [(a, b) spans already being reduced])
        self._done = set()           # id(param) whose gradient came out of a packed buffer that was reduced during backward

    # -- overlap with backward: the engine reports finished weight-gradient regions of plan.gpflat (NetPlan.reduce_span),
    #    each becomes one asynchronous all-reduce bucket that runs while the rest of backward executes -------------------
    def attach(self, *modules):
        """Let these Generator / Discriminator modules reduce their packed weight gradients during backward."""
        for m in modules:
            m._get_plan().early_sync = self
        return self

    def reduce_async(self, plan, a, b):
        works, spans = self._pending.setdefault(id(plan), ([], []))
        works.append(dist.all_reduce(plan.gpflat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        spans.append((a, b))

    def finish(self, plan):
        """Reduce whatever part of plan.gpflat has not been started yet, then make the current stream wait for all buckets."""
        works, spans = self._pending.pop(id(plan), ([], []))
        pos, n = 0, plan.gpflat.numel()
        for a, b in sorted(spans) + [(n, n)]:
            if a > pos:
                for c in plan.gpflat[pos:a].chunk(max(1, min(self.nbuckets, (a - pos) >> 20))):
                    works.append(dist.all_reduce(c, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            pos = max(pos, b)
        self.bytes_last = n * plan.gpflat.element_size()
        for w in works:
            w.wait()
        self._done.update(id(p) for p in getattr(plan, "params", ()))

    def __call__(self, params):
        if self.world == 1:
            return 1.0
        # gradients whose packed form was reduced during backward (finish) are already global sums
        params = list(params)
        grads = [p.grad for p in params if p.grad is not None and id(p) not in self._done]
        self._done.difference_update(id(p) for p in params)
        if not grads:
            return 1.0 / self.world
        flat = _flat_base(grads)
        copied = flat is None
        if copied:
            flat = torch.cat([g.reshape(-1) for g in grads])
        self.bytes_last = flat.numel() * flat.element_size()
        works = [dist.all_reduce(c, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                 for c in flat.chunk(self.nbuckets)]
        for w in works:
            w.wait()
        if copied:
            off = 0
            for g in grads:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        return 1.0 / self.world


def broadcast_parameters(modules, src=0):
    """Identical initial parameters on every rank (broadcast once; never per call as DataParallel does)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for m in modules:
        for p in m.parameters():
            dist.broadcast(p.data, src)
            p._ag_epoch = getattr(p, "_ag_epoch", 0) + 1      # written through .data: invalidate the packed-operand cache
