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


class PeerGradSync:
    """Gradient all-reduce over NVLink peer memory (csrc/peer.cu), a drop-in for GradSync in d_update / g_update / core_step.

    Each net's flat gradient buffer is allocated in symmetric memory (torch.distributed._symmetric_memory: every rank maps every
    rank's buffer; plumbing only) and the plan writes its parameter gradients straight into it (NetPlan.set_grad_buffer:
    ``p.grad`` are persistent views, no per-step copy).  The reduction itself is the library's two-shot kernel -- barrier,
    reduce own range from all peers + push the sums to all peers, barrier -- three plain launches on the step's stream, so
    the whole data-parallel step (collectives included) is captured as ONE CUDA graph (``capturable``)."""

    capturable = True

    def __init__(self, modules, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _abi as A
        self._A = A
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        dev = torch.device("cuda", torch.cuda.current_device())
        try:
            symm.enable_symm_mem_for_group(self.group.group_name)
        except Exception:                                         # noqa: BLE001 -- newer torch enables groups on rendezvous
            pass
        def mapped(t):
            """device array of `world` pointers: where each rank's copy of the symmetric tensor `t` is mapped in THIS process"""
            hdl = symm.rendezvous(t, self.group)
            off = int(getattr(hdl, "offset", 0) or 0)
            ptrs = [int(p) + off for p in hdl.buffer_ptrs]
            if ptrs[self.rank] != t.data_ptr():
                raise RuntimeError("symmetric memory: local mapping %#x != tensor %#x" % (ptrs[self.rank], t.data_ptr()))
            return hdl, torch.tensor(ptrs, dtype=torch.int64, device=dev)

        self._sig = symm.empty(64, dtype=torch.int32, device=dev)
        self._sig.zero_()
        self._sig_hdl, self._sig_ptrs = mapped(self._sig)
        self._nets = {}                                           # id(first parameter) -> (plan, buffer, handle, pointer table)
        for m in modules:
            plan = m._get_plan()
            buf = symm.empty(plan.ngrad, dtype=torch.float32, device=dev)
            hdl, ptrs = mapped(buf)
            plan.set_grad_buffer(buf)
            self._nets[id(plan.params[0])] = (plan, buf, hdl, ptrs)
        self.bytes_last = 0
        self._checked = set()
        torch.cuda.synchronize()
        dist.barrier(self.group)

    def __call__(self, params):
        if self.world == 1:
            return 1.0
        params = list(params)
        plan, buf, _, ptrs = self._nets[id(params[0])]
        if id(plan) not in self._checked:                         # once: autograd must have kept our views (it does not copy them)
            lo, hi = buf.data_ptr(), buf.data_ptr() + buf.numel() * 4
            for p in params:
                if p.grad is not None and not (lo <= p.grad.data_ptr() < hi):
                    raise RuntimeError("PeerGradSync: a gradient does not live in the symmetric buffer (was p.grad set by hand?)")
            self._checked.add(id(plan))
        A = self._A
        A.call("ag_peer_allreduce", ptrs.data_ptr(), self._sig_ptrs.data_ptr(), self.rank, self.world, buf.numel(), 0, A.stream())
        self.bytes_last = buf.numel() * 4
        return 1.0 / self.world


def broadcast_parameters(modules, src=0):
    """Identical initial parameters on every rank (broadcast once; never per call as DataParallel does)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for m in modules:
        for p in m.parameters():
            dist.broadcast(p.data, src)
            p._ag_epoch = getattr(p, "_ag_epoch", 0) + 1      # written through .data: invalidate the packed-operand cache


# --------------------------------------------------------------------------------------------------------------------
# Small reductions that keep single-process semantics: the accuracy gates' counters, calc_dists' batch
# statistics, the NaN / overflow flag of check_grad
# --------------------------------------------------------------------------------------------------------------------
def reduce_stats(*stats, group=None):
    """Sum [correct, num]-style counters (train.masked_bce_mean's `stats`, audiogan.py:741-742, :781-782, :813) over the ranks,
    in ONE all-reduce; returns the reduced tensors.  Accuracy = correct / num is then that of the global minibatch."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return stats if len(stats) > 1 else stats[0]
    flat = torch.cat([s.reshape(-1).float() for s in stats])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    out, off = [], 0
    for s in stats:
        out.append(flat[off:off + s.numel()].view_as(s))
        off += s.numel()
    return tuple(out) if len(out) > 1 else out[0]


def any_rank(flag, group=None):
    """True on every rank when `flag` (bool / 0-dim tensor) is set on ANY rank: check_grad's NaN / |g| > 1e5 verdict and the
    generator's early-exit decision must be taken by all ranks together, or they diverge at the next collective."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return bool(flag)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([1.0 if bool(flag) else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return bool(t.item() > 0)


class _GatherBatch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, dim, group):
        world = dist.get_world_size(group)
        parts = [torch.empty_like(q) for _ in range(world)]
        dist.all_gather(parts, q.contiguous(), group=group)
        ctx.dim, ctx.n, ctx.rank, ctx.world = dim, q.shape[dim], dist.get_rank(group), world
        return torch.cat(parts, dim)

    @staticmethod
    def backward(ctx, g):
        # every rank evaluates the SAME function of the gathered tensor; its parameters only see the rank's own rows.  The
        # optimizer later divides the rank-summed gradient by `world` (mean of per-rank means for the per-sample loss terms),
        # so a term that is already a function of the global batch is multiplied by `world` here.
        return g.narrow(ctx.dim, ctx.rank * ctx.n, ctx.n) * float(ctx.world), None, None


def gather_batch(q, dim=1, group=None):
    """Differentiable all-gather of per-sample statistics along `dim` (equal shards): the hook train / modules.calc_dists take
    so that the feature-matching batch moments (audiogan.py:350-358) are those of the GLOBAL minibatch."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return q
    return _GatherBatch.apply(q, dim, group)
