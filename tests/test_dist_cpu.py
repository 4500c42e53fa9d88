"""World-size-2 gloo tests (CPU) of the data-parallel host logic: batch sharding, flat-bucket gradient
all-reduce + 1/world scaling, parameter broadcast.  The kernels themselves need a GPU; the collective
plumbing does not."""
import os
import socket

import torch as T
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from audiogan_b200 import dist as agd
    r, w, _ = agd.init(backend="gloo")
    assert (r, w) == (rank, world)
    # batch sharding: equal shares that tile the global batch
    off, per = agd.shard_batch(8, rank, world)
    assert per == 4 and off == rank * 4
    try:
        agd.shard_batch(7, rank, world)
        raise AssertionError("expected ValueError")
    except ValueError:
        pass
    # parameters: identical after broadcast
    T.manual_seed(100 + rank)
    m = T.nn.Linear(5, 3)
    agd.broadcast_parameters([m])
    ref = [p.detach().clone() for p in m.parameters()]
    gathered = [T.zeros_like(ref[0]) for _ in range(world)]
    dist.all_gather(gathered, ref[0])
    assert T.equal(gathered[0], gathered[1])
    # gradients as views of ONE flat buffer (what plan.pack_backward produces): reduced in place, no copy
    flat = T.arange(18, dtype=T.float32) * (rank + 1)
    ps = list(m.parameters())
    ps[0].grad = flat[:15].view(3, 5)
    ps[1].grad = flat[15:18].view(3)
    sync = agd.GradSync(nbuckets=4)
    scale = sync(ps)
    assert scale == 0.5
    assert T.equal(flat, T.arange(18, dtype=T.float32) * 3)            # 1x + 2x summed over the two ranks
    assert ps[0].grad.data_ptr() == flat.data_ptr()
    # separately allocated gradients take the copy path and come back reduced
    for i, p in enumerate(ps):
        p.grad = T.full_like(p, float(rank + 1 + i))
    scale = sync(ps)
    for i, p in enumerate(ps):
        assert T.equal(p.grad, T.full_like(p, float(3 + 2 * i)))
    # overlap path: spans of a plan's packed gradient buffer reduced early (asynchronously), finish() reduces the rest once
    class _Plan:
        pass
    fp = _Plan()
    n = 2_500_000
    fp.gpflat = T.arange(n, dtype=T.float32) % 1000 * (rank + 1)
    sync2 = agd.GradSync(nbuckets=4)
    sync2.reduce_async(fp, 1_000_000, 1_500_000)
    sync2.reduce_async(fp, 100, 2000)
    sync2.finish(fp)
    assert T.equal(fp.gpflat, T.arange(n, dtype=T.float32) % 1000 * 3) and sync2.bytes_last == 4 * n
    sync2.finish(fp)                                                     # nothing started early: the whole buffer, in buckets
    assert T.equal(fp.gpflat, T.arange(n, dtype=T.float32) % 1000 * 6)
    # parameters of a plan whose packed gradients were reduced during backward are skipped once (still scaled by 1/world)
    fp.params = ps
    sync2.finish(fp)
    ps[0].grad, ps[1].grad = T.ones(3, 5), T.ones(3)
    assert sync2(ps) == 0.5 and T.equal(ps[0].grad, T.ones(3, 5))
    assert sync2(ps) == 0.5 and T.equal(ps[0].grad, T.full((3, 5), 2.0))      # the next call reduces again
    # mean-of-means == global mean with equal shards
    x = T.arange(8, dtype=T.float32)
    local_mean = x[off:off + per].mean().reshape(1)
    dist.all_reduce(local_mean)
    assert abs(float(local_mean) * scale - float(x.mean())) < 1e-6
    # semantic reductions: accuracy counters, NaN flag, batch statistics over the GLOBAL minibatch
    st_d, st_g = agd.reduce_stats(T.tensor([3.0 + rank, 10.0]), T.tensor([1.0, 4.0 + rank]))
    assert T.equal(st_d, T.tensor([7.0, 20.0])) and T.equal(st_g, T.tensor([2.0, 9.0]))
    assert agd.any_rank(rank == 1) is True and agd.any_rank(False) is False
    full = T.arange(24, dtype=T.float32).view(3, 4, 2) ** 1.5
    mine = full[:, rank * 2:(rank + 1) * 2].clone().requires_grad_(True)
    gq = agd.gather_batch(mine, dim=1)
    assert T.equal(gq.detach(), full)
    (gq.std(1) ** 2).sum().backward()                      # a function of the global batch, identical on both ranks
    ref = full.clone().requires_grad_(True)
    (ref.std(1) ** 2).sum().backward()
    assert T.allclose(mine.grad, ref.grad[:, rank * 2:(rank + 1) * 2] * world)     # x world: the optimizer divides by it
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, "ok"))


def test_gradsync_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, "ok"), (1, "ok")]


def test_abi_library_loads_and_exports_every_symbol():
    """No compute: the C-ABI library loads on a CPU-only box and exports every symbol of include/audiogan_b200.h."""
    import re
    from audiogan_b200 import _abi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "audiogan_b200.h")).read()
    declared = set(re.findall(r"\b(ag_[a-z0-9_]+)\s*\(", hdr))
    L = _abi.lib()
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert L.ag_version() >= 100
    assert set(_abi.exported_symbols()) <= declared | {"ag_last_error_string"}
