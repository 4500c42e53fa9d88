"""Data-parallel parity on real GPUs (needs >= 2; skipped on a single-GPU box): N ranks x per-rank batch B over NCCL must
compute what ONE rank computes on the concatenated N*B minibatch (batch sharding + summed gradient all-reduce
+ 1/world in the optimizer == the single-process mean), including the feature-matching batch statistics over the global
batch (dist.gather_batch), the reduced accuracy counters, the all-reduce overlapped with backward (GradSync.attach) and the
step captured as a CUDA graph with its NCCL all-reduces inside."""
import os
import socket

import pytest
import torch as T
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
B, L = 3, 1600
KW = dict(gk={"state_size": 128}, dk={"state_size": 128})


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _models(mode, dev):
    import audiogan_b200 as ag
    from oracle import restated as O
    Pg = O.pin_stopper(O.init_generator(11, **KW["gk"]))
    Pd = O.init_discriminator(12, **KW["dk"])
    g = ag.Generator(embed_size=100, **KW["gk"]); g.load_state_dict(Pg)
    d = ag.Discriminator(embed_size=100, **KW["dk"]); d.load_state_dict(Pd)
    g, d = g.to(dev).set_mode(mode), d.to(dev).set_mode(mode)
    return g, d, ag.FusedRMSprop(d.parameters(), lr=1e-4), ag.FusedRMSprop(g.parameters(), lr=1e-4)


def _shard(inp, rank, world, dev):
    out = {}
    for k, v in inp.items():
        n = v.shape[0] // world
        v = v[rank * n:(rank + 1) * n]
        out[k] = v if k.endswith("_len") else v.to(dev)
    out["u_stop"] = None
    return out


def _flat(mods):
    return T.cat([p.detach().reshape(-1) for m in mods for p in m.parameters()])


def _close(a, b, what, frac_max=2e-3):
    # fp32 mode, one sign-like RMSprop step: elements whose gradient is rounding noise may step the other way
    err = (a - b).abs()
    assert float((err > 2e-6).float().mean()) < frac_max and float(err.max()) < 2.1e-3, (what, float(err.max()),
                                                                                         float((err > 2e-6).float().mean()))


def _worker(rank, world, port, q):
    try:
        _run(rank, world, port)
        q.put((rank, "ok"))
    except BaseException:                                        # noqa: BLE001 -- reported to the parent, which stops the peers
        import traceback
        q.put((rank, "FAILED\n" + traceback.format_exc()))


def _run(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import audiogan_b200 as ag
    from audiogan_b200 import dist as agd
    from audiogan_b200.synthetic import step_inputs
    _, _, local = agd.init()
    dev = T.device("cuda", local)
    T.cuda.set_device(dev)
    glob = step_inputs(world * B, L, seed=900, full_length=True)
    whole, mine = _shard(glob, 0, 1, dev), _shard(glob, rank, world, dev)
    gb = lambda di: {"c_g": di["g_c_g"], "c_d": di["g_c_d"], "z": di["g_z"], "noise_fake": di["g_noise_fake"], "u_stop": None,
                     "real": di["real"], "real_len": di["real_len"], "noise_real": di["noise_real"]}
    # ---- (1) core step: one rank on the global batch vs `world` ranks on its shards (post-backward all-reduce)
    g0, d0, od0, og0 = _models("fp32", dev)
    r1, r2 = ag.core_step(g0, d0, od0, og0, whole, clip_d=1.0, clip_g=0.1)
    ref = _flat([g0, d0])
    for early in (False, True):
        g1, d1, od1, og1 = _models("fp32", dev)
        agd.broadcast_parameters([g1, d1])
        sync = agd.GradSync(nbuckets=4)
        if early:
            sync.attach(g1, d1)                                     # all-reduce packed regions while backward is running
        m1, m2 = ag.core_step(g1, d1, od1, og1, mine, clip_d=1.0, clip_g=0.1, grad_sync=sync)
        _close(_flat([g1, d1]), ref, "core step, early=%s" % early)
        assert abs(float(m1["d_grad_norm"]) - float(r1["d_grad_norm"])) <= 2e-4 * float(r1["d_grad_norm"]), (early, "d norm")
        assert abs(float(m2["g_grad_norm"]) - float(r2["g_grad_norm"])) <= 2e-4 * float(r2["g_grad_norm"]), (early, "g norm")
        st_d, st_g = agd.reduce_stats(m1["stats_d"], m1["stats_g"])                 # accuracy gates (audiogan.py:813)
        assert T.allclose(st_d, r1["stats_d"]) and T.allclose(st_g, r1["stats_g"])
        lm = T.stack([m1["loss_d"], m1["loss_g"], m2["loss"]])
        T.distributed.all_reduce(lm)
        assert T.allclose(lm / world, T.stack([r1["loss_d"], r1["loss_g"], r2["loss"]]), rtol=2e-6, atol=1e-7)
    assert agd.any_rank(rank == world - 1) and not agd.any_rank(False)
    # ---- (2) G-update with the feature-matching penalty: batch statistics over the GLOBAL minibatch
    g0, d0, od0, og0 = _models("fp32", dev)
    rr = ag.g_update(g0, d0, og0, gb(whole), clip=0.0, feature_matching=True, lambda_fp=20.0)
    g1, d1, od1, og1 = _models("fp32", dev)
    mm = ag.g_update(g1, d1, og1, gb(mine), clip=0.0, feature_matching=True, lambda_fp=20.0, grad_sync=agd.GradSync(),
                     gather=agd.gather_batch)
    assert abs(float(mm["feature_penalty"]) - float(rr["feature_penalty"])) <= 1e-4 * abs(float(rr["feature_penalty"]))
    for (k, p), (_, p0) in zip(g1.named_parameters(), g0.named_parameters()):
        if k.split(".")[-1].startswith("bias") and k.endswith("_v"):
            continue
        a, b = p.grad / world, p0.grad                               # summed over ranks; the optimizer applies 1/world
        assert float((a - b).abs().max()) <= 1e-4 * float(b.abs().max()) + 1e-12, ("fm grad", k)
    # ---- (3) the step as a CUDA graph with the NCCL all-reduces captured inside, bf16 mode (the benched configuration)
    g2, d2, od2, og2 = _models("bf16", dev)
    agd.broadcast_parameters([g2, d2])
    gs = ag.GraphedStep(g2, d2, od2, og2, mine, grad_sync=agd.GradSync(nbuckets=4), warmup=1)
    for _ in range(2):
        out = gs.run(mine)
    T.cuda.synchronize()
    assert bool(T.isfinite(out["losses"]).all())
    flat = _flat([g2, d2])
    other = flat.clone()
    T.distributed.broadcast(other, 0)
    assert T.equal(flat, other), "ranks diverged under the captured all-reduce"
    # ---- (4) the library's all-reduce over NVLink peer memory (csrc/peer.cu) instead of NCCL: same parameters as the single
    # rank on the global batch, and the whole data-parallel step as ONE graph (the collectives are plain kernels)
    g3, d3, od3, og3 = _models("fp32", dev)
    agd.broadcast_parameters([g3, d3])
    peer = agd.PeerGradSync([g3, d3])
    m1, m2 = ag.core_step(g3, d3, od3, og3, mine, clip_d=1.0, clip_g=0.1, grad_sync=peer)
    _close(_flat([g3, d3]), ref, "core step, peer-memory all-reduce")
    assert abs(float(m1["d_grad_norm"]) - float(r1["d_grad_norm"])) <= 2e-4 * float(r1["d_grad_norm"])
    g4, d4, od4, og4 = _models("bf16", dev)
    agd.broadcast_parameters([g4, d4])
    gs2 = ag.GraphedStep(g4, d4, od4, og4, mine, grad_sync=agd.PeerGradSync([g4, d4]), warmup=1)
    assert len(gs2.graphs) == 1
    for _ in range(3):
        out = gs2.run(mine)
    T.cuda.synchronize()
    assert bool(T.isfinite(out["losses"]).all())
    flat = _flat([g4, d4])
    other = flat.clone()
    T.distributed.broadcast(other, 0)
    assert T.equal(flat, other), "ranks diverged under the peer-memory all-reduce"
    T.distributed.barrier()
    T.distributed.destroy_process_group()


@pytest.mark.skipif(T.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_two_ranks_equal_one_rank_on_the_global_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = []
    try:
        for _ in range(world):                                   # a rank that fails leaves its peer inside a collective:
            got.append(q.get(timeout=300))                       # stop at the first report of a failure
            if got[-1][1] != "ok":
                break
    finally:
        for p in procs:
            p.join(5 if all(g[1] == "ok" for g in got) and len(got) == world else 0.1)
            if p.is_alive():
                p.terminate()
    assert sorted(got) == [(r, "ok") for r in range(world)], "\n".join("rank %d: %s" % g for g in got)
