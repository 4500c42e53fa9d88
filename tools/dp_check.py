"""Data-parallel consistency check (run under torchrun, 2+ GPUs): one core step with the packed gradient regions all-reduced
DURING backward (GradSync.attach) must leave the same parameters as the same step with the gradients reduced after backward,
and both must equal on every rank.  Prints the largest relative parameter difference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiogan_b200 as ag
from audiogan_b200 import dist as agd
from audiogan_b200.synthetic import step_inputs

rank, world, local = agd.init()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
res = {}
for mode in ("fp32", "bf16"):
    for early in (False, True):
        torch.manual_seed(7)
        g = ag.pin_stopper(ag.Generator(embed_size=100, state_size=128)).to(dev)
        d = ag.Discriminator(embed_size=100, state_size=128).to(dev)
        g.set_mode(mode); d.set_mode(mode)
        agd.broadcast_parameters([g, d])
        opt_d, opt_g = ag.FusedRMSprop(d.parameters(), lr=1e-3), ag.FusedRMSprop(g.parameters(), lr=1e-3)
        sync = agd.GradSync(nbuckets=4)
        if early:
            sync.attach(g, d)
        inp = step_inputs(4, 3200, seed=100 + rank, full_length=True)
        di = {k: (v if k.endswith("_len") else v.to(dev)) for k, v in inp.items()}
        di["u_stop"] = None
        norms = []
        for _ in range(2):
            m1 = ag.d_update(g, d, opt_d, di, clip=1.0, grad_sync=sync)
            gb = {"c_g": di["g_c_g"], "c_d": di["g_c_d"], "z": di["g_z"], "noise_fake": di["g_noise_fake"], "u_stop": None}
            m2 = ag.g_update(g, d, opt_g, gb, clip=0.1, grad_sync=sync)
            norms.append((float(m1["d_grad_norm"]), float(m2["g_grad_norm"])))
        torch.cuda.synchronize()
        res[(mode, early, "n")] = norms
        res[(mode, early)] = torch.cat([p.detach().reshape(-1) for p in list(g.parameters()) + list(d.parameters())])
    a, b = res[(mode, False)], res[(mode, True)]
    rel = float((a - b).abs().max() / a.abs().max())
    other = a.clone()
    torch.distributed.broadcast(other, 0)
    across = float((a - other).abs().max())
    moved = float((a - torch.cat([p.detach().reshape(-1) for p in []] or [a * 0]).to(dev)).abs().max())
    if rank == 0:
        n0, n1 = res[(mode, False, "n")], res[(mode, True, "n")]
        print("%s: sum of per-tensor gradient norms (reduced gradients) first step: late D %.6f G %.6f | early D %.6f G %.6f" % (
            mode, n0[0][0], n0[0][1], n1[0][0], n1[0][1]), flush=True)
        print("%s: early-vs-late reduction max rel param diff %.3e ; rank0-vs-rank%d max abs diff %.3e" % (mode, rel, rank, across), flush=True)
    else:
        print("%s: rank %d vs rank 0 max abs param diff %.3e (early-vs-late %.3e)" % (mode, rank, across, rel), flush=True)
torch.distributed.destroy_process_group()
