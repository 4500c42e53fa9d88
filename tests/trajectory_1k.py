"""1000-step loss trajectories (north_star: "loss trajectories must track over 1k steps"): the CUDA path (fp32 and bf16
modes) against the CPU oracle, small nets, a fresh synthetic minibatch every step (8 distinct batches in rotation), same
initial parameters.  Single steps are chaotic (see tests/test_parity_gpu.py::test_loss_trajectory_tracks_oracle), so the
curves are compared as 50-step moving averages.  Writes gpurun_out/trajectory_1k.txt (copied to profiles/)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch as T
from oracle import restated as O
from audiogan_b200.synthetic import step_inputs
import audiogan_b200 as ag
from test_parity_gpu import build, to_dev

NSTEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
cs = dict(B=2, L=800, full=True, gk={"state_size": 32}, dk={"state_size": 32})
batches = [step_inputs(cs["B"], cs["L"], seed=100 + i, full_length=True) for i in range(8)]
gb = lambda dd: {"c_g": dd["g_c_g"], "c_d": dd["g_c_d"], "z": dd["g_z"], "noise_fake": dd["g_noise_fake"]}
T.set_num_threads(os.cpu_count() or 8)
Pg, Pd, _, _ = build(cs, dev="cpu")
Pg_r = {k: v.clone() for k, v in Pg.items()}
Pd_r = {k: v.clone() for k, v in Pd.items()}
st_d, st_g, ref = {}, {}, []
t0 = time.time()
for i in range(NSTEPS):
    inp = batches[i % 8]
    o1 = O.d_update(Pg_r, Pd_r, st_d, inp)
    o2 = O.g_update(Pg_r, Pd_r, st_g, gb(inp))
    ref.append((o1["loss_d"], o1["loss_g"], o2["loss"]))
ref = T.tensor(ref)
print("oracle: %d steps in %.1f s" % (NSTEPS, time.time() - t0), flush=True)
curves = {"cpu oracle": ref}
for mode in ("fp32", "bf16"):
    _, _, g, d = build(cs)
    g.set_mode(mode); d.set_mode(mode)
    opt_d, opt_g = ag.FusedRMSprop(d.parameters(), lr=1e-4), ag.FusedRMSprop(g.parameters(), lr=1e-4)
    dbs = []
    for b in batches:
        di = to_dev(b); di["u_stop"] = None
        gbd = gb(di); gbd["u_stop"] = None
        dbs.append((di, gbd))
    got = []
    for i in range(NSTEPS):
        di, gbd = dbs[i % 8]
        m1 = ag.d_update(g, d, opt_d, di, clip=1.0)
        m2 = ag.g_update(g, d, opt_g, gbd, clip=0.1)
        got.append(T.stack([m1["loss_d"], m1["loss_g"], m2["loss"]]))
    curves["cuda " + mode] = T.stack(got).cpu()
W = 50
ma = {k: v.unfold(0, W, W).mean(-1) for k, v in curves.items()}       # [NSTEPS/W, 3]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "trajectory_1k.txt"), "w") as f:
    f.write("# %d core steps (1 D-update + 1 G-update), small nets (state 32), B=2, L=800, 8 batches in rotation, lr 1e-4\n" % NSTEPS)
    f.write("# %d-step moving averages of loss_d(real) / loss_g(D on fake) / loss(G):  cpu oracle | cuda fp32 | cuda bf16\n" % W)
    for i in range(ma["cpu oracle"].shape[0]):
        f.write("%5d  " % ((i + 1) * W) + " | ".join(" ".join("%.4f" % x for x in ma[k][i].tolist()) for k in ("cpu oracle", "cuda fp32", "cuda bf16")) + "\n")
    for k in ("cuda fp32", "cuda bf16"):
        dev = ((ma[k] - ma["cpu oracle"]).abs() / ma["cpu oracle"].abs())
        f.write("# %s vs oracle, moving averages: max rel deviation %.3f, mean %.3f; last window %s vs %s\n" % (
            k, float(dev.max()), float(dev.mean()), ["%.4f" % x for x in ma[k][-1].tolist()], ["%.4f" % x for x in ma["cpu oracle"][-1].tolist()]))
print(open(os.path.join(ROOT, "gpurun_out", "trajectory_1k.txt")).read())
