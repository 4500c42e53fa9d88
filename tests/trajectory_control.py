"""1000-step loss trajectories with an ORACLE-vs-ORACLE control band (north_star: "loss trajectories must track over 1k steps";
VERDICT r1: the divergence of the round-1 curves was called chaos without a control run).

    python tests/trajectory_control.py oracle   # CPU, build container: writes tests/golden/aux/trajectory_control.pt
    python tests/trajectory_control.py cuda     # GPU box: CUDA fp32 / bf16 curves against the committed control band

Control = the SAME CPU oracle (oracle/restated.py) run four ways that differ only at rounding level: fp32, fp64, and fp32
from initial parameters perturbed by a relative 1e-7 (two seeds).  Whatever separates those curves is the game's sensitivity to
rounding, not an implementation difference; the CUDA path is judged against the envelope of the four.  B = 16 samples per
minibatch (round 1 used 2), small nets (state 32), L = 800, eight minibatches in rotation, lr 1e-4, clip 1 / 0.1."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch as T
from oracle import restated as O
from audiogan_b200.synthetic import step_inputs

NSTEPS, W = 1000, 50
CS = dict(B=16, L=800, gk={"state_size": 32}, dk={"state_size": 32})
GOLD = os.path.join(ROOT, "tests", "golden", "aux", "trajectory_control.pt")
batches = [step_inputs(CS["B"], CS["L"], seed=100 + i, full_length=True) for i in range(8)]
gb = lambda dd: {"c_g": dd["g_c_g"], "c_d": dd["g_c_d"], "z": dd["g_z"], "noise_fake": dd["g_noise_fake"]}


def init():
    return O.pin_stopper(O.init_generator(11, **CS["gk"])), O.init_discriminator(12, **CS["dk"])


def oracle_curve(dtype, perturb_seed=None):
    Pg, Pd = init()
    if perturb_seed is not None:
        gen = T.Generator().manual_seed(perturb_seed)
        for P in (Pg, Pd):
            for k in P:
                P[k] = P[k] * (1 + 1e-7 * T.randn(P[k].shape, generator=gen))
    cv = (lambda v: v.to(dtype) if T.is_tensor(v) and v.is_floating_point() else v)
    Pg, Pd = {k: cv(v) for k, v in Pg.items()}, {k: cv(v) for k, v in Pd.items()}
    bs = [{k: cv(v) for k, v in b.items()} for b in batches]
    T.set_default_dtype(dtype)
    st_d, st_g, out = {}, {}, []
    try:
        for i in range(NSTEPS):
            o1 = O.d_update(Pg, Pd, st_d, bs[i % 8])
            o2 = O.g_update(Pg, Pd, st_g, gb(bs[i % 8]))
            out.append((o1["loss_d"], o1["loss_g"], o2["loss"]))
    finally:
        T.set_default_dtype(T.float32)
    return T.tensor(out, dtype=T.float32)


def windows(c):
    return c.unfold(0, W, W).mean(-1)                                  # [NSTEPS / W, 3]


if sys.argv[1] == "oracle":
    T.set_num_threads(os.cpu_count() or 8)
    curves = {}
    for name, args in (("oracle fp32", (T.float32,)), ("oracle fp64", (T.float64,)), ("oracle fp32 init*(1+1e-7 n) seed 1", (T.float32, 1)),
                       ("oracle fp32 init*(1+1e-7 n) seed 2", (T.float32, 2))):
        t0 = time.time()
        curves[name] = oracle_curve(*args)
        print(name, "%.0f s" % (time.time() - t0), windows(curves[name])[-1].tolist(), flush=True)
    T.save({"case": CS, "nsteps": NSTEPS, "curves": curves, "torch": str(T.__version__)}, GOLD)
else:
    import audiogan_b200 as ag
    from test_parity_gpu import to_dev
    gold = T.load(GOLD)
    ctrl = {k: windows(v) for k, v in gold["curves"].items()}
    stack = T.stack(list(ctrl.values()))                               # [4, windows, 3]
    lo, hi, mid = stack.min(0).values, stack.max(0).values, stack.mean(0)
    res = {}
    for mode in ("fp32", "bf16"):
        Pg, Pd = init()
        g = ag.Generator(embed_size=100, **CS["gk"]); g.load_state_dict(Pg)
        d = ag.Discriminator(embed_size=100, **CS["dk"]); d.load_state_dict(Pd)
        g, d = g.cuda().set_mode(mode), d.cuda().set_mode(mode)
        od, og = ag.FusedRMSprop(d.parameters(), lr=1e-4), ag.FusedRMSprop(g.parameters(), lr=1e-4)
        dbs = []
        for b in batches:
            di = to_dev(b); di["u_stop"] = None; di["real_len"] = di["real_len"].cpu()
            dbs.append(di)
        got = []
        for i in range(NSTEPS):
            m1, m2 = ag.core_step(g, d, od, og, dbs[i % 8], clip_d=1.0, clip_g=0.1)
            got.append(T.stack([m1["loss_d"], m1["loss_g"], m2["loss"]]))
        res["cuda " + mode] = windows(T.stack(got).cpu())
    lines = ["# %d core steps, B=%d, L=%d, state 32 nets, 8 minibatches in rotation; %d-step window means of loss_d(real) / loss_g(D on fake) / loss(G)"
             % (NSTEPS, CS["B"], CS["L"], W),
             "# control band = envelope of 4 CPU-oracle runs that differ at rounding level only (fp32, fp64, fp32 with initial parameters * (1 + 1e-7 n) x 2)",
             "# step | control lo .. hi per loss | cuda fp32 | cuda bf16"]
    for i in range(lo.shape[0]):
        lines.append("%5d | %s | %s | %s" % ((i + 1) * W, "  ".join("%.4f..%.4f" % (a, b) for a, b in zip(lo[i].tolist(), hi[i].tolist())),
                                             " ".join("%.4f" % x for x in res["cuda fp32"][i].tolist()),
                                             " ".join("%.4f" % x for x in res["cuda bf16"][i].tolist())))
    spread = (hi - lo)
    for name, c in list(ctrl.items()) + list(res.items()):
        # distance from the control's centre in units of the control's own spread (window by window), and relative to the value
        out_by = T.clamp(T.maximum(lo - c, c - hi), min=0)
        lines.append("# %-38s mean |x - centre| / centre %.3f ; windows inside the band %.0f %% ; mean excursion outside the band / centre %.3f ; last window %s"
                     % (name, float(((c - mid).abs() / mid.abs()).mean()), 100 * float((out_by == 0).float().mean()),
                        float((out_by / mid.abs()).mean()), ["%.4f" % x for x in c[-1].tolist()]))
    lines.append("# control spread itself: mean (hi - lo) / centre %.3f, max %.3f" % (float((spread / mid.abs()).mean()), float((spread / mid.abs()).max())))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "trajectory_control.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))
