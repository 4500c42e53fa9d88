"""GPU check of graph.GraphedStep: the captured step must produce what the eager step produces (same seeds, same inputs),
and its timing beside the eager loop.  python tools/graph_check.py [B] [L] [mode]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audiogan_b200 as ag
from audiogan_b200.synthetic import step_inputs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
L = int(sys.argv[2]) if len(sys.argv) > 2 else 3200
mode = sys.argv[3] if len(sys.argv) > 3 else "bf16"
dev = torch.device("cuda")


def make():
    torch.manual_seed(7)
    g = ag.pin_stopper(ag.Generator(embed_size=100)).to(dev).set_mode(mode)
    d = ag.Discriminator(embed_size=100).to(dev).set_mode(mode)
    return g, d, ag.FusedRMSprop(d.parameters(), lr=1e-4), ag.FusedRMSprop(g.parameters(), lr=1e-4)


batches = [step_inputs(B, L, seed=100 + i) for i in range(4)]
on_dev = lambda h: {k: (v if k.endswith("_len") else v.to(dev)) for k, v in h.items()}
dbat = [on_dev(b) for b in batches]


def eager_step(g, d, od, og, di):
    di = dict(di); di["u_stop"] = None
    m1 = ag.d_update(g, d, od, di, clip=1.0)
    gb = {"c_g": di["g_c_g"], "c_d": di["g_c_d"], "z": di["g_z"], "noise_fake": di["g_noise_fake"], "u_stop": None}
    m2 = ag.g_update(g, d, og, gb, clip=0.1)
    return torch.stack([m1["loss_d"], m1["loss_g"], m2["loss"]])


WU = 2
# eager: WU warm-up steps on batch 0 (what GraphedStep does), then batches 1..3
g1, d1, od1, og1 = make()
for _ in range(WU):
    eager_step(g1, d1, od1, og1, dbat[0])
le = [eager_step(g1, d1, od1, og1, dbat[i]).cpu() for i in (1, 2, 3)]
g2, d2, od2, og2 = make()
gs = ag.GraphedStep(g2, d2, od2, og2, dbat[0], warmup=WU)
lg = [gs.run(dbat[i])["losses"].clone().cpu() for i in (1, 2, 3)]
print("graph launches per step", gs.launches)
for a, b in zip(le, lg):
    print("eager", a.tolist(), "graph", b.tolist())
worst = 0.0
for (k, p), (_, q) in zip(list(g1.named_parameters()) + list(d1.named_parameters()), list(g2.named_parameters()) + list(d2.named_parameters())):
    e = float((p - q).abs().max()) / (float(p.abs().max()) + 1e-30)
    worst = max(worst, e)
print("max relative parameter difference after 3 steps: %.3e" % worst)
tol = 1e-6 if mode == "fp32" else 5e-3
assert all(float((a - b).abs().max()) <= tol for a, b in zip(le, lg)), "graph replay diverges from the eager step"


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_host / n * 1e3


i = [0]
def f_e():
    i[0] += 1; eager_step(g1, d1, od1, og1, dbat[i[0] % 4])
def f_g():
    i[0] += 1; gs.run(dbat[i[0] % 4])
print("eager  %.3f ms/step (host enqueue %.3f ms)" % timeit(f_e))
print("graph  %.3f ms/step (host enqueue %.3f ms)" % timeit(f_g))
print("graph check ok")
