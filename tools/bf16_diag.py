"""fp32-mode vs bf16-mode of the same modules on the GPU: where does the bf16 error enter?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as T
import audiogan_b200 as ag
from audiogan_b200.synthetic import step_inputs

def rels(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30)), float((a - b).norm() / (b.norm() + 1e-30))

T.manual_seed(0)
B, L = 4, 3200
g = ag.pin_stopper(ag.Generator(embed_size=100)).cuda()
d = ag.Discriminator(embed_size=100).cuda()
inp = {k: v.cuda() for k, v in step_inputs(B, L, seed=3).items()}
out = {}
for mode in ("fp32", "bf16"):
    g.set_mode(mode); d.set_mode(mode)
    g.zero_grad(); d.zero_grad()
    # D alone on a real waveform
    x = (inp["real"] + inp["noise_real"]).clone().requires_grad_(True)
    cls, hs, hl, nf = d(x, inp["real_len"], inp["c_real"])
    loss, _, _ = ag.masked_bce_mean(cls, nf, 0.9, 1.0)
    gh = T.autograd.grad(loss, [x] + list(hs), retain_graph=True, allow_unused=True)
    loss.backward()
    r = {"D.logits": cls.detach(), "D.dx": gh[0]}
    for i, h in enumerate(hs):
        r["D.act%d" % i] = h.detach()
    for k, p in d.named_parameters():
        if not (k.split(".")[-1].startswith("bias") and k.endswith("_v")):
            r["dD/" + k] = p.grad.clone()
    # G alone with a fixed upstream gradient
    d.zero_grad()
    z = inp["g_z"].clone().requires_grad_(True)
    xg, s, _, _ = g(z=z, c=inp["g_c_g"], u_stop=None)
    T.manual_seed(1)
    up = T.randn_like(xg) * 1e-3
    (xg * up).sum().backward()
    r["G.x"] = xg.detach(); r["G.dz"] = z.grad.clone()
    for k, p in g.named_parameters():
        if not (k.split(".")[-1].startswith("bias") and k.endswith("_v")) and p.grad is not None:
            r["dG/" + k] = p.grad.clone()
    out[mode] = r
for k in out["fp32"]:
    a, b = rels(out["bf16"][k], out["fp32"][k])
    print("%-50s max-rel %.3e  fro-rel %.3e" % (k, a, b))
