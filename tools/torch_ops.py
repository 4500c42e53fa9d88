"""Which torch (ATen) ops still launch kernels inside one eager core step, with the Python line that issues them
(torch.profiler, one step after warm-up).  python tools/torch_ops.py [B] [L] > profiles/r2_torch_ops.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import audiogan_b200 as ag
from audiogan_b200.synthetic import step_inputs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
L = int(sys.argv[2]) if len(sys.argv) > 2 else 16000
dev = torch.device("cuda")
torch.manual_seed(1)
g = ag.pin_stopper(ag.Generator(embed_size=100)).to(dev).set_mode("bf16")
d = ag.Discriminator(embed_size=100).to(dev).set_mode("bf16")
od, og = ag.FusedRMSprop(d.parameters(), lr=1e-4), ag.FusedRMSprop(g.parameters(), lr=1e-4)
inp = step_inputs(B, L, seed=3)
di = {k: (v if k.endswith("_len") else v.to(dev)) for k, v in inp.items()}
di["u_stop"] = None


def step():
    ag.d_update(g, d, od, di, clip=1.0)
    gb = {"c_g": di["g_c_g"], "c_d": di["g_c_d"], "z": di["g_z"], "noise_fake": di["g_noise_fake"], "u_stop": None}
    ag.g_update(g, d, og, gb, clip=0.1)


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith("aten::")]
rows = {}
for e in ev:
    kt = sum(k.duration for k in e.kernels) if hasattr(e, "kernels") else 0
    if not getattr(e, "kernels", None):
        continue
    st = [s for s in (e.stack or []) if "audiogan_b200" in s or "bench.py" in s or "tools/" in s]
    key = (e.name, st[0].split("/root/repo/")[-1] if st else "?")
    r = rows.setdefault(key, [0, 0.0])
    r[0] += len(e.kernels)
    r[1] += kt
tot_n = sum(r[0] for r in rows.values())
tot_t = sum(r[1] for r in rows.values())
print("# torch ops that launch kernels in one eager core step (B=%d L=%d): %d kernels, %.1f us of device time" % (B, L, tot_n, tot_t))
for (name, where), (n, t) in sorted(rows.items(), key=lambda kv: -kv[1][0]):
    print("%4d kernels %8.1f us  %-28s %s" % (n, t, name, where))
