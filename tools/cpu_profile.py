import sys, os, cProfile, pstats
sys.path.insert(0, "/root/repo")
os.chdir("/root/repo")
import torch
import audiogan_b200 as ag
from audiogan_b200.synthetic import step_inputs
dev = torch.device("cuda")
g = ag.pin_stopper(ag.Generator(embed_size=100)).to(dev); d = ag.Discriminator(embed_size=100).to(dev)
g.set_mode("bf16"); d.set_mode("bf16")
opt_d, opt_g = ag.FusedRMSprop(d.parameters(), lr=1e-4), ag.FusedRMSprop(g.parameters(), lr=1e-4)
inp = step_inputs(2, 1600, seed=1, full_length=True)
di = {k: (v if k.endswith("_len") else v.to(dev)) for k, v in inp.items()}; di["u_stop"] = None
gb = {"c_g": di["g_c_g"], "c_d": di["g_c_d"], "z": di["g_z"], "noise_fake": di["g_noise_fake"], "u_stop": None}
def step():
    ag.d_update(g, d, opt_d, di, clip=1.0)
    ag.g_update(g, d, opt_g, gb, clip=0.1)
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20): step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
