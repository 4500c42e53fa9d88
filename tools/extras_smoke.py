import sys, os
sys.path.insert(0, "/root/repo"); os.chdir("/root/repo")
import torch
import audiogan_b200 as ag
from audiogan_b200.synthetic import step_inputs
dev = torch.device("cuda")
for mode in ("bf16", "fp32"):
    torch.manual_seed(3)
    g = ag.pin_stopper(ag.Generator(embed_size=100, state_size=128)).to(dev); d = ag.Discriminator(embed_size=100, state_size=128).to(dev)
    g.set_mode(mode); d.set_mode(mode)
    opt_d, opt_g = ag.FusedRMSprop(d.parameters(), lr=1e-4), ag.FusedRMSprop(g.parameters(), lr=1e-4)
    inp = step_inputs(4, 3200, seed=1, full_length=False)
    di = {k: (v if k.endswith("_len") else v.to(dev)) for k, v in inp.items()}; di["u_stop"] = None
    m1 = ag.d_update(g, d, opt_d, di, clip=1.0, fgsm=True, with_x_grad_norm=True, check=True)
    gb = {"c_g": di["g_c_g"], "c_d": di["g_c_d"], "z": di["g_z"], "noise_fake": di["g_noise_fake"], "u_stop": None,
          "real": di["real"], "noise_real": di["noise_real"], "real_len": di["real_len"], "noise_adv": di["noise_fake"]}
    m2 = ag.g_update(g, d, opt_g, gb, clip=0.1, feature_matching=True, adv_z=True, check=True)
    torch.cuda.synchronize()
    print(mode, "d_update(fgsm, x_grad_norm): loss_d %.5f loss_g %.5f x_grad_norm %.3e | g_update(feature matching, adv z): loss %.5f fp %.5f" % (
        float(m1["loss_d"]), float(m1["loss_g"]), float(m1["x_grad_norm"]), float(m2["loss"]), float(m2["feature_penalty"])))
