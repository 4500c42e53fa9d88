"""The "vendor-kernel bar" of SURVEY 2.2 / BASELINE.md 4: the reference's step (the CPU oracle's restatement, stock torch ops --
cuBLASLt / cuDNN / ATen sm_100 kernels, no audiogan_b200 code) executed ON THE B200 at configs[1], timed beside this repo's path.
Writes gpurun_out/stock_torch_gpu.txt; asserts only that the hand-written path is not slower than stock torch bf16 autocast."""
import os
import time

import pytest
import torch as T

from oracle import restated as O
from audiogan_b200.synthetic import step_inputs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _time_oracle(Pg, Pd, inp, steps, autocast):
    gb = {"c_g": inp["g_c_g"], "c_d": inp["g_c_d"], "z": inp["g_z"], "noise_fake": inp["g_noise_fake"]}
    sd, sg = {}, {}
    def one():
        with T.autocast("cuda", dtype=T.bfloat16, enabled=autocast):
            O.d_update(Pg, Pd, sd, inp)
            O.g_update(Pg, Pd, sg, gb)
    one()
    T.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    T.cuda.synchronize()
    return (time.perf_counter() - t0) / steps


def test_stock_torch_step_on_the_same_gpu():
    import audiogan_b200 as ag
    B, L = 64, 16000
    dev = T.device("cuda")
    inp = step_inputs(B, L, seed=1234, full_length=True)
    di = {k: v.to(dev) for k, v in inp.items()}
    rows = []
    try:
        for name, tf32, ac in (("fp32 (allow_tf32=False)", False, False), ("tf32", True, False), ("bf16 autocast", True, True)):
            T.backends.cuda.matmul.allow_tf32 = tf32
            T.backends.cudnn.allow_tf32 = tf32
            Pg = {k: v.to(dev) for k, v in O.pin_stopper(O.init_generator(11)).items()}
            Pd = {k: v.to(dev) for k, v in O.init_discriminator(12).items()}
            rows.append((name, _time_oracle(Pg, Pd, di, 2, ac)))
            del Pg, Pd
            T.cuda.empty_cache()
    except Exception as e:                      # the oracle is CPU test infrastructure: not every op path must run on CUDA
        pytest.skip("oracle did not run on CUDA: %r" % (e,))
    finally:
        T.backends.cuda.matmul.allow_tf32 = False
        T.backends.cudnn.allow_tf32 = True
    # this repo's path, bf16 mode, same shapes
    g = ag.pin_stopper(ag.Generator(embed_size=100)).to(dev)
    d = ag.Discriminator(embed_size=100).to(dev)
    g.set_mode("bf16"); d.set_mode("bf16")
    opt_d, opt_g = ag.FusedRMSprop(d.parameters(), lr=1e-4), ag.FusedRMSprop(g.parameters(), lr=1e-4)
    dj = {k: (v if k.endswith("_len") else v.to(dev)) for k, v in inp.items()}
    dj["u_stop"] = None
    gb = {"c_g": dj["g_c_g"], "c_d": dj["g_c_d"], "z": dj["g_z"], "noise_fake": dj["g_noise_fake"], "u_stop": None}
    def ours():
        ag.d_update(g, d, opt_d, dj, clip=1.0)
        ag.g_update(g, d, opt_g, gb, clip=0.1)
    for _ in range(3):
        ours()
    T.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        ours()
    T.cuda.synchronize()
    t_ours = (time.perf_counter() - t0) / 10
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "stock_torch_gpu.txt"), "w") as f:
        f.write("# core GAN step (1 D-update + 1 G-update), default nets, B=64, L=16000, one B200, wall clock around synchronised steps\n")
        for name, t in rows:
            f.write("stock torch %-24s %9.2f ms/step  %8.1f audio-s/s\n" % (name, t * 1e3, B * L / 8000 / t))
        f.write("audiogan_b200 bf16 mode            %9.2f ms/step  %8.1f audio-s/s  (%.1fx stock torch bf16 autocast, %.1fx stock fp32)\n" % (
            t_ours * 1e3, B * L / 8000 / t_ours, rows[2][1] / t_ours, rows[0][1] / t_ours))
    assert t_ours < rows[2][1]
