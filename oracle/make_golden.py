"""Generate tests/golden/*.pt from the REFERENCE'S OWN classes (audiogan.py:1-552 exec'd by
oracle/ref_loader.py) on stock torch fp32 CPU.  Run in the build container only
(`python -m oracle.make_golden`); /root/reference does not exist on the GPU box, so the
fixtures are committed.  TEST INFRASTRUCTURE ONLY.

Parameters are *not* stored (59 MB): they are regenerated from
oracle.restated.init_generator / init_discriminator seeds, which are deterministic
(torch CPU Generator).  Each fixture stores inputs' seeds, outputs, losses, and for every
parameter gradient its L2 norm plus the first 32 entries.
"""
import os
import sys
import warnings

import torch as T

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader as R            # noqa: E402
from oracle import restated as O              # noqa: E402
from audiogan_b200.synthetic import step_inputs  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (B, L, full_length, g_kwargs, d_kwargs)
    "default_b3_l1200_mixed": (3, 1200, False, {}, {}),
    "default_b2_l1600_full": (2, 1600, True, {}, {}),
    "small_h64_b4_l1000_mixed": (4, 1000, False, {"state_size": 64}, {"state_size": 64}),
    # BASELINE configs[1] geometry (2 s waveforms: T_g = 80 generator frames, T_d = 250 discriminator frames), small batch
    "default_b4_l16000_full": (4, 16000, True, {}, {}),
    "default_b4_l16000_mixed": (4, 16000, False, {}, {}),
}


def grad_summary(g):
    if g is None:
        return None
    return {"norm": float(g.norm()), "head": g.flatten()[:32].clone(), "absmax": float(g.abs().max())}


def run_case(name, B, L, full, gk, dk):
    ns = R.load()
    g = ns["Generator"](embed_size=100, **gk)
    d = ns["Discriminator"](embed_size=100, **dk)
    Pg = O.pin_stopper(O.init_generator(11, **gk))
    Pd = O.init_discriminator(12, **dk)
    g.load_state_dict(Pg)
    d.load_state_dict(Pd)
    inp = step_inputs(B, L, seed=1234, full_length=full)
    bce = ns["binary_cross_entropy_with_logits_per_sample"]
    lm = ns["length_mask"]
    out = {"case": dict(B=B, L=L, full=full, gk=gk, dk=dk, g_seed=11, d_seed=12, input_seed=1234)}
    with R.py2_tensor_semantics():
        # --- generator forward (audiogan.py:412-468), G-update-style loss through D (:841-864, :897)
        z = inp["g_z"].clone().requires_grad_(True)
        x, s, stop_list, glen = g(z=z, c=inp["g_c_g"])
        fake = x + inp["g_noise_fake"]
        ln = glen if full else inp["real_len"]       # mixed lengths exercise D's masks on fake data too
        cls_g, hs, hl, nf = d(fake, ln, inp["g_c_d"])
        loss_g = (bce(cls_g, T.full_like(cls_g, 0.5), lm(cls_g.size(), nf)) / nf.float()).mean()
        g.zero_grad(); d.zero_grad()
        loss_g.backward()
        out["G"] = {"x": x.detach().clone(), "s": s.detach().clone(), "len": glen.clone(),
                    "cls_g": cls_g.detach().clone(), "nframes": nf.clone(), "loss": float(loss_g),
                    "dz": grad_summary(z.grad),
                    "grads": {k: grad_summary(p.grad) for k, p in g.named_parameters()},
                    "cnn_norms": [float(h.norm()) for h in hs],
                    "cnn_heads": [h.detach().flatten()[:32].clone() for h in hs]}
        # --- discriminator loss on real + (detached) fake, D-update style (:723-728, :761-785)
        real = (inp["real"] + inp["noise_real"]).requires_grad_(True)
        g.zero_grad(); d.zero_grad()
        cls_d, _, _, nfd = d(real, inp["real_len"], inp["c_real"])
        loss_d = (bce(cls_d, T.full_like(cls_d, 0.9), lm(cls_d.size(), nfd)) / nfd.float()).mean()
        with T.no_grad():
            xf, _, _, flen = g(z=inp["z"], c=inp["c_g"])
        fk = (xf + inp["noise_fake"]).detach().requires_grad_(True)
        cls_f, _, _, nff = d(fk, flen, inp["c_d2"])
        loss_f = (bce(cls_f, T.zeros_like(cls_f), lm(cls_f.size(), nff)) / nff.float()).mean()
        (loss_d + loss_f).backward()
        out["D"] = {"cls_d": cls_d.detach().clone(), "cls_f": cls_f.detach().clone(),
                    "loss_d": float(loss_d), "loss_f": float(loss_f),
                    "dreal": grad_summary(real.grad), "dfake": grad_summary(fk.grad),
                    "grads": {k: grad_summary(p.grad) for k, p in d.named_parameters()}}
    T.save(out, os.path.join(OUT, name + ".pt"))
    print(name, "loss_g", out["G"]["loss"], "loss_d", out["D"]["loss_d"], "loss_f", out["D"]["loss_f"])


def run_embedder():
    """Embedder (audiogan.py:302-334): the reference class's own parameters (53.6 k floats), inputs and output."""
    ns = R.load()
    T.manual_seed(7)
    ref = ns["Embedder"](output_size=100)
    chars = T.randint(0, 256, (6, 12))
    lens = T.tensor([12, 3, 7, 1, 5, 12])
    with R.py2_tensor_semantics():
        c = ref(chars, lens)
        up = T.randn(6, 100)
        (c * up).sum().backward()
    os.makedirs(os.path.join(OUT, "aux"), exist_ok=True)
    T.save({"state_dict": {k: v.clone() for k, v in ref.state_dict().items()}, "chars": chars, "lens": lens, "c": c.detach().clone(),
            "up": up, "grads": {k: p.grad.clone() for k, p in ref.named_parameters()}},
           os.path.join(OUT, "aux", "embedder.pt"))
    print("embedder", float(c.norm()))


if __name__ == "__main__":
    warnings.filterwarnings("ignore")
    os.makedirs(OUT, exist_ok=True)
    T.manual_seed(0)
    only = sys.argv[1:]
    for name, args in CASES.items():
        if not only or name in only:
            run_case(name, *args)
    if not only or "embedder" in only:
        run_embedder()
