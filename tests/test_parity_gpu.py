"""GPU parity: the drop-in modules (CUDA kernels through the C ABI) against the CPU oracle
(oracle/restated.py, itself pinned to the reference classes by tests/test_oracle.py) and against the
committed golden vectors generated from the reference's own classes.  fp32 mode: <= 1e-5 relative
(BASELINE.json north_star); relative = max|a-b| / max|b| per tensor."""
import glob
import os
import warnings

import pytest
import torch as T

from oracle import restated as O
from audiogan_b200.synthetic import step_inputs

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")
RTOL = 1e-5
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-30)


class Report:
    def __init__(self):
        self.bad, self.rows = [], []

    def check(self, name, a, b, tol=RTOL):
        if tuple(a.shape) != tuple(b.shape):
            self.bad.append("%s: shape %s vs %s" % (name, tuple(a.shape), tuple(b.shape)))
            return
        r = rel(a, b)
        self.rows.append((name, r))
        if not (r <= tol):
            self.bad.append("%s: rel err %.3e > %.1e" % (name, r, tol))

    def done(self, name=None):
        if name:                      # keep the table as evidence (copied into profiles/ per round)
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "parity_%s.txt" % name), "w") as f:
                f.write("".join("%-52s %.3e\n" % r for r in self.rows))
        assert not self.bad, "\n".join(self.bad)


def noise_only(k):
    """bias `_v` tensors: w = g*sign(v), so d/dv is identically 0 and both sides hold rounding noise."""
    return k.split(".")[-1].startswith("bias") and k.endswith("_v")


def grad_tol(k):
    # the final conv's scalar bias gradient is a sum of B*L signed terms that cancel to ~1e-3 of their
    # absolute sum; fp32 summation order alone moves it by ~1e-4 relative
    return 2e-3 if k == "dense_res_gen.4.module.bias_g" else 5e-5


def build(case, dev="cuda"):
    import audiogan_b200 as ag
    gk, dk = case.get("gk", {}), case.get("dk", {})
    Pg = O.pin_stopper(O.init_generator(case.get("g_seed", 11), **gk))
    Pd = O.init_discriminator(case.get("d_seed", 12), **dk)
    g = ag.Generator(embed_size=100, **gk)
    d = ag.Discriminator(embed_size=100, **dk)
    g.load_state_dict(Pg)
    d.load_state_dict(Pd)
    if dev == "cpu":
        return Pg, Pd, g, d
    return Pg, Pd, g.to(dev), d.to(dev)


def to_dev(inp, dev="cuda"):
    return {k: (v.to(dev) if isinstance(v, T.Tensor) else v) for k, v in inp.items()}


CASES = {
    "small_h64_b4_l1000_mixed": dict(B=4, L=1000, full=False, gk={"state_size": 64}, dk={"state_size": 64}),
    "default_b3_l1200_mixed": dict(B=3, L=1200, full=False),
    "default_b2_l1600_full": dict(B=2, L=1600, full=True),
}


def bce_mean(cls, nf, tgt):
    w = O.length_mask(cls.shape, nf)
    return (O.bce_with_logits_per_sample(cls, T.full_like(cls, tgt), w) / nf.float()).mean()


@pytest.mark.parametrize("name", list(CASES))
def test_forward_and_grads_match_oracle(name):
    import audiogan_b200 as ag
    cs = CASES[name]
    Pg, Pd, g, d = build(cs)
    inp = step_inputs(cs["B"], cs["L"], seed=1234, full_length=cs["full"])
    di = to_dev(inp)
    R = Report()
    # ---------------- oracle (CPU fp32)
    Pg_r = {k: v.clone().requires_grad_(True) for k, v in Pg.items()}
    Pd_r = {k: v.clone().requires_grad_(True) for k, v in Pd.items()}
    z_r = inp["g_z"].clone().requires_grad_(True)
    x_r, s_r, stop_r, glen_r = O.generator_forward(Pg_r, inp["g_c_g"], z=z_r)
    ln = glen_r if cs["full"] else inp["real_len"]
    cls_r, hs_r, hl_r, nf_r = O.discriminator_forward(Pd_r, x_r + inp["g_noise_fake"], ln, inp["g_c_d"])
    loss_r = bce_mean(cls_r, nf_r, 0.5)
    gk, dk = list(Pg_r), list(Pd_r)
    grads_r = T.autograd.grad(loss_r, [Pg_r[k] for k in gk] + [Pd_r[k] for k in dk] + [z_r], allow_unused=True)
    # ---------------- CUDA path
    z = di["g_z"].clone().requires_grad_(True)
    x, s, stop_list, glen = g(z=z, c=di["g_c_g"], u_stop=None)
    R.check("G.s (stop logits)", s, s_r)
    R.check("G.x", x, x_r)
    assert T.equal(glen.cpu(), glen_r), (glen, glen_r)
    assert len(stop_list) == s_r.shape[1] and tuple(stop_list[0].shape) == (cs["B"], 1)
    ln_d = glen if cs["full"] else di["real_len"]
    cls, hs, hl, nf = d(x + di["g_noise_fake"], ln_d, di["g_c_d"])
    assert T.equal(nf.cpu(), nf_r)
    for i, (a, b) in enumerate(zip(hs, hs_r)):
        R.check("D.cnn[%d]" % i, a, b)
        assert T.equal(hl[i].cpu(), hl_r[i])
    R.check("D.logits", cls, cls_r)
    loss, _, _ = ag.masked_bce_mean(cls, nf, 0.5, -1.0)
    R.check("loss", loss.reshape(1), loss_r.reshape(1))
    loss.backward()
    sg, sd = dict(g.named_parameters()), dict(d.named_parameters())
    for k, gr in zip(gk, grads_r[:len(gk)]):
        got = sg[k].grad
        if noise_only(k):
            continue
        if gr is None:
            assert got is None or float(got.abs().max()) == 0.0, k
            continue
        R.check("dG/" + k, got, gr, tol=grad_tol(k))
    for k, gr in zip(dk, grads_r[len(gk):len(gk) + len(dk)]):
        if noise_only(k):
            continue
        R.check("dD/" + k, sd[k].grad, gr, tol=5e-5)
    R.check("dz", z.grad, grads_r[-1], tol=5e-5)
    R.done("oracle_" + name)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_against_reference_golden(path):
    """Outputs of the REFERENCE's own classes (tests/golden, made by oracle/make_golden.py)."""
    import audiogan_b200 as ag
    gold = T.load(path)
    cs = gold["case"]
    Pg, Pd, g, d = build(cs)
    di = to_dev(step_inputs(cs["B"], cs["L"], seed=cs["input_seed"], full_length=cs["full"]))
    R = Report()
    G = gold["G"]
    z = di["g_z"].clone().requires_grad_(True)
    x, s, _, glen = g(z=z, c=di["g_c_g"], u_stop=None)
    R.check("x", x, G["x"])
    R.check("s", s, G["s"])
    assert T.equal(glen.cpu(), G["len"])
    ln = glen if cs["full"] else di["real_len"]
    cls_g, hs, hl, nf = d(x + di["g_noise_fake"], ln, di["g_c_d"])
    R.check("cls_g", cls_g, G["cls_g"])
    for h, n in zip(hs, G["cnn_norms"]):
        # the golden norm is torch's CPU fp32 reduction over the contiguous (B, C, T) tensor (itself 3-6e-5 away from the fp64
        # norm at this size): reduce our activation in the same memory order
        assert abs(float(h.detach().float().cpu().contiguous().norm()) - n) <= 2e-5 * n
    loss, _, _ = ag.masked_bce_mean(cls_g, nf, 0.5, -1.0)
    assert abs(float(loss) - G["loss"]) < 2e-6
    loss.backward()
    # The L = 16000 fixtures (T_g = 80, T_d = 250) hold the reference's fp32 CPU gradients, which at that depth are themselves
    # 1e-4 ... 1e-3 away from an fp64 run of the same code (the control band measured in tests/test_config1_gpu.py, where the
    # CUDA gradients are judged against fp64): gradient summaries of those fixtures are held to 5e-4 (norm) / 1e-3 (elements).
    long_seq = cs["L"] > 2000

    def check_summary(tag, k, grad, gs):
        # the golden norm was taken with torch's CPU fp32 reduction: use the same reduction on our gradient
        tn, th = (max(grad_tol(k), 5e-4), max(grad_tol(k), 1e-3)) if long_seq else (grad_tol(k), grad_tol(k))
        R.check(tag + k + " |norm|", grad.cpu().norm().reshape(1), T.tensor([gs["norm"]]), tol=tn)
        R.rows.append((tag + k + " head", float((grad.flatten()[:32].cpu() - gs["head"]).abs().max()) / (gs["absmax"] + 1e-30)))
        if R.rows[-1][1] > th:
            R.bad.append("%s head err %.3e" % (tag + k, R.rows[-1][1]))

    for k, p in g.named_parameters():
        gs = G["grads"][k]
        if gs is None or noise_only(k):
            continue
        check_summary("dG/", k, p.grad, gs)
    # input gradients through 80 / 250 recurrent steps (the L = 16000 fixtures): the reference's own fp32 result is 1e-4 ... 1e-2
    # (element-wise) away from an fp64 run of the same code at this geometry (tests/test_config1_gpu.py's control band)
    tol_in = 5e-5 if cs["L"] <= 2000 else 3e-4
    assert abs(float(z.grad.cpu().norm()) - G["dz"]["norm"]) <= tol_in * G["dz"]["norm"]
    # D-update style losses on real + detached fake
    D = gold["D"]
    g.zero_grad(); d.zero_grad()
    real = (di["real"] + di["noise_real"]).requires_grad_(True)
    cls_d, _, _, nfd = d(real, di["real_len"], di["c_real"])
    with T.no_grad():
        xf, _, _, flen = g(z=di["z"], c=di["c_g"], u_stop=None)
    fk = (xf + di["noise_fake"]).detach().requires_grad_(True)
    cls_f, _, _, nff = d(fk, flen, di["c_d2"])
    R.check("cls_d", cls_d, D["cls_d"])
    R.check("cls_f", cls_f, D["cls_f"])
    loss_d, _, _ = ag.masked_bce_mean(cls_d, nfd, 0.9, 1.0)
    loss_f, _, _ = ag.masked_bce_mean(cls_f, nff, 0.0, -1.0)
    assert abs(float(loss_d) - D["loss_d"]) < 2e-6 and abs(float(loss_f) - D["loss_f"]) < 2e-6
    (loss_d + loss_f).backward()
    for k, p in d.named_parameters():
        if noise_only(k):
            continue
        check_summary("dD/", k, p.grad, D["grads"][k])
    assert abs(float(real.grad.cpu().norm()) - D["dreal"]["norm"]) <= tol_in * D["dreal"]["norm"]
    assert abs(float(fk.grad.cpu().norm()) - D["dfake"]["norm"]) <= tol_in * D["dfake"]["norm"]
    R.done("golden_" + os.path.basename(path)[:-3])


def test_core_step_matches_oracle_updates():
    """1 D-update + 1 G-update (SURVEY 8(d) core step): losses and post-step parameters."""
    import audiogan_b200 as ag
    cs = dict(B=3, L=1200, full=True, gk={"state_size": 64}, dk={"state_size": 64})
    Pg, Pd, g, d = build(cs)
    inp = step_inputs(cs["B"], cs["L"], seed=77, full_length=True)
    di = to_dev(inp)
    Pg_r = {k: v.clone() for k, v in Pg.items()}
    Pd_r = {k: v.clone() for k, v in Pd.items()}
    st_d, st_g = {}, {}
    gb = lambda dd: {"c_g": dd["g_c_g"], "c_d": dd["g_c_d"], "z": dd["g_z"], "noise_fake": dd["g_noise_fake"]}
    o1 = O.d_update(Pg_r, Pd_r, st_d, inp)
    o2 = O.g_update(Pg_r, Pd_r, st_g, gb(inp))
    opt_d = ag.FusedRMSprop(d.parameters(), lr=1e-4)
    opt_g = ag.FusedRMSprop(g.parameters(), lr=1e-4)
    di["u_stop"] = None
    m1 = ag.d_update(g, d, opt_d, di, clip=1.0, check=True)          # batched: real + fake in one 2B discriminator pass
    # the literal two-call sequence gives the same update
    Pg_b, Pd_b, g_b, d_b = build(cs)
    mb = ag.d_update(g_b, d_b, ag.FusedRMSprop(d_b.parameters(), lr=1e-4), di, clip=1.0, check=True, batched=False)
    assert abs(float(mb["loss"]) - float(m1["loss"])) < 1e-6
    for (k, p), (_, q) in zip(d.named_parameters(), d_b.named_parameters()):
        if not noise_only(k):
            assert rel(p, q) < 1e-5, k
    gbd = gb(di); gbd["u_stop"] = None
    m2 = ag.g_update(g, d, opt_g, gbd, clip=0.1, check=True)
    R = Report()
    R.check("loss_d", m1["loss_d"].reshape(1), T.tensor([o1["loss_d"]]))
    R.check("loss_g(D)", m1["loss_g"].reshape(1), T.tensor([o1["loss_g"]]))
    R.check("d_grad_norm", m1["d_grad_norm"].reshape(1), T.tensor([o1["d_grad_norm"]]), tol=5e-5)
    R.check("loss(G)", m2["loss"].reshape(1), T.tensor([o2["loss"]]))
    R.check("g_grad_norm", m2["g_grad_norm"].reshape(1), T.tensor([o2["g_grad_norm"]]), tol=5e-5)
    # RMSprop's first step is lr*g/(sqrt(0.01 g^2)+eps) ~ +-10*lr: compare the parameters themselves (the step
    # is 1e-3 absolute, so 1e-5 relative parity of p means the step agrees wherever the gradient is not noise)
    for k, p in d.named_parameters():
        if not noise_only(k):
            R.check("D after step " + k, p, Pd_r[k], tol=2e-5)
    for k, p in g.named_parameters():
        if not noise_only(k):
            R.check("G after step " + k, p, Pg_r[k], tol=2e-5)
    R.done("core_step")


@pytest.mark.parametrize("name", ["default_b3_l1200_mixed"])
def test_bf16_mode_within_2e2(name):
    """bf16 mode (tcgen05 GEMMs, fp32 accumulate / state): <= 2e-2 relative to the fp32 oracle (north_star)."""
    import audiogan_b200 as ag
    cs = CASES[name]
    Pg, Pd, g, d = build(cs)
    g.set_mode("bf16"); d.set_mode("bf16")
    inp = step_inputs(cs["B"], cs["L"], seed=1234, full_length=cs["full"])
    di = to_dev(inp)
    R = Report()
    Pg_r = {k: v.clone().requires_grad_(True) for k, v in Pg.items()}
    Pd_r = {k: v.clone().requires_grad_(True) for k, v in Pd.items()}
    z_r = inp["g_z"].clone().requires_grad_(True)
    x_r, s_r, _, glen_r = O.generator_forward(Pg_r, inp["g_c_g"], z=z_r)
    ln = glen_r if cs["full"] else inp["real_len"]
    cls_r, hs_r, hl_r, nf_r = O.discriminator_forward(Pd_r, x_r + inp["g_noise_fake"], ln, inp["g_c_d"])
    loss_r = bce_mean(cls_r, nf_r, 0.5)
    gk, dk = list(Pg_r), list(Pd_r)
    grads_r = T.autograd.grad(loss_r, [Pg_r[k] for k in gk] + [Pd_r[k] for k in dk] + [z_r], allow_unused=True)
    z = di["g_z"].clone().requires_grad_(True)
    x, s, _, glen = g(z=z, c=di["g_c_g"], u_stop=None)
    tol = 2e-2
    R.check("G.s", s, s_r, tol)
    R.check("G.x", x, x_r, tol)
    ln_d = glen if cs["full"] else di["real_len"]
    cls, hs, hl, nf = d(x + di["g_noise_fake"], ln_d, di["g_c_d"])
    for i, (a, b) in enumerate(zip(hs, hs_r)):
        R.check("D.cnn[%d]" % i, a, b, tol)
    R.check("D.logits", cls, cls_r, tol)
    loss, _, _ = ag.masked_bce_mean(cls, nf, 0.5, -1.0)
    R.check("loss", loss.reshape(1), loss_r.reshape(1), tol)
    loss.backward()
    # Gradients: bf16 operand rounding (~4e-3 per GEMM) flips the sign of the LeakyReLU pre-activations that lie
    # within that distance of zero (a fraction ~3e-3 of the units per layer); a flipped unit changes its derivative
    # from 1 to 0.01, so the gradient picks up a relative error ~sqrt(3e-3 * depth) ~ 0.1 -- inherent to ANY bf16
    # forward, and far above 2e-2.  The test therefore pins direction and norm: cosine >= 0.97, |norm ratio - 1| <= 0.1.
    def check_dir(tag, a, b):
        a, b = a.detach().float().cpu().flatten(), b.detach().float().cpu().flatten()
        cos = float(T.dot(a, b) / (a.norm() * b.norm() + 1e-30))
        ratio = float(a.norm() / (b.norm() + 1e-30))
        R.rows.append((tag + " (1-cos)", 1 - cos))
        R.rows.append((tag + " |norm ratio-1|", abs(ratio - 1)))
        if not (cos >= 0.97 and abs(ratio - 1) <= 0.1):
            R.bad.append("%s: cos %.4f norm ratio %.4f" % (tag, cos, ratio))

    sg, sd = dict(g.named_parameters()), dict(d.named_parameters())
    for k, gr in zip(gk, grads_r[:len(gk)]):
        if noise_only(k) or gr is None or gr.numel() < 8:
            continue
        check_dir("dG/" + k, sg[k].grad, gr)
    for k, gr in zip(dk, grads_r[len(gk):len(gk) + len(dk)]):
        if noise_only(k) or gr.numel() < 8:
            continue
        check_dir("dD/" + k, sd[k].grad, gr)
    check_dir("dz", z.grad, grads_r[-1])
    R.done("bf16_" + name)


def test_reference_faithful_extras_match_oracle():
    """The extra passes of the reference loop: FGSM moves + x_grad_norm in the D-update
    (audiogan.py:729-736, :752-759, :769-775), adversarially sampled z and the feature-matching penalty over
    calc_dists in the G-update (:836, :847-855).  Sign steps make the inputs of later passes discontinuous in the
    gradient, so the comparison is on losses / norms / penalties (1e-4) rather than on every parameter."""
    import audiogan_b200 as ag
    cs = dict(B=3, L=1200, full=True, gk={"state_size": 64}, dk={"state_size": 64})
    Pg, Pd, g, d = build(cs)
    inp = step_inputs(cs["B"], cs["L"], seed=99, full_length=True)
    di = to_dev(inp)
    di["u_stop"] = None
    Pg_r = {k: v.clone() for k, v in Pg.items()}
    Pd_r = {k: v.clone() for k, v in Pd.items()}
    o1 = O.d_update(Pg_r, Pd_r, {}, inp, with_x_grad_norm=True)
    opt_d = ag.FusedRMSprop(d.parameters(), lr=1e-4)
    m1 = ag.d_update(g, d, opt_d, di, clip=1.0, with_x_grad_norm=True, check=True)
    R = Report()
    R.check("loss_d", m1["loss_d"].reshape(1), T.tensor([o1["loss_d"]]), 1e-5)
    R.check("loss_g", m1["loss_g"].reshape(1), T.tensor([o1["loss_g"]]), 1e-5)
    R.check("x_grad_norm", m1["x_grad_norm"].reshape(1), T.tensor([o1["x_grad_norm"]]), 1e-4)
    # FGSM branch: runs the two extra D passes + sign steps; the real-branch loss is taken before the move
    Pg2, Pd2, g2, d2 = build(cs)
    o3 = O.d_update({k: v.clone() for k, v in Pg2.items()}, {k: v.clone() for k, v in Pd2.items()}, {}, inp, fgsm=True)
    m3 = ag.d_update(g2, d2, ag.FusedRMSprop(d2.parameters(), lr=1e-4), di, clip=1.0, fgsm=True, check=True)
    R.check("fgsm loss_d", m3["loss_d"].reshape(1), T.tensor([o3["loss_d"]]), 1e-5)
    R.check("fgsm loss_g", m3["loss_g"].reshape(1), T.tensor([o3["loss_g"]]), 2e-3)     # sign(grad) flips move x by 1e-3
    # G-update with feature matching (no sign steps: exact parity) and with adversarial z (sign steps: loose)
    gb_r = {"c_g": inp["g_c_g"], "c_d": inp["g_c_d"], "z": inp["g_z"], "noise_fake": inp["g_noise_fake"], "real": inp["real"],
            "real_len": inp["real_len"], "noise_real": inp["noise_real"], "noise_adv": inp["noise_fake"]}
    gb_d = to_dev(gb_r)
    gb_d["u_stop"] = None
    Pg3, Pd3, g3, d3 = build(cs)
    o2 = O.g_update({k: v.clone() for k, v in Pg3.items()}, {k: v.clone() for k, v in Pd3.items()}, {}, gb_r,
                    feature_matching=True)
    m2 = ag.g_update(g3, d3, ag.FusedRMSprop(g3.parameters(), lr=1e-4), gb_d, clip=0.1, feature_matching=True, check=True)
    R.check("G loss", m2["loss"].reshape(1), T.tensor([o2["loss"]]), 1e-5)
    R.check("feature_penalty", m2["feature_penalty"].detach().reshape(1), T.tensor([o2["feature_penalty"]]), 1e-4)
    R.check("g_grad_norm (fm)", m2["g_grad_norm"].reshape(1), T.tensor([o2["g_grad_norm"]]), 1e-4)
    # every raw generator gradient with the penalty's gradient injected into all six conv activations (calc_dists kernels)
    Pg5, Pd5, g5, d5 = build(cs)
    o5 = O.g_update({k: v.clone() for k, v in Pg5.items()}, {k: v.clone() for k, v in Pd5.items()}, {}, gb_r,
                    feature_matching=True, clip=0, lambda_fp=50.0)
    m5 = ag.g_update(g5, d5, ag.FusedRMSprop(g5.parameters(), lr=1e-4), gb_d, clip=0.0, feature_matching=True, lambda_fp=50.0)
    R.check("feature_penalty (lambda 50)", m5["feature_penalty"].detach().reshape(1), T.tensor([o5["feature_penalty"]]), 1e-4)
    for k, p in g5.named_parameters():
        if not noise_only(k) and o5["grads"].get(k) is not None:
            R.check("fm dG/" + k, p.grad, o5["grads"][k], tol=max(grad_tol(k), 1e-4))
    Pg4, Pd4, g4, d4 = build(cs)
    o4 = O.g_update({k: v.clone() for k, v in Pg4.items()}, {k: v.clone() for k, v in Pd4.items()}, {}, gb_r, adv_z=True)
    m4 = ag.g_update(g4, d4, ag.FusedRMSprop(g4.parameters(), lr=1e-4), gb_d, clip=0.1, adv_z=True, check=True)
    R.check("G loss (adv z)", m4["loss"].reshape(1), T.tensor([o4["loss"]]), 2e-3)
    # ---- the autograd.grad passes above must not leak weight gradients into the update (the reference never lets
    # autograd.grad touch p.grad): gradient norms and post-step parameters against the oracle's
    R.check("d_grad_norm (x_grad_norm)", m1["d_grad_norm"].reshape(1), T.tensor([o1["d_grad_norm"]]), 5e-5)
    for k, p in d.named_parameters():
        if not noise_only(k):
            R.check("D after step (x_grad_norm) " + k, p, Pd_r[k], tol=2e-5)
    R.check("d_grad_norm (fgsm)", m3["d_grad_norm"].reshape(1), T.tensor([o3["d_grad_norm"]]), 5e-3)
    R.check("g_grad_norm (adv z)", m4["g_grad_norm"].reshape(1), T.tensor([o4["g_grad_norm"]]), 5e-3)
    R.done("extras")


def test_autograd_grad_passes_leave_the_update_unchanged():
    """x_grad_norm (audiogan.py:769-775) is a logging-only data-gradient pass: with it on, the D-update must be the
    update without it, bit for bit up to atomics order (weight gradients leaked from autograd.grad passes)."""
    import audiogan_b200 as ag
    cs = dict(B=3, L=1200, full=True, gk={"state_size": 64}, dk={"state_size": 64})
    inp = step_inputs(cs["B"], cs["L"], seed=101, full_length=True)
    di = to_dev(inp)
    di["u_stop"] = None
    res = []
    for flag in (False, True):
        _, _, g, d = build(cs)
        m = ag.d_update(g, d, ag.FusedRMSprop(d.parameters(), lr=1e-4), di, clip=1.0, with_x_grad_norm=flag, batched=False)
        res.append((float(m["d_grad_norm"]), {k: p.detach().clone() for k, p in d.named_parameters()}))
    assert abs(res[0][0] - res[1][0]) <= 1e-5 * res[0][0], (res[0][0], res[1][0])
    for k in res[0][1]:
        if not noise_only(k):
            assert rel(res[1][1][k], res[0][1][k]) < 1e-6, k


@pytest.mark.parametrize("all_stop", [False, True])
def test_reinforce_stop_head_update_matches_oracle(all_stop):
    """REINFORCE for the stop head (audiogan.py:873-908) with live stop sampling on supplied uniforms: stop decisions,
    lengths, loss, EMA baseline over two calls, gradient norm and the post-step stop-head parameters."""
    import audiogan_b200 as ag
    cs = dict(B=4, L=1600, full=True, gk={"state_size": 64}, dk={"state_size": 64})
    Pg = O.init_generator(11, **cs["gk"])            # stop head NOT pinned: p(stop) ~ 0.5 per frame
    Pd = O.init_discriminator(12, **cs["dk"])
    g = ag.Generator(embed_size=100, **cs["gk"]); g.load_state_dict(Pg); g = g.cuda()
    d = ag.Discriminator(embed_size=100, **cs["dk"]); d.load_state_dict(Pd); d = d.cuda()
    inp = step_inputs(cs["B"], cs["L"], seed=55, full_length=True)
    gen = T.Generator().manual_seed(9)
    u = T.rand(cs["B"], 8, generator=gen) * 0.2 + 0.37          # some samples stop early, at different frames
    if all_stop:
        u[:, 4] = 0.0                                           # everyone has stopped after frame 5: early exit, T < Tcap
    else:
        u[0] = 1.0                                              # sample 0 never stops: runs all 8 frames
    gb_r = {"c_g": inp["g_c_g"], "c_d": inp["g_c_d"], "z": inp["g_z"], "noise_fake": inp["g_noise_fake"], "u_stop": u}
    gb_d = to_dev(gb_r)
    Pg_r = {k: v.clone() for k, v in Pg.items()}
    Pd_r = {k: v.clone() for k, v in Pd.items()}
    st_g = {}
    opt_g = ag.FusedRMSprop(g.parameters(), lr=1e-4)
    R = Report()
    base_r, base = None, None
    for it in range(2):
        o = O.g_update(Pg_r, Pd_r, st_g, gb_r, reinforce=True, baseline=base_r)
        m = ag.g_update(g, d, opt_g, gb_d, clip=0.1, reinforce=True, baseline=base, check=True)
        base_r, base = o["baseline"], m["baseline"]
        assert tuple(m["fake"].shape) == tuple(o["fake"].shape)
        lens_r = O.generator_forward({k: v for k, v in Pg.items()}, inp["g_c_g"], z=inp["g_z"], u_stop=u)[3] if it == 0 else None
        if lens_r is not None:
            assert T.equal(m["fake_len"].cpu(), lens_r), (m["fake_len"], lens_r)
            assert len(set(lens_r.tolist())) > 1, "the case must exercise ragged stop lengths"
        R.check("loss[%d]" % it, m["loss"].reshape(1), T.tensor([o["loss"]]))
        R.check("baseline[%d]" % it, m["baseline"].reshape(1), T.tensor([o["baseline"]]))
        # The REINFORCE advantage is reward - baseline = a difference of nearly equal numbers (|adv| ~ 1e-3 |reward| on this
        # case): the cancellation amplifies fp32 rounding of the per-sample losses ~1000x, so the stop head's gradient carries
        # ~2-4e-4 relative noise on ANY fp32 implementation (measured: the CPU oracle's own fp32 result is 2.4e-4 away from its
        # fp64 result on stopper.*, 1.7e-6 on proj.*).  stopper.* and the norm sum that contains it are held to 2e-3.
        R.check("g_grad_norm[%d]" % it, m["g_grad_norm"].reshape(1), T.tensor([o["g_grad_norm"]]), tol=2e-3)
        for k, p in g.named_parameters():
            if not noise_only(k):
                R.check("G after step %d %s" % (it, k), p, Pg_r[k],
                        tol=(1e-4 if k.startswith("stopper.") else 2e-5) if it == 0 else 2e-4)
        # every raw gradient of the same update (no clipping on either side; the oracle clips its copies in place)
        Pg_c = {k: v.detach().clone() for k, v in g.state_dict().items()}
        o_c = O.g_update({k: v.cpu() for k, v in Pg_c.items()}, Pd_r, {}, gb_r, clip=0, reinforce=True, baseline=base_r)
        g2 = ag.Generator(embed_size=100, **cs["gk"]); g2.load_state_dict(Pg_c); g2 = g2.cuda()
        ag.g_update(g2, d, ag.FusedRMSprop(g2.parameters(), lr=1e-4), gb_d, clip=0.0, reinforce=True, baseline=base)
        for k, p in g2.named_parameters():
            if not noise_only(k) and o_c["grads"].get(k) is not None:
                R.check("raw dG[%d]/%s" % (it, k), p.grad, o_c["grads"][k], tol=2e-3 if k.startswith("stopper.") else grad_tol(k))
    R.done("reinforce_all_stop" if all_stop else "reinforce")


def test_loss_trajectory_tracks_oracle():
    """Loss trajectories over repeated core steps (north_star: "loss trajectories must track").  Small nets, the same
    batch every step, CUDA path against the CPU oracle.  RMSprop divides by sqrt(E[g^2]): during the first steps every
    parameter moves by ~10*lr*sign(g), so a gradient element whose SIGN differs by rounding (|g| ~ 1e-7 of the
    tensor's scale) moves its parameter the other way -- any two fp32 implementations with different summation order
    (two BLAS libraries included) separate after a handful of steps.  Pinned here: the first 4 steps agree to 2e-5
    (fp32) / 2e-2 (bf16), the first 12 steps stay within 15 % of the oracle's curve, all 25 within 30 % (8 % on average)
    while the loss goes down."""
    import audiogan_b200 as ag
    cs = dict(B=2, L=800, full=True, gk={"state_size": 32}, dk={"state_size": 32})
    nsteps = 25
    inp = step_inputs(cs["B"], cs["L"], seed=5, full_length=True)
    gb = lambda dd: {"c_g": dd["g_c_g"], "c_d": dd["g_c_d"], "z": dd["g_z"], "noise_fake": dd["g_noise_fake"]}
    Pg, Pd, _, _ = build(cs, dev="cpu")
    Pg_r = {k: v.clone() for k, v in Pg.items()}
    Pd_r = {k: v.clone() for k, v in Pd.items()}
    st_d, st_g, ref = {}, {}, []
    for _ in range(nsteps):
        o1 = O.d_update(Pg_r, Pd_r, st_d, inp)
        o2 = O.g_update(Pg_r, Pd_r, st_g, gb(inp))
        ref.append((o1["loss_d"], o1["loss_g"], o2["loss"]))
    ref = T.tensor(ref)
    di = to_dev(inp)
    di["u_stop"] = None
    gbd = gb(di)
    gbd["u_stop"] = None
    for mode, tol in (("fp32", 2e-5), ("bf16", 2e-2)):
        _, _, g, d = build(cs)
        g.set_mode(mode); d.set_mode(mode)
        opt_d, opt_g = ag.FusedRMSprop(d.parameters(), lr=1e-4), ag.FusedRMSprop(g.parameters(), lr=1e-4)
        got = []
        for _ in range(nsteps):
            m1 = ag.d_update(g, d, opt_d, di, clip=1.0)
            m2 = ag.g_update(g, d, opt_g, gbd, clip=0.1)
            got.append(T.stack([m1["loss_d"], m1["loss_g"], m2["loss"]]))
        got = T.stack(got).cpu()
        dev_rel = (got - ref).abs() / ref.abs()
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "trajectory_%s.txt" % mode), "w") as f:
            f.write("# step: loss_d loss_g(D) loss(G)  [cuda %s] | [cpu oracle]\n" % mode)
            for a, b in zip(got.tolist(), ref.tolist()):
                f.write("%s | %s\n" % (" ".join("%.6f" % v for v in a), " ".join("%.6f" % v for v in b)))
        assert float(dev_rel[:4].max()) <= tol, (mode, "first steps", float(dev_rel[:4].max()))
        # past ~10 steps the two runs are different samples of a chaotic trajectory (see the docstring; the weight-gradient
        # atomics alone change the summation order from run to run): pin the first half tightly, the rest as a band
        assert float(dev_rel[:12].max()) <= 0.15, (mode, "first half", float(dev_rel[:12].max()))
        assert float(dev_rel.mean()) <= 0.08 and float(dev_rel.max()) <= 0.30, (mode, "whole curve", float(dev_rel.mean()), float(dev_rel.max()))
        assert float(got[-1, 0]) < float(got[0, 0])
    assert float(ref[-1, 0]) < float(ref[0, 0])            # the discriminator is actually learning on this batch


def test_scaled_discriminator_with_stride1_layers():
    """BASELINE configs[4] shape family ("2x conv channels and depth"): extra kernel-7 stride-1 layers between the
    stride-2 ones (SURVEY 8(d) cfg 5).  Forward activations, logits and every gradient against the oracle."""
    import audiogan_b200 as ag
    struct = [[7, 2, 8], [7, 1, 8], [7, 2, 16], [7, 1, 16], [7, 2, 32], [7, 1, 32]]
    cs = dict(B=3, L=1000, full=False, gk={"state_size": 32}, dk={"state_size": 32, "cnn_struct": struct})
    Pg, Pd, g, d = build(cs)
    inp = step_inputs(cs["B"], cs["L"], seed=21, full_length=False)
    di = to_dev(inp)
    R = Report()
    Pd_r = {k: v.clone().requires_grad_(True) for k, v in Pd.items()}
    x_r = (inp["real"] + inp["noise_real"]).requires_grad_(True)
    cls_r, hs_r, hl_r, nf_r = O.discriminator_forward(Pd_r, x_r, inp["real_len"], inp["c_real"], cnn_struct=struct)
    loss_r = bce_mean(cls_r, nf_r, 0.9)
    dk = list(Pd_r)
    grads_r = T.autograd.grad(loss_r, [Pd_r[k] for k in dk] + [x_r])
    x = (di["real"] + di["noise_real"]).requires_grad_(True)
    cls, hs, hl, nf = d(x, di["real_len"], di["c_real"])
    assert T.equal(nf.cpu(), nf_r) and len(hs) == len(struct)
    for i, (a, b) in enumerate(zip(hs, hs_r)):
        R.check("D.cnn[%d]" % i, a, b)
    R.check("D.logits", cls, cls_r)
    loss, _, _ = ag.masked_bce_mean(cls, nf, 0.9, 1.0)
    loss.backward()
    sd = dict(d.named_parameters())
    for k, gr in zip(dk, grads_r[:-1]):
        if not noise_only(k):
            R.check("dD/" + k, sd[k].grad, gr, tol=5e-5)
    R.check("dx", x.grad, grads_r[-1], tol=5e-5)
    R.done("scaled_d")


def test_double_hidden_generator_streams_weights():
    """BASELINE configs[3] shape family (2x recurrent hidden size): with H = 2048 a 64-row slice of [whh | wx] no longer
    fits in shared memory, so the recurrent kernels take the non-resident path (weights streamed from L2 each step).
    Generator outputs and gradients against the oracle."""
    cs = dict(B=2, L=600, full=True, gk={"state_size": 2048}, dk={"state_size": 64})
    Pg, Pd, g, d = build(cs)
    inp = step_inputs(cs["B"], cs["L"], seed=31, full_length=True)
    di = to_dev(inp)
    Pg_r = {k: v.clone().requires_grad_(True) for k, v in Pg.items()}
    z_r = inp["g_z"].clone().requires_grad_(True)
    x_r, s_r, _, _ = O.generator_forward(Pg_r, inp["g_c_g"], z=z_r)
    T.manual_seed(2)
    up = T.randn_like(x_r)
    gk = list(Pg_r)
    grads_r = T.autograd.grad((x_r * up).sum(), [Pg_r[k] for k in gk] + [z_r], allow_unused=True)
    R = Report()
    z = di["g_z"].clone().requires_grad_(True)
    x, s, _, _ = g(z=z, c=di["g_c_g"], u_stop=None)
    R.check("G.x", x, x_r)
    R.check("G.s", s, s_r)
    (x * up.cuda()).sum().backward()
    sg = dict(g.named_parameters())
    for k, gr in zip(gk, grads_r[:-1]):
        if gr is not None and not noise_only(k):
            R.check("dG/" + k, sg[k].grad, gr, tol=5e-5)
    R.check("dz", z.grad, grads_r[-1], tol=5e-5)
    R.done("h2048")


def test_cuda_graph_step_equals_eager_step():
    """graph.GraphedStep (the whole core step captured as one CUDA graph) against the eager launch sequence in fp32 mode:
    same seeds and batches -> same losses and the same parameters after three steps (atomics order only)."""
    import audiogan_b200 as ag
    cs = dict(B=3, L=1200, full=True, gk={"state_size": 64}, dk={"state_size": 64})
    batches = [to_dev(step_inputs(cs["B"], cs["L"], seed=300 + i, full_length=True)) for i in range(4)]
    for b in batches:
        b["real_len"] = b["real_len"].cpu()

    def eager_step(g, d, od, og, di):
        di = dict(di); di["u_stop"] = None
        m1 = ag.d_update(g, d, od, di, clip=1.0)
        gb = {"c_g": di["g_c_g"], "c_d": di["g_c_d"], "z": di["g_z"], "noise_fake": di["g_noise_fake"], "u_stop": None}
        m2 = ag.g_update(g, d, og, gb, clip=0.1)
        return T.stack([m1["loss_d"], m1["loss_g"], m2["loss"]]).cpu()

    _, _, g1, d1 = build(cs)
    od1, og1 = ag.FusedRMSprop(d1.parameters(), lr=1e-4), ag.FusedRMSprop(g1.parameters(), lr=1e-4)
    eager_step(g1, d1, od1, og1, batches[0])                      # GraphedStep's warm-up step
    le = [eager_step(g1, d1, od1, og1, batches[i]) for i in (1, 2, 3)]
    _, _, g2, d2 = build(cs)
    od2, og2 = ag.FusedRMSprop(d2.parameters(), lr=1e-4), ag.FusedRMSprop(g2.parameters(), lr=1e-4)
    gs = ag.GraphedStep(g2, d2, od2, og2, batches[0], warmup=1)
    assert gs.launches > 100
    lg = [gs.run(batches[i])["losses"].clone().cpu() for i in (1, 2, 3)]
    for a, b in zip(le, lg):
        # the two runs differ in the order of their atomic sums only; an RMSprop step is sign-like, so a parameter whose gradient is
        # rounding noise may step the other way (bounded below) and the losses of the later steps move by a few 1e-6 (observed <= 2.7e-6)
        assert float((a - b).abs().max()) < 1e-5, (a, b)
    for (k, p), (_, q) in zip(list(g1.named_parameters()) + list(d1.named_parameters()),
                              list(g2.named_parameters()) + list(d2.named_parameters())):
        if not noise_only(k):
            # three sign-like RMSprop steps: an element whose gradient is rounding noise may step the other way
            # (a capture bug -- a missing dependency, a reused buffer -- shows as O(1e-2) losses and most elements off)
            err = (p - q).abs()
            nbad = int((err > 5e-6).sum())
            assert nbad <= max(3, 2e-3 * err.numel()) and float(err.max()) < 6.1e-3, (k, nbad, err.numel(), float(err.max()))
    with pytest.raises(ValueError):
        bad = dict(batches[1]); bad["real_len"] = bad["real_len"] - 200
        gs.load(bad)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_core_step_with_batched_generator_pass_equals_the_two_updates(mode):
    """train.core_step runs the D-update's detached generator pass and the G-update's generator pass as ONE 2B forward
    (same parameters; only the second half is differentiated).  It must reproduce d_update followed by g_update: to fp32
    rounding in fp32 mode; in bf16 mode (other tile shapes -> other bf16 roundings) the batched variant must be as close to the
    fp32-mode result as the call-by-call variant is (control: both bf16 variants are measured against fp32 mode)."""
    import audiogan_b200 as ag
    cs = dict(B=5, L=2400, full=True, gk={"state_size": 128}, dk={"state_size": 128}) if mode == "fp32" else dict(B=6, L=3200, full=True)
    di = to_dev(step_inputs(cs["B"], cs["L"], seed=77, full_length=True))
    di["real_len"] = di["real_len"].cpu()
    di["u_stop"] = None

    def run(md, fused):
        _, _, g, d = build(cs)
        g.set_mode(md); d.set_mode(md)
        od, og = ag.FusedRMSprop(d.parameters(), lr=1e-4), ag.FusedRMSprop(g.parameters(), lr=1e-4)
        if fused:
            m1, m2 = ag.core_step(g, d, od, og, di, clip_d=0.0, clip_g=0.0)
        else:
            m1 = ag.d_update(g, d, od, di, clip=0.0)
            gb = {"c_g": di["g_c_g"], "c_d": di["g_c_d"], "z": di["g_z"], "noise_fake": di["g_noise_fake"], "u_stop": None}
            m2 = ag.g_update(g, d, og, gb, clip=0.0)
        return dict(ld=m1["loss_d"], lg=m1["loss_g"], l=m2["loss"], fake1=m1["fake"], fake2=m2["fake"],
                    gd={k: p.grad.clone() for k, p in d.named_parameters()}, gg={k: p.grad.clone() for k, p in g.named_parameters()})

    R = Report()
    if mode == "fp32":
        a, b = run("fp32", False), run("fp32", True)
        for k in ("ld", "lg", "l", "fake1", "fake2"):
            R.check(k, b[k].reshape(-1), a[k].reshape(-1), 2e-6)
        for tag, key in (("dD/", "gd"), ("dG/", "gg")):
            for k in a[key]:
                if not noise_only(k):
                    R.check(tag + k, b[key][k], a[key][k], 2e-5)
    else:
        ref, a, b = run("fp32", False), run("bf16", False), run("bf16", True)
        for k in ("ld", "lg", "l", "fake1", "fake2"):
            R.check(k, b[k].reshape(-1), ref[k].reshape(-1), 2e-2)
        l2 = lambda x, y: float((x.float() - y.float()).norm() / (y.float().norm() + 1e-30))
        for tag, key in (("dD/", "gd"), ("dG/", "gg")):
            for k in ref[key]:
                if noise_only(k) or ref[key][k].numel() < 8:
                    continue
                e_split, e_fused = l2(a[key][k], ref[key][k]), l2(b[key][k], ref[key][k])
                R.rows.append((tag + k + " relL2 vs fp32: call-by-call", e_split))
                R.rows.append((tag + k + " relL2 vs fp32: batched G pass", e_fused))
                if e_fused > max(2.0 * e_split, 0.05):
                    R.bad.append("%s%s: batched %.3f vs call-by-call %.3f (relative L2 error against fp32 mode)" % (tag, k, e_fused, e_split))
    R.done("core_step_%s" % mode)


def test_generator_batch_beyond_one_launch_runs_in_chunks_on_the_fast_kernels():
    """B = 136 > the 128 samples (forward) / 64 samples (BPTT) one launch of the TMEM-resident generator recurrence takes:
    the engine runs consecutive launches over row slices of the batch (samples are independent) instead of dropping to the
    grid-barrier kernels (per-GPU batch 128 fell off the fast path silently).  Checked against the CPU oracle."""
    import audiogan_b200 as ag
    Bn, L = 136, 800
    Pg = O.pin_stopper(O.init_generator(11))
    g = ag.Generator(embed_size=100); g.load_state_dict(Pg); g = g.cuda().set_mode("bf16")
    gen = T.Generator().manual_seed(3)
    z = T.randn(Bn, L // 200, 100, generator=gen)
    c = T.randn(Bn, 100, generator=gen)
    up = T.randn(Bn, L, generator=gen)
    zr = z.clone().requires_grad_(True)
    Pr = {k: v.clone().requires_grad_(True) for k, v in Pg.items()}
    xr, sr, _, _ = O.generator_forward(Pr, c, z=zr)
    (xr * up).sum().backward()
    zd = z.cuda().requires_grad_(True)
    x, s_, _, _ = g(z=zd, c=c.cuda(), u_stop=None)
    (x * up.cuda()).sum().backward()
    assert g._plan.last_path.get("g_fwd") == "tmem" and g._plan.last_path.get("g_bwd") == "tmem", g._plan.last_path
    R = Report()
    R.check("x (136 samples)", x, xr, 2e-2)
    R.check("s", s_, sr, 2e-2)
    for lo, hi in ((0, 64), (64, 128), (128, 136)):                 # every BPTT chunk
        a, b = zd.grad[lo:hi].flatten().cpu(), zr.grad[lo:hi].flatten()
        cos = float(T.dot(a, b) / (a.norm() * b.norm()))
        R.rows.append(("dz cos rows %d:%d" % (lo, hi), cos))
        if cos < 0.97:
            R.bad.append("dz rows %d:%d cos %.4f" % (lo, hi, cos))
    for k in ("rnn.0.module.weight_hh_v", "proj.module.weight_v", "dense_res_gen.1.module.conv.weight_v"):
        a, b = dict(g.named_parameters())[k].grad.flatten().cpu(), Pr[k].grad.flatten()
        cos = float(T.dot(a, b) / (a.norm() * b.norm()))
        R.rows.append(("d%s cos" % k, cos))
        if cos < 0.97:
            R.bad.append("%s cos %.4f" % (k, cos))
    R.done("gen_batch_chunks")


def test_double_hidden_generator_runs_stepwise_on_tensor_cores_in_bf16_mode():
    """configs[3]'s generator (state 2048): its recurrent weights (36.8 MB in bf16) cannot stay resident on chip, so in bf16 mode
    the recurrence runs frame by frame on tcgen05 GEMMs over the bf16 weights (engine._gen_stepwise_*, csrc/lstm_step.cu)
    instead of the fp32 grid-barrier kernel that streams fp32 weights.  Against the CPU oracle: forward <= 2e-2, gradients by
    direction and norm."""
    import audiogan_b200 as ag
    gk = {"state_size": 2048}
    Bn, L = 5, 1000
    Pg = O.pin_stopper(O.init_generator(11, **gk))
    g = ag.Generator(embed_size=100, **gk); g.load_state_dict(Pg); g = g.cuda().set_mode("bf16")
    gen = T.Generator().manual_seed(3)
    z, c, up = T.randn(Bn, L // 200, 100, generator=gen), T.randn(Bn, 100, generator=gen), T.randn(Bn, L, generator=gen)
    zr = z.clone().requires_grad_(True)
    Pr = {k: v.clone().requires_grad_(True) for k, v in Pg.items()}
    xr, sr, _, _ = O.generator_forward(Pr, c, z=zr)
    ((xr * up).sum() + 3.0 * sr.sum()).backward()                   # the stop logits take part too (they do under REINFORCE)
    zd = z.cuda().requires_grad_(True)
    x, s_, _, ln = g(z=zd, c=c.cuda(), u_stop=None)
    ((x * up.cuda()).sum() + 3.0 * s_.sum()).backward()
    assert g._plan.last_path["g_fwd"].startswith("stepwise") and g._plan.last_path["g_bwd"].startswith("stepwise"), g._plan.last_path
    assert ln.tolist() == [L] * Bn
    R = Report()
    R.check("x", x, xr, 2e-2)
    R.check("s", s_, sr, 2e-2)
    for tag, a, b in [("dz", zd.grad, zr.grad)] + [("d" + k, dict(g.named_parameters())[k].grad, Pr[k].grad) for k in
                                                  ("rnn.0.module.weight_hh_v", "rnn.0.module.weight_ih_v", "proj.module.weight_v",
                                                   "rnn.0.module.bias_hh_g", "dense_res_gen.0.module.conv.weight_v")]:
        if a is None or b is None:
            continue
        a, b = a.flatten().float().cpu(), b.flatten()
        if float(b.norm()) == 0:
            continue
        cos, ratio = float(T.dot(a, b) / (a.norm() * b.norm())), float(a.norm() / b.norm())
        R.rows.append((tag + " cos", cos))
        if cos < 0.97 or abs(ratio - 1) > 0.1:
            R.bad.append("%s: cos %.4f norm ratio %.4f" % (tag, cos, ratio))
    R.done("gen_stepwise_h2048")
