"""CPU: pin oracle/restated.py against (a) the committed golden vectors generated from the
reference's own classes and (b) those classes executed live when /root/reference exists."""
import glob
import os
import warnings

import pytest
import torch as T

from oracle import ref_loader as R
from oracle import restated as O
from audiogan_b200.synthetic import step_inputs

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))
warnings.filterwarnings("ignore")


def _close(a, b, rtol=2e-5, what=""):
    scale = float(b.abs().max()) + 1e-30
    err = float((a - b).abs().max())
    assert err <= rtol * scale, "%s: err %.3e scale %.3e" % (what, err, scale)


def _check_grads(got, gold, bias_scale):
    for k, gs in gold.items():
        g = got.get(k)
        if gs is None:
            assert g is None or float(g.abs().max()) == 0.0, k
            continue
        if k.endswith("bias_v"):     # d/dv of g*sign(v) == 0: both sides hold rounding noise only
            assert float(g.abs().max()) <= 1e-4 * bias_scale[k[:-1] + "g"] + 1e-12, k
            continue
        assert abs(float(g.norm()) - gs["norm"]) <= 2e-5 * gs["norm"] + 1e-12, k
        assert float((g.flatten()[:32] - gs["head"]).abs().max()) <= 2e-5 * gs["absmax"] + 1e-12, k


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_restated_matches_reference_golden(path):
    gold = T.load(path)
    cs = gold["case"]
    Pg = {k: v.requires_grad_(True) for k, v in O.pin_stopper(O.init_generator(cs["g_seed"], **cs["gk"])).items()}
    Pd = {k: v.requires_grad_(True) for k, v in O.init_discriminator(cs["d_seed"], **cs["dk"]).items()}
    inp = step_inputs(cs["B"], cs["L"], seed=cs["input_seed"], full_length=cs["full"])
    z = inp["g_z"].clone().requires_grad_(True)
    x, s, stop, glen = O.generator_forward(Pg, inp["g_c_g"], z=z)
    G = gold["G"]
    _close(x, G["x"], what="x")
    _close(s, G["s"], what="s")
    assert T.equal(glen, G["len"])
    ln = glen if cs["full"] else inp["real_len"]
    cls_g, hs, hl, nf = O.discriminator_forward(Pd, x + inp["g_noise_fake"], ln, inp["g_c_d"])
    _close(cls_g, G["cls_g"], what="cls_g")
    assert T.equal(nf, G["nframes"])
    for h, n, hd in zip(hs, G["cnn_norms"], G["cnn_heads"]):
        assert abs(float(h.norm()) - n) <= 2e-5 * n
        _close(h.flatten()[:32], hd, what="cnn head")
    loss = (O.bce_with_logits_per_sample(cls_g, T.full_like(cls_g, 0.5), O.length_mask(cls_g.shape, nf)) / nf.float()).mean()
    assert abs(float(loss) - G["loss"]) < 1e-6
    keys = list(Pg)
    grads = T.autograd.grad(loss, [Pg[k] for k in keys] + [z], allow_unused=True)
    got = dict(zip(keys, grads[:-1]))
    _check_grads(got, G["grads"], {k: v["absmax"] for k, v in G["grads"].items() if v is not None})
    assert abs(float(grads[-1].norm()) - G["dz"]["norm"]) <= 2e-5 * G["dz"]["norm"]
    # D-update style
    D = gold["D"]
    real = (inp["real"] + inp["noise_real"]).requires_grad_(True)
    cls_d, _, _, nfd = O.discriminator_forward(Pd, real, inp["real_len"], inp["c_real"])
    loss_d = (O.bce_with_logits_per_sample(cls_d, T.full_like(cls_d, 0.9), O.length_mask(cls_d.shape, nfd)) / nfd.float()).mean()
    with T.no_grad():
        xf, _, _, flen = O.generator_forward(Pg, inp["c_g"], z=inp["z"])
    fk = (xf + inp["noise_fake"]).detach().requires_grad_(True)
    cls_f, _, _, nff = O.discriminator_forward(Pd, fk, flen, inp["c_d2"])
    loss_f = (O.bce_with_logits_per_sample(cls_f, T.zeros_like(cls_f), O.length_mask(cls_f.shape, nff)) / nff.float()).mean()
    _close(cls_d, D["cls_d"], what="cls_d")
    _close(cls_f, D["cls_f"], what="cls_f")
    assert abs(float(loss_d) - D["loss_d"]) < 1e-6 and abs(float(loss_f) - D["loss_f"]) < 1e-6
    keys = list(Pd)
    grads = T.autograd.grad(loss_d + loss_f, [Pd[k] for k in keys] + [real, fk])
    _check_grads(dict(zip(keys, grads[:-2])), D["grads"], {k: v["absmax"] for k, v in D["grads"].items()})
    assert abs(float(grads[-2].norm()) - D["dreal"]["norm"]) <= 2e-5 * D["dreal"]["norm"]
    assert abs(float(grads[-1].norm()) - D["dfake"]["norm"]) <= 2e-5 * D["dfake"]["norm"]


@pytest.mark.skipif(not R.available(), reason="/root/reference not present (GPU box)")
def test_restated_helpers_match_live_reference():
    ns = R.load()
    T.manual_seed(3)
    x = T.randn(4, 9)
    t = T.rand(4, 9)
    w = O.length_mask((4, 9), T.tensor([9, 5, 1, 7]))
    with R.py2_tensor_semantics():
        ref = ns["binary_cross_entropy_with_logits_per_sample"](x, t, w)
        wr = ns["length_mask"]((4, 9), T.tensor([9, 5, 1, 7]))
    assert T.equal(w, wr)
    assert T.allclose(O.bce_with_logits_per_sample(x, t, w), ref, atol=1e-7)
    with pytest.raises(ValueError):
        O.bce_with_logits_per_sample(x, t[:, :3])
    hs = [T.randn(4, 5, 11).abs() * O.length_mask((4, 11), T.tensor([11, 6, 3, 9])).unsqueeze(1)]
    ls = [T.tensor([11, 6, 3, 9])]
    with R.py2_tensor_semantics():
        dr = ns["calc_dists"](hs, ls)
    do = O.calc_dists(hs, ls)
    assert len(dr) == len(do) == 9
    for a, b in zip(dr, do):
        assert T.allclose(a[0], b[0], atol=1e-6) and T.allclose(a[1], b[1], atol=1e-6)
    # clip_grad: per-tensor clip, returns the sum of norms (audiogan.py:243-253)
    ps = [T.nn.Parameter(T.randn(7, 3)), T.nn.Parameter(T.randn(5))]
    for p in ps:
        p.grad = T.randn_like(p) * 3
    mine = [p.grad.clone() for p in ps]
    with R.py2_tensor_semantics():
        nr = ns["clip_grad"](ps, 0.5)
    no = O.clip_grad(mine, 0.5)
    assert abs(float(nr) - no) < 1e-5
    for p, m in zip(ps, mine):
        assert T.allclose(p.grad, m, atol=1e-6)


def test_rmsprop_matches_torch():
    T.manual_seed(0)
    p = {"a": T.randn(33), "b": T.randn(4, 5)}
    q = [T.nn.Parameter(v.clone()) for v in p.values()]
    opt = T.optim.RMSprop(q, lr=1e-4)
    st = {}
    for _ in range(3):
        gs = {k: T.randn_like(v) for k, v in p.items()}
        for qq, g in zip(q, gs.values()):
            qq.grad = g.clone()
        opt.step()
        O.rmsprop_step(p, gs, st, lr=1e-4)
    for qq, v in zip(q, p.values()):
        assert T.allclose(qq.data, v, atol=1e-7)


def test_core_step_runs_and_losses_finite():
    Pg = O.pin_stopper(O.init_generator(1, state_size=32))
    Pd = O.init_discriminator(2, state_size=32)
    inp = step_inputs(2, 800, seed=5)
    sd, sg = {}, {}
    o1 = O.d_update(Pg, Pd, sd, inp, with_x_grad_norm=True)
    gb = {"c_g": inp["g_c_g"], "c_d": inp["g_c_d"], "z": inp["g_z"], "noise_fake": inp["g_noise_fake"],
          "real": inp["real"], "real_len": inp["real_len"], "noise_real": inp["noise_real"],
          "noise_adv": inp["noise_fake"]}
    o2 = O.g_update(Pg, Pd, sg, gb, feature_matching=True, adv_z=True, reinforce=True)
    o3 = O.d_update(Pg, Pd, sd, inp, fgsm=True)
    for o in (o1, o2, o3):
        assert all(v == v for v in [o["loss"]])
    assert o1["x_grad_norm"] >= 0


@pytest.mark.skipif(not R.available(), reason="/root/reference not present (GPU box)")
def test_embedder_and_pickling_match_reference_api():
    """The drop-in Embedder has the reference's state_dict keys (its numbers are checked on the GPU against the golden
    fixture written from the reference class: tests/test_recurrent_vs_torch_gpu.py) and, like every module of the package,
    refuses to run without CUDA; the modules pickle without their device plans (audiogan.py:936-939 saves whole modules)."""
    import io
    import audiogan_b200 as ag
    ns = R.load()
    T.manual_seed(0)
    ref = ns["Embedder"](output_size=100)
    mine = ag.Embedder(output_size=100)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    mine.load_state_dict(ref.state_dict())
    chars = T.randint(0, 256, (5, 9))
    lens = T.tensor([9, 3, 7, 1, 5])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mine(chars, lens)
    g = ag.Generator(embed_size=100, state_size=32)
    buf = io.BytesIO()
    T.save(g, buf)
    buf.seek(0)
    g2 = T.load(buf, weights_only=False)
    assert list(g2.state_dict().keys()) == list(g.state_dict().keys()) and g2._plan is None


def test_reference_class_step_matches_restated_step():
    """oracle/ref_step.py (the reference's own modules through the core step's call order: what `bench.py --impl reference`
    times) against the functional restatement the GPU parity tests use: same losses, same parameters after the step."""
    from oracle import ref_step as RS
    if RS.locate() is None:
        pytest.skip("reference definitions not present (neither /root/reference nor oracle/_ref)")
    from audiogan_b200.synthetic import step_inputs
    inp = step_inputs(3, 1600, seed=5)
    rs = RS.ReferenceStep()
    got = rs.step(inp)
    Pg, Pd = O.pin_stopper(O.init_generator(11)), O.init_discriminator(12)
    o1 = O.d_update(Pg, Pd, {}, inp)
    o2 = O.g_update(Pg, Pd, {}, {"c_g": inp["g_c_g"], "c_d": inp["g_c_d"], "z": inp["g_z"], "noise_fake": inp["g_noise_fake"]})
    for a, b in zip(got, (o1["loss_d"], o1["loss_g"], o2["loss"])):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (got, o1["loss_d"], o1["loss_g"], o2["loss"])
    for sd, P in ((rs.d.state_dict(), Pd), (rs.g.state_dict(), Pg)):
        for k, v in sd.items():
            if k.split(".")[-1].startswith("bias") and k.endswith("_v"):
                continue                   # w = g * sign(v): d/dv is identically 0, both sides step on rounding noise
            err = (v - P[k]).abs()
            # the first RMSprop step is sign-like: elements whose gradient is rounding noise may step the other way
            assert float((err > 5e-6).float().mean()) < 1e-3 and float(err.max()) < 2.1e-3, (k, float(err.max()))


def test_feed_collation_and_checkpoint_names_match_reference_conventions(tmp_path):
    """feed.make_sample / collate against dataset.py:48-71 (zero row of maxlen, peak normalisation, length rounded up to the
    frame, over-long and silent waveforms rejected) and the checkpoint round trip with the reference's file names and keys."""
    import numpy as np
    import audiogan_b200 as ag
    from audiogan_b200 import feed
    rng = np.random.default_rng(0)
    w = np.concatenate([rng.standard_normal(1234) * 3.0, np.zeros(50)])
    row, ln = feed.make_sample(w, 2000, frame_size=200)
    assert row.shape == (2000,) and row.dtype == np.float32 and ln == 1400 and abs(float(np.abs(row).max()) - 1.0) < 1e-6
    assert np.all(row[1234:] == 0) and np.allclose(row[:1234], (w[:1234] / np.abs(w).max()).astype(np.float32))
    assert feed.make_sample(w, 1000) == (None, None) and feed.make_sample(np.zeros(10), 100) == (None, None)
    assert feed.make_sample(w, 2000)[1] == 1234
    b = feed.collate([(row, ln), (row, ln)], words=["hello", "hi"])
    assert b["real"].shape == (2, 2000) and b["real_len"].tolist() == [1400, 1400]
    assert b["chars"].shape == (2, 5) and b["chars"][1].tolist() == [104, 105, 0, 0, 0] and b["char_len"].tolist() == [5, 2]
    g = ag.Generator(embed_size=100, state_size=32)
    d = ag.Discriminator(embed_size=100, state_size=32)
    paths = feed.save_checkpoint(str(tmp_path / "m"), 500, g=g, d=d)
    assert paths["g"].endswith("m-gen-00500") and paths["d"].endswith("m-dis-00500")          # audiogan.py:936-937
    g2 = ag.Generator(embed_size=100, state_size=32)
    d2 = ag.Discriminator(embed_size=100, state_size=32)
    feed.load_checkpoint(str(tmp_path / "m"), 500, g=g2, d=d2)
    for a, b_ in zip(list(g.parameters()) + list(d.parameters()), list(g2.parameters()) + list(d2.parameters())):
        assert T.equal(a, b_)
    if R.available():                                  # the reference's own module takes the file as it is
        ns = R.load()
        ref = ns["Generator"](embed_size=100, state_size=32)
        ref.load_state_dict(T.load(paths["g"]))
