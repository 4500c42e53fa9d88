"""The recurrent kernels against PLAIN TORCH at the benched sequence lengths (the TMEM-resident kernels were
only compared with the repo's own fp32 kernels):

  * discriminator BiLSTM (csrc/lstm_cluster.cu in bf16 mode, csrc/lstm.cu in fp32 mode) vs torch.nn.LSTM(bidirectional)
    over packed sequences on the CPU -- what audiogan.py:214-229 / :498-503 run -- at T = 250, mixed lengths;
  * generator recurrence (csrc/lstm_gen.cu in bf16 mode, csrc/lstm.cu in fp32 mode) vs a torch loop of the cell with
    output feedback (audiogan.py:437-444) under autograd at T = 80.

Forward states AND the backward products (d pre-activations, d input, weight / bias gradients assembled from the kernels'
outputs in fp32) are compared.  fp32 mode <= 1e-4, bf16 mode <= 2e-2 forward / 3e-2 backward (relative to each tensor's max)."""
import os
import warnings

import pytest
import torch as T
import torch.nn as NN
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-30)


def _report(name, rows):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_%s.txt" % name), "w") as f:
        f.write("".join("%-40s %.3e\n" % r for r in rows))


@pytest.mark.parametrize("prec,H,B,Tn", [(0, 512, 24, 250), (1, 512, 24, 250), (1, 512, 70, 250), (1, 256, 9, 40)])
def test_bilstm_kernels_vs_torch_nn_lstm(prec, H, B, Tn):
    from audiogan_b200 import kernels as Kn
    T.manual_seed(17)
    I = 96
    rnn = NN.LSTM(I, H, 1, bidirectional=True)
    x = T.randn(Tn, B, I)
    lens = T.randint(Tn // 4, Tn + 1, (B,))
    lens[0] = Tn
    x_r = x.clone().requires_grad_(True)
    out, _ = rnn(pack_padded_sequence(x_r, lens, enforce_sorted=False))
    out = pad_packed_sequence(out, total_length=Tn)[0]                      # [T, B, 2H], zeros past each length
    mask = (T.arange(Tn)[:, None] < lens[None, :]).float()[:, :, None]
    dh = T.randn(Tn, B, 2 * H) * mask
    (out * dh).sum().backward()
    # ---- kernels: hoisted input projection (plain torch fp32 here: it is not what is under test), then the recurrence
    dev = "cuda"
    wih = T.cat([rnn.weight_ih_l0, rnn.weight_ih_l0_reverse], 0).detach().to(dev)           # [8H, I]
    bias = T.cat([rnn.bias_ih_l0 + rnn.bias_hh_l0, rnn.bias_ih_l0_reverse + rnn.bias_hh_l0_reverse], 0).detach().to(dev)
    w1 = T.stack([rnn.weight_hh_l0, rnn.weight_hh_l0_reverse], 0).detach().to(dev).contiguous()   # [2, 4H, H]
    w1t = w1.permute(0, 2, 1).contiguous()
    xb = x.permute(1, 0, 2).contiguous().to(dev)                                               # [B, T, I]
    pre = (xb @ wih.t() + bias).contiguous()                                                   # [B, T, 8H]
    lens_d = lens.to(dev, T.int32)
    hbuf, gates, cbuf = T.zeros(B, Tn + 2, 2 * H, device=dev), T.zeros(B, Tn, 8 * H, device=dev), T.zeros(B, Tn, 2 * H, device=dev)
    misc = T.zeros(16, dtype=T.int32, device=dev)
    hbuf16 = T.zeros(B, Tn + 2, 2 * H, device=dev, dtype=T.bfloat16) if prec else None
    Kn.lstm_fwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=2, F=0, pre=pre, w1=w1, hbuf=hbuf, gates=gates, cbuf=cbuf, len=lens_d,
                barrier=misc, prec=prec, flags=0, hbuf16=hbuf16)
    path_f = Kn.lstm_last_path()
    dh_ext = dh.permute(1, 0, 2).contiguous().to(dev)
    dgates = T.full((B, Tn, 8 * H), float("nan"), device=dev)
    dgates16 = T.zeros(B, Tn, 8 * H, device=dev, dtype=T.bfloat16) if prec else None
    Kn.lstm_bwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=2, F=0, gates=gates, cbuf=cbuf, len=lens_d, dh_ext=dh_ext, dgates=dgates,
                w1t=w1t, barrier=misc, prec=prec, flags=0, dgates16=dgates16)
    path_b = Kn.lstm_last_path()
    T.cuda.synchronize()
    if prec:
        assert path_f == "cluster" and path_b == "cluster", (path_f, path_b)
    else:
        assert path_f.startswith("grid-fp32") and path_b.startswith("grid-fp32"), (path_f, path_b)
    h = hbuf[:, 1:Tn + 1]                                                                      # [B, T, 2H]
    assert bool(T.isfinite(dgates).all())
    # gradients assembled from the kernel's d(pre-activations) in fp32 torch
    dx = dgates @ wih                                                                          # [B, T, I]
    dwih = dgates.reshape(B * Tn, 8 * H).t() @ xb.reshape(B * Tn, I)
    dbias = dgates.sum((0, 1))
    dwhh_f = dgates[:, :, :4 * H].reshape(B * Tn, 4 * H).t() @ hbuf[:, 0:Tn, :H].reshape(B * Tn, H)          # h_{t-1}
    dwhh_r = dgates[:, :, 4 * H:].reshape(B * Tn, 4 * H).t() @ hbuf[:, 2:Tn + 2, H:].reshape(B * Tn, H)      # h_{t+1}
    tf, tb = (1e-4, 1e-4) if prec == 0 else (2e-2, 3e-2)
    rows = [("h (T=%d)" % Tn, rel(h, out.permute(1, 0, 2))),
            ("dx", rel(dx, x_r.grad.permute(1, 0, 2))),
            ("dW_ih", rel(dwih, T.cat([rnn.weight_ih_l0.grad, rnn.weight_ih_l0_reverse.grad], 0))),
            ("db", rel(dbias, T.cat([rnn.bias_ih_l0.grad, rnn.bias_ih_l0_reverse.grad], 0))),
            ("dW_hh fwd", rel(dwhh_f, rnn.weight_hh_l0.grad)),
            ("dW_hh rev", rel(dwhh_r, rnn.weight_hh_l0_reverse.grad))]
    _report("bilstm_vs_torch_prec%d_H%d_B%d_T%d" % (prec, H, B, Tn), rows)
    assert rows[0][1] <= tf, rows
    for nm, r in rows[1:]:
        assert r <= tb, rows
    # packed-sequence semantics: exactly zero past each sample's length
    past = (T.arange(Tn, device=dev)[None, :, None] >= lens_d[:, None, None])
    assert float((h * past).abs().max()) == 0.0 and float((dgates * past).abs().max()) == 0.0


@pytest.mark.parametrize("prec,B,Tn", [(0, 24, 80), (1, 24, 80), (1, 64, 80)])
def test_generator_recurrence_kernels_vs_torch_loop(prec, B, Tn):
    from audiogan_b200 import kernels as Kn
    T.manual_seed(23)
    H, Fr = 1024, 200
    FP = (Fr + 1 + 7) // 8 * 8
    sc = 1.0 / H ** 0.5
    pre = (T.randn(B, Tn, 4 * H) * 0.5).requires_grad_(True)
    w1 = ((T.rand(4 * H, H + Fr) * 2 - 1) * sc).requires_grad_(True)           # [whh | wx]
    w2 = ((T.rand(Fr + 1, H) * 2 - 1) * sc).requires_grad_(True)               # [wp ; ws]
    b2 = ((T.rand(Fr + 1) * 2 - 1) * sc).requires_grad_(True)
    dx_ext, ds_ext = T.randn(B, Tn, Fr), T.randn(B, Tn)
    # ---- torch loop (audiogan.py:437-444) on the CPU, fp32, autograd
    h, c, x = T.zeros(B, H), T.zeros(B, H), T.zeros(B, Fr)
    hs, xs, ss = [], [], []
    for t in range(Tn):
        gts = pre[:, t] + h @ w1[:, :H].t() + x @ w1[:, H:].t()
        i, f, g, o = gts.chunk(4, 1)
        c = T.sigmoid(f) * c + T.sigmoid(i) * T.tanh(g)
        h = T.sigmoid(o) * T.tanh(c)
        x = T.tanh(h @ w2[:Fr].t() + b2[:Fr])
        s = h @ w2[Fr] + b2[Fr]
        hs.append(h); xs.append(x); ss.append(s)
    hs, xs, ss = T.stack(hs, 1), T.stack(xs, 1), T.stack(ss, 1)
    ((xs * dx_ext).sum() + (ss * ds_ext).sum()).backward()
    # ---- kernels
    dev = "cuda"
    cu = lambda v: v.detach().to(dev).contiguous()
    w1d, w2d, b2d = cu(w1).unsqueeze(0).contiguous(), cu(w2), cu(b2)
    hbuf, gates, cbuf = T.zeros(B, Tn + 2, H, device=dev), T.zeros(B, Tn, 4 * H, device=dev), T.zeros(B, Tn, H, device=dev)
    xbuf, sbuf = T.zeros(B, Tn + 1, Fr, device=dev), T.zeros(B, Tn, device=dev)
    stop, glen = T.zeros(B, Tn, dtype=T.int32, device=dev), T.zeros(B, dtype=T.int32, device=dev)
    misc = T.zeros(1024, dtype=T.int32, device=dev)
    hbuf16 = T.zeros(B, Tn + 2, H, device=dev, dtype=T.bfloat16) if prec else None
    xbuf16 = T.zeros(B, Tn + 1, Fr, device=dev, dtype=T.bfloat16) if prec else None
    ws = Kn.lstm_workspace(B, H, Fr, False, dev) if prec else None
    Kn.lstm_fwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=1, F=Fr, pre=cu(pre), w1=w1d, w2=w2d, b2=b2d, hbuf=hbuf, gates=gates, cbuf=cbuf,
                xbuf=xbuf, sbuf=sbuf, u=None, stop=stop, glen=glen, t_end=(misc, 8), barrier=misc, prec=prec, flags=2,
                hbuf16=hbuf16, xbuf16=xbuf16, ll_ws=ws, ll_ws_bytes=ws.numel() if ws is not None else 0)
    path_f = Kn.lstm_last_path()
    w1t = T.cat([w1d[0, :, :H].t(), w2d[:Fr].t(), w2d[Fr:].t(), T.zeros(H, FP - Fr - 1, device=dev)], 1).contiguous()
    wxt = w1d[0, :, H:].t().contiguous()
    dgates, dpx = T.zeros(B, Tn, 4 * H, device=dev), T.zeros(B, Tn, FP, device=dev)
    dgates16 = T.zeros(B, Tn, 4 * H, device=dev, dtype=T.bfloat16) if prec else None
    dpx16 = T.zeros(B, Tn, FP, device=dev, dtype=T.bfloat16) if prec else None
    wb = Kn.lstm_workspace(B, H, Fr, True, dev) if prec else None
    Kn.lstm_bwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=1, F=Fr, gates=gates, cbuf=cbuf, xbuf=xbuf, dx_ext=cu(dx_ext), ds_ext=cu(ds_ext),
                dgates=dgates, dpx=dpx, w1t=w1t, wxt=wxt, barrier=T.zeros(1024, dtype=T.int32, device=dev), prec=prec, flags=2,
                dgates16=dgates16, dpx16=dpx16, ll_ws=wb, ll_ws_bytes=wb.numel() if wb is not None else 0)
    path_b = Kn.lstm_last_path()
    T.cuda.synchronize()
    if prec:
        assert path_f == "tmem" and path_b == "tmem", (path_f, path_b)
    else:
        assert path_f.startswith("grid-fp32") and path_b.startswith("grid-fp32"), (path_f, path_b)
    M = B * Tn
    dw2 = dpx[:, :, :Fr + 1].reshape(M, Fr + 1).t() @ hbuf[:, 1:Tn + 1].reshape(M, H)
    db2 = dpx[:, :, :Fr + 1].sum((0, 1))
    hx_prev = T.cat([hbuf[:, 0:Tn], xbuf[:, 0:Tn]], 2).reshape(M, H + Fr)                       # [h_{t-1} | x_{t-1}]
    dw1 = dgates.reshape(M, 4 * H).t() @ hx_prev
    tf, tb = (1e-4, 2e-4) if prec == 0 else (2e-2, 3e-2)
    rows = [("h (T=%d)" % Tn, rel(hbuf[:, 1:Tn + 1], hs)), ("x frames", rel(xbuf[:, 1:Tn + 1], xs)), ("stop logits", rel(sbuf, ss)),
            ("d pre (dgates)", rel(dgates, pre.grad)), ("d[whh|wx]", rel(dw1, w1.grad)), ("d[wp;ws]", rel(dw2, w2.grad)),
            ("d[bp;bs]", rel(db2, b2.grad))]
    _report("gen_recurrence_vs_torch_prec%d_B%d_T%d" % (prec, B, Tn), rows)
    for nm, r in rows[:3]:
        assert r <= tf, rows
    for nm, r in rows[3:]:
        assert r <= tb, rows


def test_fused_adam_matches_torch_optim_adam():
    """ag_mt_adam (north_star: "fused Adam update"; the obsolete TF path's optimizer, computation_graph.py:58-59) against
    torch.optim.Adam over several steps, odd sizes (vector + tail paths), with and without the per-tensor clip."""
    import audiogan_b200 as ag
    T.manual_seed(5)
    shapes = [(1000, 37), (65536 * 2 + 5,), (3,), (128, 64, 7)]
    for clip in (0.0, 0.5):
        ps = [NN.Parameter(T.randn(*s, device="cuda")) for s in shapes]
        qs = [NN.Parameter(p.detach().clone()) for p in ps]
        ref = T.optim.Adam(qs, lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
        opt = ag.FusedRMSprop(ps, lr=1e-3, adam=True, betas=(0.9, 0.999), eps=1e-8)
        for it in range(5):
            gs = [T.randn_like(p) * (10.0 ** (it - 2)) for p in ps]
            for p, q, g in zip(ps, qs, gs):
                p.grad = g.clone()
                g2 = g.clone()
                if clip > 0:                                    # the reference's per-tensor clip (audiogan.py:243-253)
                    n = float(g2.norm())
                    if n > clip:
                        g2 /= (n / clip)
                q.grad = g2
            opt.step(clip=clip)
            ref.step()
            for p, q in zip(ps, qs):
                assert rel(p, q) < 2e-6, (clip, it, tuple(p.shape), rel(p, q))
        for s1, s2, q in zip(opt.s1, opt.s2, qs):
            st = ref.state[q]
            assert rel(s1, st["exp_avg"]) < 1e-5 and rel(s2, st["exp_avg_sq"]) < 1e-5


def test_embedder_on_gpu_matches_reference_golden():
    """Embedder (audiogan.py:302-334) on the GPU against the output and gradients of the REFERENCE's own
    class (tests/golden/aux/embedder.pt, written by oracle/make_golden.py)."""
    import audiogan_b200 as ag
    gold = T.load(os.path.join(os.path.dirname(__file__), "golden", "aux", "embedder.pt"))
    e = ag.Embedder(output_size=100)
    assert list(e.state_dict().keys()) == list(gold["state_dict"].keys())
    e.load_state_dict(gold["state_dict"])
    e = e.cuda()
    c = e(gold["chars"].cuda(), gold["lens"].cuda())
    assert rel(c, gold["c"]) < 1e-5
    (c * gold["up"].cuda()).sum().backward()
    for k, p in e.named_parameters():
        assert rel(p.grad, gold["grads"][k]) < 5e-5, k
