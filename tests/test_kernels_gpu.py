"""GPU unit tests of individual C-ABI kernels against plain PyTorch fp32 (CPU) references."""
import pytest
import torch as T
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-30)


def test_library_loads_on_gpu():
    from audiogan_b200 import _abi
    sm, smem, cc = _abi.device_info()
    assert cc >= 100 and sm > 0 and smem >= 200 * 1024


@pytest.mark.parametrize("M,N,K", [(300, 70, 45), (128, 16, 7), (1000, 200, 333), (64, 1, 512)])
def test_gemm_nt_epilogues(M, N, K):
    from audiogan_b200 import kernels as Kn
    T.manual_seed(0)
    A, B, bias, skip = T.randn(M, K), T.randn(N, K), T.randn(N), T.randn(M, N)
    ref = F.leaky_relu(A @ B.t() + bias + skip, 0.01)
    Ad, Bd, C = A.cuda(), B.cuda(), T.empty(M, N, device="cuda")
    Kn.gemm_nt(M, N, K, Ad, (M, 0, K), Bd, K, C, (M, 0, N), bias=bias.cuda(), skip=skip.cuda(), act=1)
    assert rel(C, ref) < 1e-5
    dact = T.randn(M, N)
    Kn.gemm_nt(M, N, K, Ad, (M, 0, K), Bd, K, C, (M, 0, N), dact=dact.cuda())
    assert rel(C, (A @ B.t()) * T.where(dact > 0, 1.0, 0.01)) < 1e-5


def test_gemm_nt_conv_view_and_mask():
    """Strided conv k7 s2 p3 on a zero-padded channel-last buffer == F.conv1d (audiogan.py:531-536)."""
    from audiogan_b200 import kernels as Kn
    T.manual_seed(1)
    Bn, Cin, Cout, Tin, k, s, p = 3, 5, 9, 37, 7, 2, 3
    x, w, b = T.randn(Bn, Cin, Tin), T.randn(Cout, Cin, k), T.randn(Cout)
    lens = T.tensor([19, 7, 12], dtype=T.int32)
    Tout = (Tin + s - 1) // s
    ref = F.leaky_relu(F.conv1d(x, w, b, stride=s, padding=p), 0.01)
    ref = ref * (T.arange(Tout)[None, None, :] < lens[:, None, None]).float()
    xp = T.zeros(Bn, Tin + 2 * p, Cin)
    xp[:, p:p + Tin] = x.permute(0, 2, 1)
    wp = w.permute(0, 2, 1).reshape(Cout, k * Cin).contiguous()
    out = T.zeros(Bn, Tout, Cout, device="cuda")
    Kn.gemm_nt(Bn * Tout, Cout, k * Cin, xp.cuda(), (Tout, (Tin + 2 * p) * Cin, s * Cin), wp.cuda(), k * Cin,
               out, (Tout, Tout * Cout, Cout), bias=b.cuda(), act=1, mask_len=lens.cuda(), mask=(1, 0, 0))
    assert rel(out.permute(0, 2, 1), ref) < 1e-5


def test_gemm_tn_with_bias_column():
    from audiogan_b200 import kernels as Kn
    T.manual_seed(2)
    M, N, K = 777, 50, 131
    Y, A = T.randn(M, N), T.randn(M, K)
    dw = T.zeros(N, K + 1, device="cuda")
    Kn.gemm_tn(M, N, K, Y.cuda(), (M, 0, N), A.cuda(), (M, 0, K), dw, K + 1, ones_col=True)
    assert rel(dw[:, :K], Y.t() @ A) < 1e-5
    assert rel(dw[:, K], Y.sum(0)) < 1e-5


@pytest.mark.parametrize("H,B,Tn", [(16, 3, 9), (64, 5, 12), (512, 4, 6)])
def test_lstm_bidirectional_fwd_bwd(H, B, Tn):
    """Persistent BiLSTM with per-sample lengths == NN.LSTM on packed sequences (audiogan.py:214-229)."""
    from audiogan_b200 import kernels as Kn
    from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence
    T.manual_seed(3)
    I = 11
    rnn = T.nn.LSTM(I, H, 1, bidirectional=True)
    x = T.randn(Tn, B, I, requires_grad=True)
    lens = T.tensor(sorted([Tn] + [int(v) for v in T.randint(1, Tn + 1, (B - 1,))], reverse=True))
    out, _ = rnn(pack_padded_sequence(x, lens))
    out = pad_packed_sequence(out, total_length=Tn)[0]            # (T, B, 2H)
    gout = T.randn_like(out)
    out.backward(gout)
    ps = dict(rnn.named_parameters())
    # hoisted input projection on the host reference side (the kernel takes `pre`)
    pre = T.cat([x.detach() @ ps["weight_ih_l0" + s].t() + ps["bias_ih_l0" + s] + ps["bias_hh_l0" + s]
                 for s in ("", "_reverse")], 2).permute(1, 0, 2).contiguous().detach()      # (B, T, 8H)
    w1 = T.stack([ps["weight_hh_l0"], ps["weight_hh_l0_reverse"]], 0).detach().contiguous()
    w1t = w1.permute(0, 2, 1).contiguous()
    dev = "cuda"
    hbuf = T.zeros(B, Tn + 2, 2 * H, device=dev)
    gates = T.empty(B, Tn, 8 * H, device=dev)
    cbuf = T.empty(B, Tn, 2 * H, device=dev)
    misc = T.zeros(16, dtype=T.int32, device=dev)
    lens_d = lens.to(T.int32).to(dev)
    Kn.lstm_fwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=2, F=0, pre=pre.to(dev), w1=w1.to(dev), hbuf=hbuf, gates=gates, cbuf=cbuf,
                len=lens_d, barrier=misc)
    assert rel(hbuf[:, 1:Tn + 1], out.permute(1, 0, 2)) < 1e-5
    dgates = T.empty(B, Tn, 8 * H, device=dev)
    Kn.lstm_bwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=2, F=0, gates=gates, cbuf=cbuf, len=lens_d,
                dh_ext=gout.permute(1, 0, 2).contiguous().to(dev), dgates=dgates, w1t=w1t.to(dev), barrier=misc)
    # d pre = dgates; check through dx = dgates @ W_ih and the bias gradient
    dg = dgates.cpu()
    wih = T.cat([ps["weight_ih_l0"], ps["weight_ih_l0_reverse"]], 0).detach()
    assert rel((dg @ wih).permute(1, 0, 2), x.grad) < 2e-5
    assert rel(dg[..., :4 * H].sum((0, 1)), ps["bias_ih_l0"].grad) < 2e-5
    hprev = T.zeros(B, Tn, H)
    hprev[:, 1:] = hbuf[:, 1:Tn, :H].cpu()
    assert rel(T.einsum("btr,bth->rh", dg[..., :4 * H], hprev), ps["weight_hh_l0"].grad) < 2e-5


@pytest.mark.parametrize("H,B,Tn,Fr", [(32, 3, 5, 8), (64, 9, 7, 200)])
def test_lstm_feedback_fwd_bwd(H, B, Tn, Fr):
    """Generator recurrence with output feedback, proj + tanh and stop logit (audiogan.py:437-444)."""
    from audiogan_b200 import kernels as Kn
    T.manual_seed(4)
    sc = 1.0 / H ** 0.5
    whh, wx = (T.randn(4 * H, H) * sc).requires_grad_(), (T.randn(4 * H, Fr) * sc).requires_grad_()
    wp, bp = (T.randn(Fr, H) * sc).requires_grad_(), (T.randn(Fr) * sc).requires_grad_()
    ws, bs = (T.randn(1, H) * sc).requires_grad_(), (T.randn(1) * sc).requires_grad_()
    pre = T.randn(B, Tn, 4 * H).requires_grad_()
    h, c, x = T.zeros(B, H), T.zeros(B, H), T.zeros(B, Fr)
    xs, ss, hs = [], [], []
    for t in range(Tn):
        gt = pre[:, t] + h @ whh.t() + x @ wx.t()
        i, f, g_, o = gt.chunk(4, 1)
        c = T.sigmoid(f) * c + T.sigmoid(i) * T.tanh(g_)
        h = T.sigmoid(o) * T.tanh(c)
        x = T.tanh(h @ wp.t() + bp)
        xs.append(x); ss.append((h @ ws.t() + bs).squeeze(1)); hs.append(h)
    X, S = T.stack(xs, 1), T.stack(ss, 1)
    gX, gS = T.randn_like(X), T.randn_like(S)
    (X * gX).sum().add((S * gS).sum()).backward()
    dev = "cuda"
    FP = (Fr + 1 + 7) // 8 * 8
    w1 = T.cat([whh, wx], 1).detach().contiguous().to(dev)
    w2 = T.cat([wp, ws], 0).detach().contiguous().to(dev)
    b2 = T.cat([bp, bs], 0).detach().contiguous().to(dev)
    w1t = T.cat([whh.t(), wp.t(), ws.t(), T.zeros(H, FP - Fr - 1)], 1).detach().contiguous().to(dev)
    wxt = wx.t().detach().contiguous().to(dev)
    hbuf, xbuf = T.zeros(B, Tn + 2, H, device=dev), T.zeros(B, Tn + 1, Fr, device=dev)
    gates, cbuf = T.empty(B, Tn, 4 * H, device=dev), T.empty(B, Tn, H, device=dev)
    sbuf = T.zeros(B, Tn, device=dev)
    stop = T.zeros(B, Tn, dtype=T.int32, device=dev)
    glen = T.zeros(B, dtype=T.int32, device=dev)
    misc = T.zeros(16, dtype=T.int32, device=dev)
    Kn.lstm_fwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=1, F=Fr, pre=pre.detach().to(dev), w1=w1, w2=w2, b2=b2, hbuf=hbuf, gates=gates,
                cbuf=cbuf, xbuf=xbuf, sbuf=sbuf, stop=stop, glen=glen, t_end=(misc, 8), barrier=misc)
    assert rel(xbuf[:, 1:], X) < 1e-5 and rel(sbuf, S) < 1e-5
    assert int(misc[8]) == Tn and glen.tolist() == [Tn] * B
    dgates, dpx = T.empty(B, Tn, 4 * H, device=dev), T.empty(B, Tn, FP, device=dev)
    Kn.lstm_bwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=1, F=Fr, gates=gates, cbuf=cbuf, xbuf=xbuf, dx_ext=gX.to(dev),
                ds_ext=gS.to(dev), dgates=dgates, dpx=dpx, w1t=w1t, wxt=wxt, barrier=misc)
    assert rel(dgates, pre.grad) < 2e-5
    dpx_c = dpx.cpu()
    H_all = T.stack(hs, 1).detach()
    assert rel(T.einsum("btp,bth->ph", dpx_c[..., :Fr], H_all), wp.grad) < 2e-5
    assert rel(dpx_c[..., Fr].sum(), bs.grad.sum()) < 2e-5
    # early exit: uniforms below sigmoid(logit) at step 1 for every sample -> two frames
    u = T.ones(B, Tn, device=dev)
    u[:, 1] = 0.0
    hbuf.zero_(); xbuf.zero_()
    Kn.lstm_fwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=1, F=Fr, pre=pre.detach().to(dev), w1=w1, w2=w2, b2=b2, hbuf=hbuf, gates=gates,
                cbuf=cbuf, xbuf=xbuf, sbuf=sbuf, u=u, stop=stop, glen=glen, t_end=(misc, 8), barrier=misc)
    assert int(misc[8]) == 2 and glen.tolist() == [2] * B and stop[:, 1].tolist() == [1] * B


def test_weight_norm_bce_optimizer():
    import audiogan_b200 as ag
    T.manual_seed(5)
    x, t = T.randn(4, 9), T.rand(4, 9)
    w = (T.arange(9)[None] < T.tensor([9, 5, 1, 7])[:, None]).float()
    xr = x.clone().requires_grad_()
    mx = (-xr).clamp(min=0)
    ref = ((xr - xr * t + mx + ((-mx).exp() + (-xr - mx).exp()).log()) * w).sum(1)
    ref.backward(T.arange(1.0, 5.0))
    xd = x.cuda().requires_grad_()
    got = ag.binary_cross_entropy_with_logits_per_sample(xd, t.cuda(), w.cuda())
    got.backward(T.arange(1.0, 5.0).cuda())
    assert rel(got, ref) < 1e-6 and rel(xd.grad, xr.grad) < 1e-6
    with pytest.raises(ValueError):
        ag.binary_cross_entropy_with_logits_per_sample(xd, t.cuda()[:, :3])
    # fused RMSprop + per-tensor clip == reference clip_grad then torch.optim.RMSprop
    ps = [T.nn.Parameter(T.randn(70000)), T.nn.Parameter(T.randn(33, 5)), T.nn.Parameter(T.randn(7))]
    qs = [T.nn.Parameter(p.detach().clone().cuda()) for p in ps]
    opt_r = T.optim.RMSprop(ps, lr=1e-4)
    opt = ag.FusedRMSprop(qs, lr=1e-4)
    for _ in range(3):
        total = 0.0
        for p, q in zip(ps, qs):
            gr = T.randn_like(p) * (0.001 if p.numel() == 7 else 1.0)
            p.grad, q.grad = gr.clone(), gr.clone().cuda()
            n = float(p.grad.norm()); total += n
            if n > 0.5:
                p.grad /= n / 0.5
        opt_r.step()
        norm = opt.step(clip=0.5, check=True)
        assert abs(float(norm) - total) < 1e-4 * total
    for p, q in zip(ps, qs):
        assert rel(q, p) < 1e-6
    # clip_grad / check_grad as separate reference-style calls
    for q in qs:
        q.grad = T.randn_like(q) * 3
    before = [q.grad.clone() for q in qs]
    tot = ag.clip_grad(qs, 0.25)
    ag.check_grad(qs)
    assert abs(float(tot) - sum(float(b.norm()) for b in before)) < 1e-3
    for q, b in zip(qs, before):
        n = float(b.norm())
        assert rel(q.grad, b / (n / 0.25) if n > 0.25 else b) < 1e-6


@pytest.mark.parametrize("M,N,K,vec", [(300, 70, 64, True), (1000, 256, 448, True), (517, 512, 1024, True),
                                       (260, 16, 7, False), (129, 40, 153, False), (64, 1, 512, True)])
def test_gemm_nt_tc_matches_bf16_reference(M, N, K, vec):
    """tcgen05 path: operands rounded to bf16, fp32 accumulation in TMEM -> equals an fp32 matmul of the
    bf16-rounded operands up to summation order."""
    from audiogan_b200 import kernels as Kn
    T.manual_seed(7)
    Kp = (K + 7) // 8 * 8
    A, B, bias, skip = T.randn(M, K), T.randn(N, K), T.randn(N), T.randn(M, N)
    Ab, Bb = A.bfloat16().float(), B.bfloat16().float()
    ref = F.leaky_relu(Ab @ Bb.t() + bias + skip, 0.01)
    Bd = T.zeros(N, Kp, dtype=T.bfloat16, device="cuda")
    Bd[:, :K] = B.cuda().bfloat16()
    C = T.empty(M, N, device="cuda")
    Kn.gemm_nt(M, N, K, A.cuda(), (M, 0, K), Bd, Kp, C, (M, 0, N), bias=bias.cuda(), skip=skip.cuda(), act=1, tc=True)
    assert rel(C, ref) < 2e-5
    # bf16 A operand and bf16 output
    Cb = T.empty(M, N, device="cuda", dtype=T.bfloat16)
    A16 = T.zeros(M, Kp, dtype=T.bfloat16, device="cuda")
    A16[:, :K] = A.cuda().bfloat16()
    Kn.gemm_nt(M, N, K, A16, (M, 0, Kp), Bd, Kp, Cb, (M, 0, N), tc=True)
    assert rel(Cb, (Ab @ Bb.t())) < 1e-2


def test_gemm_nt_tc_conv_view():
    from audiogan_b200 import kernels as Kn
    T.manual_seed(8)
    Bn, Cin, Cout, Tin, k, s, p = 3, 16, 40, 301, 7, 2, 3
    x, w, b = T.randn(Bn, Cin, Tin), T.randn(Cout, Cin, k), T.randn(Cout)
    lens = T.tensor([151, 70, 12], dtype=T.int32)
    Tout = (Tin + s - 1) // s
    ref = F.leaky_relu(F.conv1d(x.bfloat16().float(), w.bfloat16().float(), b, stride=s, padding=p), 0.01)
    ref = ref * (T.arange(Tout)[None, None, :] < lens[:, None, None]).float()
    xp = T.zeros(Bn, Tin + 2 * p, Cin)
    xp[:, p:p + Tin] = x.permute(0, 2, 1)
    wp = w.permute(0, 2, 1).reshape(Cout, k * Cin).contiguous().cuda().bfloat16()
    out = T.zeros(Bn, Tout, Cout, device="cuda")
    Kn.gemm_nt(Bn * Tout, Cout, k * Cin, xp.cuda(), (Tout, (Tin + 2 * p) * Cin, s * Cin), wp, k * Cin,
               out, (Tout, Tout * Cout, Cout), bias=b.cuda(), act=1, mask_len=lens.cuda(), mask=(1, 0, 0), tc=True)
    assert rel(out.permute(0, 2, 1), ref) < 2e-5


@pytest.mark.parametrize("M,N,K", [(777, 50, 131), (4000, 512, 448), (300, 128, 64), (1030, 33, 7), (2000, 1, 339)])
def test_gemm_tn_tc_matches_bf16_reference(M, N, K):
    from audiogan_b200 import kernels as Kn
    T.manual_seed(9)
    Y, A = T.randn(M, N), T.randn(M, K)
    Yb, Ab = Y.bfloat16().float(), A.bfloat16().float()
    dw = T.zeros(N, K + 1, device="cuda")
    Kn.gemm_tn(M, N, K, Y.cuda(), (M, 0, N), A.cuda(), (M, 0, K), dw, K + 1, ones_col=True, tc=True)
    assert rel(dw[:, :K], Yb.t() @ Ab) < 2e-5
    assert rel(dw[:, K], Yb.sum(0)) < 2e-5


def test_gemm_tn_tc_conv_wgrad_view():
    """Weight gradient of a strided conv read through the im2col view (bias gradient in the extra column)."""
    from audiogan_b200 import kernels as Kn
    T.manual_seed(10)
    Bn, Cin, Cout, Tin, k, s, p = 3, 16, 40, 301, 7, 2, 3
    x = T.randn(Bn, Cin, Tin).bfloat16().float()
    Tout = (Tin + s - 1) // s
    dy = T.randn(Bn, Cout, Tout).bfloat16().float()
    w = T.zeros(Cout, Cin, k, requires_grad=True)
    bb = T.zeros(Cout, requires_grad=True)
    (F.conv1d(x, w, bb, stride=s, padding=p) * dy).sum().backward()
    xp = T.zeros(Bn, Tin + 2 * p, Cin)
    xp[:, p:p + Tin] = x.permute(0, 2, 1)
    dyc = dy.permute(0, 2, 1).contiguous().cuda()
    dw = T.zeros(Cout, k * Cin + 1, device="cuda")
    Kn.gemm_tn(Bn * Tout, Cout, k * Cin, dyc, (Tout, Tout * Cout, Cout), xp.cuda(), (Tout, (Tin + 2 * p) * Cin, s * Cin),
               dw, k * Cin + 1, ones_col=True, tc=True)
    assert rel(dw[:, :k * Cin].reshape(Cout, k, Cin).permute(0, 2, 1), w.grad) < 2e-5
    assert rel(dw[:, k * Cin], bb.grad) < 2e-5


@pytest.mark.parametrize("H,B,Tn,ndir,Fr", [(64, 5, 12, 2, 0), (512, 40, 9, 2, 0), (64, 9, 7, 1, 200), (1024, 64, 6, 1, 200)])
@pytest.mark.parametrize("prec_tc", [1, 2])
def test_lstm_bf16_mode_tracks_fp32_kernels(H, B, Tn, ndir, Fr, prec_tc):
    """prec=1 (mma.sync) / prec=2 (tcgen05, TMEM accumulator) bf16 products with fp32 state against the fp32 kernels
    on the same inputs: <= 2e-2."""
    from audiogan_b200 import kernels as Kn
    T.manual_seed(11)
    dev = "cuda"
    sc = 1.0 / H ** 0.5
    FP = (Fr + 1 + 7) // 8 * 8 if Fr else 0
    pre = T.randn(B, Tn, ndir * 4 * H, device=dev)
    w1 = (T.randn(ndir, 4 * H, H + Fr, device=dev) * sc).contiguous()
    w2 = (T.randn(Fr + 1, H, device=dev) * sc) if Fr else None
    b2 = (T.randn(Fr + 1, device=dev) * sc) if Fr else None
    lens = None if Fr else T.randint(1, Tn + 1, (B,), device=dev, dtype=T.int32)
    if lens is not None:
        lens[0] = Tn
    if Fr:
        w1t = T.cat([w1[0, :, :H].t(), w2[:Fr].t(), w2[Fr:].t(), T.zeros(H, FP - Fr - 1, device=dev)], 1).contiguous()
        wxt = w1[0, :, H:].t().contiguous()
    else:
        w1t, wxt = w1.permute(0, 2, 1).contiguous(), None
    dh_ext = None if Fr else T.randn(B, Tn, ndir * H, device=dev)
    dx_ext = T.randn(B, Tn, Fr, device=dev) if Fr else None
    res = {}
    for prec in (0, prec_tc):
        hbuf, gates, cbuf = T.zeros(B, Tn + 2, ndir * H, device=dev), T.empty(B, Tn, ndir * 4 * H, device=dev), T.empty(B, Tn, ndir * H, device=dev)
        xbuf = T.zeros(B, Tn + 1, Fr, device=dev) if Fr else None
        sbuf = T.zeros(B, Tn, device=dev) if Fr else None
        misc = T.zeros(16, dtype=T.int32, device=dev)
        kw = {}
        if prec:
            kw = dict(prec=prec, hbuf16=T.zeros(B, Tn + 2, ndir * H, device=dev, dtype=T.bfloat16),
                      xbuf16=T.zeros(B, Tn + 1, Fr, device=dev, dtype=T.bfloat16) if Fr else None)
        Kn.lstm_fwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=ndir, F=Fr, pre=pre, w1=w1, w2=w2, b2=b2, hbuf=hbuf, gates=gates, cbuf=cbuf,
                    len=lens, xbuf=xbuf, sbuf=sbuf, t_end=(misc, 8) if Fr else None, barrier=misc, **kw)
        dgates = T.empty(B, Tn, ndir * 4 * H, device=dev)
        dpx = T.empty(B, Tn, FP, device=dev) if Fr else None
        kw = {}
        if prec:
            kw = dict(prec=prec, dgates16=T.empty(B, Tn, ndir * 4 * H, device=dev, dtype=T.bfloat16),
                      dpx16=T.empty(B, Tn, FP, device=dev, dtype=T.bfloat16) if Fr else None)
        Kn.lstm_bwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=ndir, F=Fr, gates=gates, cbuf=cbuf, len=lens, xbuf=xbuf, dh_ext=dh_ext,
                    dx_ext=dx_ext, dgates=dgates, dpx=dpx, w1t=w1t, wxt=wxt, barrier=misc, **kw)
        res[prec] = (hbuf.clone(), dgates.clone(), xbuf.clone() if Fr else None, kw.get("dgates16"))
    r1 = res[prec_tc]
    assert rel(r1[0], res[0][0]) < 2e-2, "h"
    assert rel(r1[1], res[0][1]) < 3e-2, "dgates"
    if Fr:
        assert rel(r1[2], res[0][2]) < 2e-2, "x"
    assert rel(r1[3].float(), r1[1]) < 1e-2, "bf16 shadow"


@pytest.mark.parametrize("H,B,Tn", [(512, 40, 9), (128, 5, 12), (256, 70, 17), (512, 200, 5), (512, 128, 33)])
def test_lstm_cluster_kernels_track_fp32_kernels(H, B, Tn):
    """Cluster-resident BiLSTM (csrc/lstm_cluster.cu: TMEM-resident weights, DSMEM exchange; what the discriminator's
    NN.LSTM, audiogan.py:498-503, runs on in bf16 mode) against the fp32 grid-barrier kernels and against the bf16
    grid-barrier kernels (flags bit 0) on the same inputs, mixed lengths, ragged last slice, several rounds (B = 200)."""
    from audiogan_b200 import kernels as Kn, _abi as A
    assert A.lib().ag_lstm_cluster_max_active(H, 0) >= 2 and A.lib().ag_lstm_cluster_max_active(H, 1) >= 2, "cluster launch unavailable"
    T.manual_seed(5)
    dev, ndir = "cuda", 2
    pre = T.randn(B, Tn, ndir * 4 * H, device=dev)
    w1 = (T.randn(ndir, 4 * H, H, device=dev) / H ** 0.5).contiguous()
    w1t = w1.permute(0, 2, 1).contiguous()
    lens = T.randint(1, Tn + 1, (B,), device=dev, dtype=T.int32)
    lens[0] = Tn
    dh_ext = T.randn(B, Tn, ndir * H, device=dev)
    out = {}
    for name, prec, flags in (("fp32", 0, 0), ("grid", 1, 1), ("cluster", 1, 0)):
        hbuf, gates, cbuf = T.zeros(B, Tn + 2, ndir * H, device=dev), T.zeros(B, Tn, ndir * 4 * H, device=dev), T.zeros(B, Tn, ndir * H, device=dev)
        misc = T.zeros(16, dtype=T.int32, device=dev)
        dbg = T.zeros(148 * 4, 8, dtype=T.int64, device=dev) if name == "cluster" else None
        hbuf16 = T.zeros(B, Tn + 2, ndir * H, device=dev, dtype=T.bfloat16) if prec else None
        dgates = T.full((B, Tn, ndir * 4 * H), float("nan"), device=dev)
        dgates16 = T.zeros(B, Tn, ndir * 4 * H, device=dev, dtype=T.bfloat16) if prec else None
        Kn.lstm_fwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=ndir, F=0, pre=pre, w1=w1, hbuf=hbuf, gates=gates, cbuf=cbuf, len=lens,
                    barrier=misc, prec=prec, flags=flags, hbuf16=hbuf16, dbg=dbg)
        if dbg is not None:
            assert int((dbg[:, 7] > 0).sum()) > 0, "the cluster forward kernel did not run"
            dbg.zero_()
        Kn.lstm_bwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=ndir, F=0, gates=gates, cbuf=cbuf, len=lens, dh_ext=dh_ext, dgates=dgates,
                    w1t=w1t, barrier=misc, prec=prec, flags=flags, dgates16=dgates16, dbg=dbg)
        if dbg is not None:
            assert int((dbg[:, 7] > 0).sum()) > 0, "the cluster backward kernel did not run"
        out[name] = (hbuf, gates, cbuf, dgates, hbuf16, dgates16)
    f, gr, c = out["fp32"], out["grid"], out["cluster"]
    for i, nm in enumerate(("h", "gates", "c", "dgates")):
        assert bool(T.isfinite(c[i]).all()), nm
        assert rel(c[i], f[i]) < 2e-2, (nm, rel(c[i], f[i]))
        assert rel(c[i], gr[i]) < 2e-2, (nm, "vs grid bf16", rel(c[i], gr[i]))
    # steps past a sample's length emit exactly zero (packed-sequence semantics, audiogan.py:214-229)
    tt = T.arange(Tn, device=dev)[None, :, None] >= lens[:, None, None]
    assert float((c[0][:, 1:Tn + 1] * tt).abs().max()) == 0.0 and float((c[3] * tt).abs().max()) == 0.0
    assert rel(c[4].float(), c[0]) < 1e-2 and rel(c[5].float(), c[3]) < 1e-2, "bf16 shadows"


@pytest.mark.parametrize("B,Tn,stops", [(16, 3, False), (37, 12, True), (64, 9, False), (70, 7, True)])
def test_lstm_generator_tmem_kernel_tracks_fp32_kernel(B, Tn, stops):
    """TMEM-resident generator recurrence (csrc/lstm_gen.cu: weights in tensor memory + shared memory, LL exchange of
    h_t / x_t through L2) against the fp32 grid-barrier kernel on the same inputs: frames, stop logits, saved state
    <= 2e-2; with stop sampling the Bernoulli draws, lengths and the early-exit step are IDENTICAL (audiogan.py:445-460)."""
    from audiogan_b200 import kernels as Kn
    T.manual_seed(7)
    dev, H, Fr = "cuda", 1024, 200
    pre = T.randn(B, Tn, 4 * H, device=dev)
    w1 = (T.randn(1, 4 * H, H + Fr, device=dev) / H ** 0.5).contiguous()
    w2 = (T.randn(Fr + 1, H, device=dev) / H ** 0.5).contiguous()
    b2 = (T.randn(Fr + 1, device=dev) / H ** 0.5).contiguous()
    if stops:
        b2[Fr] = 0.5                                   # stop probability ~60 % per frame: every sample stops early
    u = T.rand(B, Tn, device=dev) if stops else None
    out = {}
    for name, prec, flags in (("fp32", 0, 0), ("tmem", 1, 2)):
        hbuf, gates, cbuf = T.zeros(B, Tn + 2, H, device=dev), T.zeros(B, Tn, 4 * H, device=dev), T.zeros(B, Tn, H, device=dev)
        xbuf, sbuf = T.zeros(B, Tn + 1, Fr, device=dev), T.zeros(B, Tn, device=dev)
        stop, glen = T.zeros(B, Tn, dtype=T.int32, device=dev), T.zeros(B, dtype=T.int32, device=dev)
        misc = T.zeros(1024, dtype=T.int32, device=dev)
        dbg = T.zeros(148, 8, dtype=T.int64, device=dev) if prec else None
        ll_ws = Kn.lstm_workspace(B, H, Fr, False, dev)            # sized by the library (ag_lstm_workspace_bytes)
        hbuf16 = T.zeros(B, Tn + 2, H, device=dev, dtype=T.bfloat16) if prec else None
        xbuf16 = T.zeros(B, Tn + 1, Fr, device=dev, dtype=T.bfloat16) if prec else None
        Kn.lstm_fwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=1, F=Fr, pre=pre, w1=w1, w2=w2, b2=b2, hbuf=hbuf, gates=gates, cbuf=cbuf,
                    xbuf=xbuf, sbuf=sbuf, u=u, stop=stop, glen=glen, t_end=(misc, 8), barrier=misc, prec=prec,
                    flags=flags, hbuf16=hbuf16, xbuf16=xbuf16, dbg=dbg, ll_ws=ll_ws, ll_ws_bytes=ll_ws.numel() if ll_ws is not None else 0)
        T.cuda.synchronize()
        if dbg is not None:
            assert int((dbg[:, 7] > 0).sum()) > 0, "the TMEM-resident kernel did not run"
            dbg.zero_()
        te_ = int(misc[8].item())
        # BPTT over the steps that ran (audiogan.py:437-444 through autograd in the reference)
        FP = (Fr + 1 + 7) // 8 * 8
        if name == "fp32":
            w1t = T.cat([w1[0, :, :H].t(), w2[:Fr].t(), w2[Fr:].t(), T.zeros(H, FP - Fr - 1, device=dev)], 1).contiguous()
            wxt = w1[0, :, H:].t().contiguous()
            dx_ext, ds_ext = T.randn(B, Tn, Fr, device=dev), T.randn(B, Tn, device=dev)
        dgates, dpx = T.zeros(B, Tn, 4 * H, device=dev), T.zeros(B, Tn, FP, device=dev)
        dgates16 = T.zeros(B, Tn, 4 * H, device=dev, dtype=T.bfloat16) if prec else None
        dpx16 = T.zeros(B, Tn, FP, device=dev, dtype=T.bfloat16) if prec else None
        ngr = (B + 15) // 16
        ll_wb = Kn.lstm_workspace(B, H, Fr, True, dev)
        Kn.lstm_bwd(B=B, T=te_, Tcap=Tn, H=H, ndir=1, F=Fr, gates=gates, cbuf=cbuf, xbuf=xbuf, dx_ext=dx_ext, ds_ext=ds_ext,
                    dgates=dgates, dpx=dpx, w1t=w1t, wxt=wxt, barrier=T.zeros(1024, dtype=T.int32, device=dev), prec=prec,
                    flags=flags, dgates16=dgates16, dpx16=dpx16, dbg=dbg, ll_ws=ll_wb, ll_ws_bytes=ll_wb.numel() if ll_wb is not None else 0)
        T.cuda.synchronize()
        if dbg is not None and B <= 64:
            assert int((dbg[:, 7] > 0).sum()) > 0, "the TMEM-resident BPTT kernel did not run"
        out[name] = dict(h=hbuf, x=xbuf, s=sbuf, gates=gates, c=cbuf, stop=stop, glen=glen, t_end=te_, x16=xbuf16,
                         dgates=dgates, dpx=dpx, dgates16=dgates16)
    f, c = out["fp32"], out["tmem"]
    te = f["t_end"]
    assert c["t_end"] == te and (not stops or te < Tn or B < 20)
    assert T.equal(c["glen"], f["glen"]) and T.equal(c["stop"][:, :te], f["stop"][:, :te])
    for nm in ("h", "x"):
        assert rel(c[nm][:, :te + 1], f[nm][:, :te + 1]) < 2e-2, nm
    for nm in ("s", "gates", "c"):
        assert rel(c[nm][:, :te], f[nm][:, :te]) < 2e-2, nm
    assert rel(c["x16"][:, :te + 1].float(), c["x"][:, :te + 1]) < 1e-2
    for nm in ("dgates", "dpx"):
        assert rel(c[nm][:, :te], f[nm][:, :te]) < 3e-2, nm
    assert rel(c["dgates16"][:, :te].float(), c["dgates"][:, :te]) < 1e-2


@pytest.mark.parametrize("dt", [T.float32, T.bfloat16])
def test_streaming_kernels_both_storage_types(dt):
    """ew_grad / colsum / copy3d / conv1out_* / conv1in_* with fp32 and bf16 activation storage (bf16 mode keeps the
    conv stacks' activations and gradients in HBM as bf16) against plain torch fp32 on the same (rounded) inputs."""
    from audiogan_b200 import kernels as K
    T.manual_seed(3)
    B, Tn, Cn, pl, pr = 3, 37, 24, 2, 3
    q = lambda x: x.to(dt).float()          # the values the kernel sees
    tol = 1e-6 if dt == T.float32 else 8e-3
    g1, g2, act, acc0 = (q(T.randn(B, Tn, Cn)) for _ in range(4))
    ln = T.tensor([37, 20, 1], dtype=T.int32)
    out = T.empty(B, pl + Tn + pr, Cn, device="cuda", dtype=dt)
    acc = acc0.to(dt).cuda()
    K.ew_grad(B, Tn, Cn, out=out, pad=(pl, pr), g1=g1.to(dt).cuda(), g1_str=(Tn * Cn, Cn, 1), g2=g2.to(dt).cuda(),
              g2_str=(Tn * Cn, Cn, 1), act=act.to(dt).cuda(), act_str=(Tn * Cn, Cn), length=ln.cuda(), acc=acc, acc_str=(Tn * Cn, Cn))
    m = (T.arange(Tn)[None] < ln[:, None]).float()[:, :, None]
    v = (g1 + g2) * T.where(act > 0, 1.0, 0.01) * m
    ref = T.zeros(B, pl + Tn + pr, Cn)
    ref[:, pl:pl + Tn] = v
    assert rel(out, ref) < tol and rel(acc, acc0 + v) < tol
    # fused bias gradient: column sums of the assembled gradient in the same pass (C = 4 * 2^k)
    Cs = 32
    gs_, as_ = q(T.randn(B, Tn, Cs)), q(T.randn(B, Tn, Cs))
    outs = T.empty(B, pl + Tn + pr, Cs, device="cuda", dtype=dt)
    cs = T.zeros(Cs, device="cuda")
    K.ew_grad(B, Tn, Cs, out=outs, pad=(pl, pr), g1=gs_.to(dt).cuda(), g1_str=(Tn * Cs, Cs, 1), act=as_.to(dt).cuda(),
              act_str=(Tn * Cs, Cs), colsum=cs)
    vs = gs_ * T.where(as_ > 0, 1.0, 0.01)
    assert rel(outs[:, pl:pl + Tn], vs) < tol and rel(cs, vs.sum((0, 1))) < 1e-5
    # strided (scalar-path) g1: a (B, C, T) tensor read as [b, t, c]
    g1t = g1.permute(0, 2, 1).contiguous().to(dt).cuda()
    out2 = T.empty(B, Tn, Cn, device="cuda", dtype=dt)
    K.ew_grad(B, Tn, Cn, out=out2, g1=g1t, g1_str=(Cn * Tn, 1, Tn))
    assert rel(out2, g1) < tol
    # colsum (vector and generic paths)
    for Cc in (24, 7):
        src = q(T.randn(B, Tn, Cc))
        o = T.zeros(Cc, device="cuda")
        K.colsum(src.to(dt).cuda(), Tn * Cc, Cc, B, Tn, Cc, o)
        assert rel(o, src.sum((0, 1))) < 1e-5
    # copy3d: fp32 frames -> channel 0 of a channel-last buffer of the storage type, and back (accumulating)
    fr = T.randn(B, Tn)
    buf = T.zeros(B, Tn, Cn, device="cuda", dtype=dt)
    K.copy3d(buf, (Tn * Cn, Cn, 0), fr.cuda(), (Tn, 1, 0), B, Tn, 1)
    assert rel(buf[:, :, 0], q(fr)) < 1e-6 and float(buf[:, :, 1:].abs().max()) == 0
    back = T.ones(B, Tn, device="cuda")
    K.copy3d(back, (Tn, 1, 0), buf, (Tn * Cn, Cn, 0), B, Tn, 1, accumulate=True)
    assert rel(back, q(fr) + 1) < 1e-6
    # conv1out: Conv1d(C -> 1, k = 3) over a channel-last buffer, its data and weight gradients
    k = 3
    X = q(T.randn(B, Tn + k - 1, Cn))
    w, b = T.randn(k * Cn), T.randn(1)
    y = T.empty(B, Tn, device="cuda")
    K.conv1out_fwd(X.to(dt).cuda(), (Tn + k - 1) * Cn, Cn, k, w.cuda(), b.cuda(), y, B, Tn)
    wt = w.view(k, Cn).t().unsqueeze(0)                      # [1, C, k]
    yr = F.conv1d(X.permute(0, 2, 1), wt, b)[:, 0]
    assert rel(y, yr) < 1e-5
    g = T.randn(B, Tn)
    dX = T.empty(B, Tn + k - 1, Cn, device="cuda", dtype=dt)
    K.conv1out_dgrad(g.cuda(), w.cuda(), dX, (Tn + k - 1) * Cn, Cn, k, B, Tn)
    Xr = X.clone().requires_grad_()
    wr = wt.clone().requires_grad_()
    F.conv1d(Xr.permute(0, 2, 1), wr, b)[:, 0].backward(g)
    assert rel(dX, Xr.grad) < tol
    dw = T.zeros(k * Cn + 1, device="cuda")
    K.conv1out_wgrad(g.cuda(), X.to(dt).cuda(), (Tn + k - 1) * Cn, Cn, k, dw, B, Tn)
    assert rel(dw[:-1].view(k, Cn), wr.grad[0].t()) < 1e-5 and abs(float(dw[-1]) - float(g.sum())) < 1e-4
    # the same three kernels at the dense buffer's width (C = 128: strip kernels, several strips per sample) and an odd length
    B2, T2, C2 = 2, 301, 128
    X2 = q(T.randn(B2, T2 + k - 1, C2))
    w2, g2_ = T.randn(k * C2), T.randn(B2, T2)
    y2 = T.empty(B2, T2, device="cuda")
    K.conv1out_fwd(X2.to(dt).cuda(), (T2 + k - 1) * C2, C2, k, w2.cuda(), b.cuda(), y2, B2, T2)
    wt2 = w2.view(k, C2).t().unsqueeze(0).clone().requires_grad_()
    yr2 = F.conv1d(X2.permute(0, 2, 1), wt2, b)[:, 0]
    assert rel(y2, yr2) < 1e-5
    yr2.backward(g2_)
    dw2 = T.zeros(k * C2 + 1, device="cuda")
    K.conv1out_wgrad(g2_.cuda(), X2.to(dt).cuda(), (T2 + k - 1) * C2, C2, k, dw2, B2, T2)
    assert rel(dw2[:-1].view(k, C2), wt2.grad[0].t()) < 1e-5 and abs(float(dw2[-1]) - float(g2_.sum())) < 1e-3
    # conv1in: Conv1d(1 -> C, k = 7, s = 2) + bias + LeakyReLU + mask on a zero-padded waveform, weight/bias gradient
    k, s, Co, L = 7, 2, 16, 60
    To = (L + s - 1) // s
    x = T.zeros(B, L + 6)
    x[:, 3:3 + L] = T.randn(B, L)
    w1, b1 = T.randn(Co, k), T.randn(Co)
    ln1 = T.tensor([To, 11, 2], dtype=T.int32)
    o = T.empty(B, To, Co, device="cuda", dtype=dt)
    K.conv1in_fwd(x.cuda(), L + 6, w1.cuda(), b1.cuda(), o, To * Co, k, s, Co, B, To, ln1.cuda())
    m1 = (T.arange(To)[None] < ln1[:, None]).float()[:, None]
    orf = F.leaky_relu(F.conv1d(x[:, None], w1[:, None], b1, stride=s)[:, :, :To]) * m1
    assert rel(o, orf.permute(0, 2, 1)) < tol
    dy = q(T.randn(B, To, Co))
    dw1 = T.zeros(Co, k + 1, device="cuda")
    K.conv1in_wgrad(dy.to(dt).cuda(), To * Co, x.cuda(), L + 6, dw1, k, s, Co, B, To)
    w1r, b1r = w1.clone().requires_grad_(), b1.clone().requires_grad_()
    F.conv1d(x[:, None], w1r[:, None], b1r, stride=s)[:, :, :To].backward(dy.permute(0, 2, 1))
    assert rel(dw1[:, :k], w1r.grad) < 1e-5 and rel(dw1[:, k], b1r.grad) < 1e-5


@pytest.mark.parametrize("Bn,Cin,Cout,Tin,k,s", [(3, 16, 40, 301, 7, 2), (2, 32, 256, 700, 7, 2), (5, 64, 24, 130, 3, 1),
                                                 (1, 128, 512, 1000, 7, 2)])
def test_gemm_nt_tma_conv_window_view(Bn, Cin, Cout, Tin, k, s):
    """Persistent TMA-fed kernel (bf16 activations, contiguous im2col window, rows tiled per batch): strided conv with
    bias + skip + LeakyReLU + length mask, fp32 and bf16 outputs, several tiles per CTA and a ragged last row tile."""
    from audiogan_b200 import kernels as Kn
    T.manual_seed(18)
    p = (k - 1) // 2
    x, w, b = T.randn(Bn, Cin, Tin), T.randn(Cout, Cin, k) / (Cin * k) ** 0.5, T.randn(Cout)
    Tout = (Tin + 2 * p - k) // s + 1
    lens = T.randint(1, Tout + 1, (Bn,), dtype=T.int32)
    lens[0] = Tout
    skip = T.randn(Bn, Tout, Cout).bfloat16()
    ref = F.conv1d(x.bfloat16().float(), w.bfloat16().float(), b, stride=s, padding=p).permute(0, 2, 1) + skip.float()
    ref = F.leaky_relu(ref, 0.01) * (T.arange(Tout)[None, :, None] < lens[:, None, None]).float()
    xp = T.zeros(Bn, Tin + 2 * p + s, Cin)
    xp[:, p:p + Tin] = x.permute(0, 2, 1)
    xp = xp.cuda().bfloat16()
    wp = w.permute(0, 2, 1).reshape(Cout, k * Cin).contiguous().cuda().bfloat16()
    for odt, tol in ((T.float32, 3e-5), (T.bfloat16, 8e-3)):
        out = T.zeros(Bn, Tout, Cout, device="cuda", dtype=odt)
        Kn.gemm_nt(Bn * Tout, Cout, k * Cin, xp, (Tout, xp.shape[1] * Cin, s * Cin), wp, k * Cin,
                   out, (Tout, Tout * Cout, Cout), bias=b.cuda(), skip=skip.cuda(), act=1, mask_len=lens.cuda(), mask=(1, 0, 0), tc=True)
        assert rel(out, ref) < tol


@pytest.mark.parametrize("Bn,Tm,N,K", [(4, 250, 1024, 1024), (3, 77, 512, 192), (2, 130, 8, 64)])
def test_gemm_nt_tma_flat_and_padded_rows(Bn, Tm, N, K):
    """TMA-fed kernel with a flat A ([B*Tm, K]) written into a padded per-batch C geometry ([B, Tm+2, N], rows 1..Tm),
    LeakyReLU' of a saved activation and C += skip -- the discriminator tail's data-gradient GEMMs -- and the reverse."""
    from audiogan_b200 import kernels as Kn
    T.manual_seed(19)
    M = Bn * Tm
    A, W = T.randn(M, K).bfloat16(), (T.randn(N, K) / K ** 0.5).bfloat16()
    act = T.randn(Bn, Tm + 2, N).bfloat16()
    sk = T.randn(Bn, Tm + 2, N).bfloat16()
    geo = (Tm, (Tm + 2) * N, N)
    out = T.zeros(Bn, Tm + 2, N, device="cuda", dtype=T.bfloat16)
    Kn.gemm_nt(M, N, K, A.cuda(), (M, 0, K), W.cuda(), K, (out, N), geo, skip=(sk.cuda(), N), dact=(act.cuda(), N), tc=True)
    ref = (A.float() @ W.float().t()).view(Bn, Tm, N) + sk[:, 1:Tm + 1].float()
    ref = ref * T.where(act[:, 1:Tm + 1].float() > 0, 1.0, 0.01)
    assert rel(out[:, 1:Tm + 1], ref) < 8e-3 and float(out[:, 0].abs().max()) == 0 and float(out[:, Tm + 1].abs().max()) == 0
    # padded per-batch A -> packed fp32 C with a per-batch row bias
    Ap = T.zeros(Bn, Tm + 2, K).bfloat16()
    Ap[:, 1:Tm + 1] = A.view(Bn, Tm, K)
    rb = T.randn(Bn, N)
    out2 = T.empty(M, N, device="cuda")
    Kn.gemm_nt(M, N, K, (Ap.cuda(), K), (Tm, (Tm + 2) * K, K), W.cuda(), K, out2, (Tm, Tm * N, N), rowbias=rb.cuda(), rowbias_ld=N, tc=True)
    ref2 = (A.float() @ W.float().t()).view(Bn, Tm, N) + rb[:, None]
    assert rel(out2.view(Bn, Tm, N), ref2) < 3e-5


@pytest.mark.parametrize("Bn,Cin,Cout,Tin,k,s", [(3, 16, 40, 301, 7, 2), (5, 64, 256, 700, 7, 2), (2, 128, 24, 130, 3, 1)])
def test_gemm_tn_tma_conv_wgrad_view(Bn, Cin, Cout, Tin, k, s):
    """TMA-fed weight-gradient kernel (bf16 storage, stages of 64 rows per batch, zero-filled ragged ends): strided-conv
    weight gradient through the im2col window view + the bias gradient from the column-sum kernel."""
    from audiogan_b200 import kernels as Kn
    T.manual_seed(20)
    p = (k - 1) // 2
    x = T.randn(Bn, Cin, Tin).bfloat16().float()
    Tout = (Tin + 2 * p - k) // s + 1
    dy = T.randn(Bn, Cout, Tout).bfloat16().float()
    w = T.zeros(Cout, Cin, k, requires_grad=True)
    bb = T.zeros(Cout, requires_grad=True)
    (F.conv1d(x, w, bb, stride=s, padding=p) * dy).sum().backward()
    xp = T.zeros(Bn, Tin + 2 * p + s, Cin)
    xp[:, p:p + Tin] = x.permute(0, 2, 1)
    dyc = dy.permute(0, 2, 1).contiguous().cuda().bfloat16()
    dw = T.zeros(Cout, k * Cin + 1, device="cuda")
    Kn.gemm_tn(Bn * Tout, Cout, k * Cin, dyc, (Tout, Tout * Cout, Cout), xp.cuda().bfloat16(), (Tout, xp.shape[1] * Cin, s * Cin),
               dw, k * Cin + 1, ones_col=True, tc=True)
    assert rel(dw[:, :k * Cin].reshape(Cout, k, Cin).permute(0, 2, 1), w.grad) < 3e-5
    assert rel(dw[:, k * Cin], bb.grad) < 3e-5


@pytest.mark.parametrize("Bn,Tm,N,K", [(4, 250, 1024, 1024), (3, 77, 512, 192), (2, 130, 8, 64), (6, 100, 2048, 512)])
def test_gemm_tn_tma_flat_and_padded_rows(Bn, Tm, N, K):
    """TMA-fed weight gradient with a flat Y ([B*Tm, N]) against a padded per-batch activation ([B, Tm+2, K], rows 1..Tm)
    and the reverse -- the discriminator tail's weight-gradient GEMMs."""
    from audiogan_b200 import kernels as Kn
    T.manual_seed(21)
    M = Bn * Tm
    Y, A = T.randn(M, N).bfloat16(), T.randn(M, K).bfloat16()
    ref, refb = Y.float().t() @ A.float(), Y.float().sum(0)
    Ap = T.randn(Bn, Tm + 2, K).bfloat16()
    Ap[:, 1:Tm + 1] = A.view(Bn, Tm, K)
    dw = T.zeros(N, K + 1, device="cuda")
    Kn.gemm_tn(M, N, K, Y.cuda(), (M, 0, N), (Ap.cuda(), K), (Tm, (Tm + 2) * K, K), dw, K + 1, ones_col=True, tc=True)
    assert rel(dw[:, :K], ref) < 3e-5 and rel(dw[:, K], refb) < 3e-5
    Yp = T.randn(Bn, Tm + 2, N).bfloat16()
    Yp[:, 1:Tm + 1] = Y.view(Bn, Tm, N)
    dw2 = T.zeros(N, K, device="cuda")
    Kn.gemm_tn(M, N, K, (Yp.cuda(), N), (Tm, (Tm + 2) * N, N), A.cuda(), (M, 0, K), dw2, K, tc=True)
    assert rel(dw2, ref) < 3e-5


@pytest.mark.parametrize("Bn,cin,CT,hid,L,k,s", [(3, 24, 120, 64, 1000, 9, 4), (2, 8, 120, 128, 1600, 17, 8), (2, 88, 120, 32, 520, 9, 4)])
def test_gemm_tma_channel_prefix_view(Bn, cin, CT, hid, L, k, s):
    """a_layout 1: the generator's conv over a channel PREFIX of the dense channel-last buffer through a 4-D tensor map
    {channel, tap, row, batch}; filter / gradient columns ordered (channel group, tap padded to 8, channel).  Forward
    (bias + LeakyReLU) and weight + bias gradient against conv1d on the same bf16-rounded data."""
    from audiogan_b200 import kernels as Kn
    T.manual_seed(22)
    p, Lh, PAD = (k - 1) // 2, L // s, 8
    Lp = L + 2 * PAD
    Xd = T.zeros(Bn, Lp, CT)
    Xd[:, PAD:PAD + L] = T.randn(Bn, L, CT)             # channels >= cin hold data the view must not read
    Xd = Xd.bfloat16()
    w = (T.randn(hid, cin, k) / (cin * k) ** 0.5).bfloat16()
    b = T.randn(hid)
    G_, KT = cin // 8, (k + 7) // 8
    Kq = G_ * KT * 64
    wq = T.zeros(hid, G_, KT * 8, 8)
    wq[:, :, :k] = w.float().view(hid, G_, 8, k).permute(0, 1, 3, 2)
    wq = wq.reshape(hid, Kq).bfloat16().cuda()
    x = Xd[:, PAD:PAD + L, :cin].float().permute(0, 2, 1)
    wr, br = w.float().clone().requires_grad_(), b.clone().requires_grad_()
    pre = F.conv1d(x, wr, br, stride=s, padding=p)[:, :, :Lh]
    ref = F.leaky_relu(pre, 0.01)
    Hh = T.zeros(Bn, Lh + 2, hid, device="cuda", dtype=T.bfloat16)
    Xg = Xd.cuda()
    Kn.gemm_nt(Bn * Lh, hid, k * cin, (Xg, (PAD - p) * CT), (Lh, Lp * CT, s * CT, cin, CT), wq, Kq, (Hh, hid),
               (Lh, (Lh + 2) * hid, hid), bias=b.cuda(), act=1, tc=True, a_layout=1)
    assert rel(Hh[:, 1:Lh + 1], ref.permute(0, 2, 1)) < 8e-3
    dH = T.randn(Bn, Lh, hid).bfloat16()
    pre.backward(dH.float().permute(0, 2, 1))
    dHp = T.zeros(Bn, Lh + 2, hid, dtype=T.bfloat16)
    dHp[:, 1:Lh + 1] = dH
    dw = T.zeros(hid, Kq + 1, device="cuda")
    Kn.gemm_tn(Bn * Lh, hid, k * cin, (dHp.cuda(), hid), (Lh, (Lh + 2) * hid, hid), (Xg, (PAD - p) * CT),
               (Lh, Lp * CT, s * CT, cin, CT), dw, Kq + 1, ones_col=True, tc=True, a_layout=1)
    got = dw[:, :Kq].view(hid, G_, KT * 8, 8)[:, :, :k].permute(0, 1, 3, 2).reshape(hid, cin, k)
    assert rel(got, wr.grad) < 3e-5 and rel(dw[:, Kq], br.grad) < 3e-5
    assert float(dw[:, :Kq].view(hid, G_, KT * 8, 8)[:, :, k:].abs().max()) == 0 or KT * 8 == k


@pytest.mark.parametrize("dt", [T.float32, T.bfloat16])
def test_rowdot_linear_to_one(dt):
    """ag_rowdot: forward of the classifier's Linear(512 -> 1) (audiogan.py:508-512) against torch on the same (rounded) rows"""
    from audiogan_b200 import kernels as Kn
    T.manual_seed(31)
    for M, Kd in ((1037, 512), (5, 8), (300, 1024)):
        X = T.randn(M, Kd).to(dt)
        w, b = T.randn(Kd + 3)[3:].contiguous(), T.randn(1)          # w at an odd offset: no alignment assumed
        wg = T.randn(Kd + 3).cuda()
        wg[3:] = w.cuda()
        out = T.empty(M, device="cuda")
        Kn.rowdot(X.cuda(), (wg, 3), b.cuda(), out, M, Kd)
        ref = X.double() @ w.double() + b.double()
        assert float((out.cpu().double() - ref).abs().max()) < 1e-4 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("Bn,cin,CT,hid,L,k,s", [(3, 32, 128, 64, 1000, 9, 4), (2, 64, 128, 64, 1600, 9, 4), (2, 96, 128, 32, 520, 9, 4),
                                                  (2, 88, 120, 32, 520, 9, 4), (2, 16, 128, 128, 1600, 17, 8)])
def test_gemm_tma_channel_prefix_view_wide(Bn, cin, CT, hid, L, k, s):
    """a_layout 2: the same conv over a channel prefix, through 128-byte TMA boxes {64 channels, rows, 1 tap}; filter / gradient
    columns ordered (tap, channel group of 64, channel), the channels past the prefix zero-filled by the TMA unit (the buffer
    holds data there that must not leak in).  Forward (bias + LeakyReLU), weight + bias gradient against conv1d."""
    from audiogan_b200 import kernels as Kn
    T.manual_seed(23)
    p, Lh, PAD = (k - 1) // 2, L // s, 8
    Lp = L + 2 * PAD
    Xd = T.zeros(Bn, Lp, CT)
    Xd[:, PAD:PAD + L] = T.randn(Bn, L, CT)             # channels >= cin hold data the view must not read
    Xd = Xd.bfloat16()
    w = (T.randn(hid, cin, k) / (cin * k) ** 0.5).bfloat16()
    b = T.randn(hid)
    G_ = (cin + 63) // 64
    Kq = k * G_ * 64
    wq = T.zeros(hid, k, G_ * 64)
    wq[:, :, :cin] = w.float().permute(0, 2, 1)
    wq = wq.reshape(hid, Kq).bfloat16().cuda()
    x = Xd[:, PAD:PAD + L, :cin].float().permute(0, 2, 1)
    wr, br = w.float().clone().requires_grad_(), b.clone().requires_grad_()
    pre = F.conv1d(x, wr, br, stride=s, padding=p)[:, :, :Lh]
    ref = F.leaky_relu(pre, 0.01)
    Hh = T.zeros(Bn, Lh + 2, hid, device="cuda", dtype=T.bfloat16)
    Xg = Xd.cuda()
    Kn.gemm_nt(Bn * Lh, hid, k * cin, (Xg, (PAD - p) * CT), (Lh, Lp * CT, s * CT, cin, CT), wq, Kq, (Hh, hid),
               (Lh, (Lh + 2) * hid, hid), bias=b.cuda(), act=1, tc=True, a_layout=2)
    assert rel(Hh[:, 1:Lh + 1], ref.permute(0, 2, 1)) < 8e-3
    dH = T.randn(Bn, Lh, hid).bfloat16()
    pre.backward(dH.float().permute(0, 2, 1))
    dHp = T.zeros(Bn, Lh + 2, hid, dtype=T.bfloat16)
    dHp[:, 1:Lh + 1] = dH
    dw = T.zeros(hid, Kq + 1, device="cuda")
    Kn.gemm_tn(Bn * Lh, hid, k * cin, (dHp.cuda(), hid), (Lh, (Lh + 2) * hid, hid), (Xg, (PAD - p) * CT),
               (Lh, Lp * CT, s * CT, cin, CT), dw, Kq + 1, ones_col=True, tc=True, a_layout=2)
    got = dw[:, :Kq].view(hid, k, G_ * 64)[:, :, :cin].permute(0, 2, 1)
    assert rel(got, wr.grad) < 3e-5 and rel(dw[:, Kq], br.grad) < 3e-5
    assert G_ * 64 == cin or float(dw[:, :Kq].view(hid, k, G_ * 64)[:, :, cin:].abs().max()) == 0


@pytest.mark.parametrize("dt", [T.float32, T.bfloat16])
def test_conv1in_dgrad_and_wcolsum(dt):
    """Direct kernels for the two GEMMs with 1-2 output columns: gradient of the raw waveform through the discriminator's first
    conv (k = 7, s = 2, C_in = 1), and the weight / bias gradient of the classifier's Linear(K -> 1)."""
    from audiogan_b200 import kernels as K
    T.manual_seed(4)
    B, L, k, s, Co, p = 3, 61, 7, 2, 16, 3
    To = (L + s - 1) // s
    w = T.randn(Co, k)
    dy = T.randn(B, To, Co).to(dt)
    x = T.zeros(B, 1, L, requires_grad=True)
    F.conv1d(x, w[:, None], None, stride=s, padding=p)[:, :, :To].backward(dy.float().permute(0, 2, 1))
    dx = T.full((B, L + 2 * p), 7.0, device="cuda")
    K.conv1in_dgrad(dy.cuda(), To * Co, w.cuda(), dx, L + 2 * p, k, s, p, Co, B, To, L)
    assert rel(dx[:, p:p + L], x.grad[:, 0]) < 1e-5
    assert float(dx[:, :p].abs().max()) == 0 and float(dx[:, p + L:].abs().max()) == 0
    M, Kn = 1003, 512
    g, X = T.randn(M), T.randn(M, Kn).to(dt)
    out = T.zeros(Kn + 1, device="cuda")
    K.wcolsum(g.cuda(), X.cuda(), M, Kn, out)
    assert rel(out[:Kn], g @ X.float()) < 1e-5 and abs(float(out[Kn]) - float(g.sum())) < 1e-3


@pytest.mark.parametrize("dt", [T.float32, T.bfloat16])
@pytest.mark.parametrize("Bn,Cn,Tn", [(3, 16, 700), (5, 64, 130), (2, 512, 19), (4, 24, 300)])
def test_time_moments_kernels_match_reference_formula(dt, Bn, Cn, Tn):
    """ag_time_moments_fwd / _bwd (calc_dists, audiogan.py:341-348) against the reference formula evaluated by torch in
    float64 on the same (storage-rounded) activation: values and the gradient into the activation."""
    from audiogan_b200 import engine as E
    T.manual_seed(3)
    lens = T.randint(max(1, Tn // 3), Tn + 1, (Bn,))
    lens[0] = Tn
    mask = (T.arange(Tn)[None, :] < lens[:, None]).float()
    buf = (T.randn(Bn, Tn + 6, Cn) * 0.7 + 0.3).to(dt)                 # channel-last storage with pad rows, as the engine's
    buf[:, 3:3 + Tn] *= mask[:, :, None].to(dt)
    bufd = buf.cuda()
    h = bufd[:, 3:3 + Tn].permute(0, 2, 1).requires_grad_(True)         # (B, C, T) view, channel stride 1
    q = E._TimeMomentsFn.apply(h, lens.cuda().to(T.int32))
    w = T.randn(3, Bn, Cn, generator=T.Generator().manual_seed(4))
    (q * w.cuda()).sum().backward()
    hr = buf[:, 3:3 + Tn].permute(0, 2, 1).double().requires_grad_(True)
    lf = lens.double().unsqueeze(1)
    m = hr.sum(2) / lf
    dv = hr - m.unsqueeze(2) * mask.unsqueeze(1).double()
    sr = ((dv ** 2).sum(2) ** 0.5) / lf
    fr = ((dv ** 4).sum(2) ** 0.25) / lf
    (T.stack([m, sr, fr]) * w.double()).sum().backward()
    tol = 2e-5
    for i, ref in enumerate((m, sr, fr)):
        assert rel(q[i], ref.detach().float()) < tol, (i, rel(q[i], ref.detach().float()))
    gref = (hr.grad * mask.unsqueeze(1).double()).float()               # the gradient past a sample's length is masked away
    got = h.grad.float().cpu() * mask.unsqueeze(1)
    assert rel(got, gref) < (2e-5 if dt == T.float32 else 8e-3), rel(got, gref)


def test_step_feed_delivers_every_batch_in_order():
    """feed.StepFeed: pinned staging + copy stream + two device slots; batches arrive intact and in order while the consumer
    keeps the previous slot busy; host metadata (`*_len`) passes through."""
    from audiogan_b200.feed import StepFeed
    gen = T.Generator().manual_seed(0)
    src = [{"real": T.randn(4, 1000, generator=gen), "z": T.randn(4, 5, 100, generator=gen), "real_len": T.full((4,), 1000)}
           for _ in range(7)]
    feed = StepFeed(iter(src), "cuda")
    seen, acc = 0, T.zeros((), device="cuda")
    for i, b in enumerate(feed):
        assert b["real"].is_cuda and not b["real_len"].is_cuda
        acc = acc + (b["real"].double().sum() + b["z"].double().sum()).float()          # consumer work on the current stream
        assert T.equal(b["real"].cpu(), src[i]["real"]) and T.equal(b["z"].cpu(), src[i]["z"])
        seen += 1
    assert seen == 7 and feed.h2d_bytes == 4 * (4 * 1000 + 4 * 5 * 100)
