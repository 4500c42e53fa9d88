"""CPU restatement of the audiogan GAN training step -- TEST INFRASTRUCTURE ONLY.

See oracle/__init__.py: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this file, and only as the checker
or the CPU baseline.  It is a *functional* py3 restatement (parameters live in a dict
keyed by the reference's own state_dict names) of:

    weight_norm                audiogan.py:77-80  (torch.nn.utils.weight_norm, dim=0)
    Generator.forward          audiogan.py:412-468
    dense_res_bottleneck       audiogan.py:266-283
    Discriminator.forward      audiogan.py:514-551
    Residual                   audiogan.py:256-264
    dynamic_rnn                audiogan.py:214-229
    length_mask                audiogan.py:204-211
    BCE-with-logits per sample audiogan.py:187-197
    calc_dists / fourth_moment audiogan.py:336-359
    check_grad / clip_grad     audiogan.py:232-253
    adversarial_movement_d     audiogan.py:139-150
    adversarially_sample_z     audiogan.py:99-137
    D-update loop body         audiogan.py:706-788
    G-update loop body         audiogan.py:816-921
    torch.optim.RMSprop        audiogan.py:693-694 (torch defaults alpha=.99 eps=1e-8)

The arithmetic of the reference lives in un-pinned third-party PyTorch (<=0.3 era,
not in /root/reference); the formulas used here (weight_norm over dim 0, LSTM gate
order i,f,g,o, RMSprop with eps outside the sqrt, unbiased std) are unchanged between
that era and torch 2.x.  **Parity unpinned by the reference's own tests** (it has none):
this file is pinned against the reference classes executed here (oracle/ref_loader.py)
via tests/golden/*.pt and tests/test_oracle.py.

Deliberate, documented deviation: the stop head's ``multinomial`` draw
(audiogan.py:450) is restated as ``stop = u < sigmoid(logit)`` on *supplied*
uniforms ``u`` so that CPU oracle and CUDA path can share one random stream; the
distribution is identical, the RNG stream is not.  Parity runs pin the stop head
(bias -30 -> never stops) as SURVEY 8(c) prescribes.
"""
import math

import torch as T
import torch.nn as NN
import torch.nn.functional as F
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

G_STRUCT = [[17, 8, 128, 16], [9, 4, 64, 32], [9, 4, 64, 32], [9, 4, 32, 32]]   # audiogan.py:368
D_STRUCT = [[7, 2, 16], [7, 2, 32], [7, 2, 64], [7, 2, 128], [7, 2, 256], [7, 2, 512]]  # :476
LRELU = 0.01            # NN.LeakyReLU() / F.leaky_relu defaults (audiogan.py:261, :532)


def div_roundup(x, d):                      # audiogan.py:172-173 (py2 integer division)
    return (x + d - 1) // d


def wn(P, name):
    """w = g * v / ||v|| with the norm over every dim but 0 (audiogan.py:77-80)."""
    g, v = P[name + "_g"], P[name + "_v"]
    if v.dim() == 1:
        norm = v.abs()
    else:
        norm = v.reshape(v.shape[0], -1).norm(2, 1).reshape([-1] + [1] * (v.dim() - 1))
    return v * (g / norm)


def length_mask(size, length):              # audiogan.py:204-211
    ar = T.arange(size[1], device=length.device).unsqueeze(0)
    return (ar < length.reshape(-1, 1)).float()


def bce_with_logits_per_sample(inp, target, weight=None):   # audiogan.py:187-197
    if target.shape != inp.shape:
        raise ValueError("Target size ({}) must be the same as input size ({})".format(
            target.size(), inp.size()))
    max_val = (-inp).clamp(min=0)
    loss = inp - inp * target + max_val + ((-max_val).exp() + (-inp - max_val).exp()).log()
    if weight is not None:
        loss = loss * weight
    return loss.sum(1)


# ----------------------------------------------------------------------------- Generator
def generator_forward(P, c, z=None, batch_size=None, length=None, frame_size=200, noise_size=100,
                      struct=G_STRUCT, u_stop=None, early_exit=True):
    """audiogan.py:412-468.  P: generator state_dict (reference key names).

    Returns (x (B, t*frame), s (B, t) stop logits, stop (B, t) int64, length (B,) samples).
    """
    if z is None:
        nframes = div_roundup(length, frame_size)
        z = T.randn(batch_size, nframes, noise_size, device=c.device)
    else:
        batch_size, nframes, _ = z.shape
    pre = "rnn.0.module."
    w_ih, w_hh = wn(P, pre + "weight_ih"), wn(P, pre + "weight_hh")
    b_ih, b_hh = wn(P, pre + "bias_ih"), wn(P, pre + "bias_hh")
    w_p, b_p = wn(P, "proj.module.weight"), wn(P, "proj.module.bias")
    w_s, b_s = wn(P, "stopper.module.weight"), wn(P, "stopper.module.bias")
    H = w_hh.shape[1]
    zc = T.cat([z, c.unsqueeze(1).expand(batch_size, nframes, c.shape[1])], 2)   # :425-426
    dev = c.device                  # CPU in every parity use; tests/test_stock_torch_gpu.py times the same code on the GPU
    h = T.zeros(batch_size, H, device=dev)
    cc = T.zeros(batch_size, H, device=dev)
    x_t = T.zeros(batch_size, frame_size, device=dev)
    generating = T.ones(batch_size, dtype=T.long, device=dev)
    nlen = T.zeros(batch_size, dtype=T.long, device=dev)
    xs, ss, stops = [], [], []
    for t in range(nframes):                                                    # :437
        inp = T.cat([x_t, zc[:, t]], 1)                                         # :439
        gates = F.linear(inp, w_ih, b_ih) + F.linear(h, w_hh, b_hh)             # LSTMCell :440
        i, f, g, o = gates.chunk(4, 1)
        cc = T.sigmoid(f) * cc + T.sigmoid(i) * T.tanh(g)
        h = T.sigmoid(o) * T.tanh(cc)
        x_t = T.tanh(F.linear(h, w_p, b_p))                                     # :443
        logit = F.linear(h, w_s, b_s)                                           # :444
        if u_stop is None:
            stop_t = T.bernoulli(T.sigmoid(logit.detach())).long()               # :445-450
        else:
            stop_t = (u_stop[:, t:t + 1] < T.sigmoid(logit.detach())).long()
        nlen = nlen + generating                                                # :451
        xs.append(x_t)
        ss.append(logit.squeeze(1))
        stops.append(stop_t)
        generating = generating * (stop_t.squeeze(1) == 0).long()               # :458
        if early_exit and int(generating.sum()) == 0:                           # :459-460
            break
    x = T.cat(xs, 1).unsqueeze(1)                                               # :462, :464
    s = T.stack(ss, 1)                                                          # :463
    infilters = 1
    for li, (k, st, hid, out) in enumerate(struct):                             # :465-467
        pfx = "dense_res_gen.%d.module." % li
        act = F.leaky_relu(F.conv1d(x, wn(P, pfx + "conv.weight"), wn(P, pfx + "conv.bias"),
                                    stride=st, padding=(k - 1) // 2), LRELU)     # :279
        act = F.conv_transpose1d(act, wn(P, pfx + "deconv.weight"), wn(P, pfx + "deconv.bias"),
                                 stride=st, padding=st // 2)                     # :280
        if infilters >= out:
            act = act + x[:, -out:, :]                                          # :281-282
        act = F.leaky_relu(act, LRELU)
        x = T.cat([x, act], 1)
        infilters += out
    pfx = "dense_res_gen.%d.module." % len(struct)
    x_next = F.conv1d(x, wn(P, pfx + "weight"), wn(P, pfx + "bias"), padding=1)  # :403-407
    return x_next.squeeze(1), s, T.cat(stops, 1), nlen * frame_size


# ------------------------------------------------------------------------- Discriminator
_LSTM_CACHE = {}


def _bilstm(P, x_tbc, lengths):
    """dynamic_rnn (audiogan.py:214-229) around NN.LSTM(bidirectional) (:498-503)."""
    in_size, hid = P["rnn.weight_ih_l0"].shape[1], P["rnn.weight_hh_l0"].shape[1]
    key = (in_size, hid)
    if key not in _LSTM_CACHE:
        _LSTM_CACHE[key] = NN.LSTM(in_size, hid, 1, bidirectional=True)
    rnn = _LSTM_CACHE[key]
    names = [n for n, _ in rnn.named_parameters()]
    params = {n: P["rnn." + n] for n in names}
    l_sorted, idx = T.sort(lengths, descending=True)
    _, inv = T.sort(idx)
    packed = pack_padded_sequence(x_tbc[:, idx], l_sorted.cpu())
    out, _ = T.func.functional_call(rnn, params, (packed,))
    out = pad_packed_sequence(out)[0]
    return out[:, inv]


def discriminator_forward(P, x, length, c, cnn_struct=D_STRUCT):
    """audiogan.py:514-551.  Returns (logits (B,T'), cnn_outputs[6], cnn_lengths[6], nframes)."""
    B = x.shape[0]
    act = x.unsqueeze(1)
    nframes = length
    outs, lens = [], []
    for li, (k, st, _) in enumerate(cnn_struct):                                 # :531-536
        pfx = "cnn.%d.module." % li
        act = F.leaky_relu(F.conv1d(act, wn(P, pfx + "weight"), wn(P, pfx + "bias"),
                                    stride=st, padding=(k - 1) // 2), LRELU)
        nframes = (nframes + st - 1) // st
        act = act * length_mask((B, act.shape[2]), nframes).unsqueeze(1)
        outs.append(act)
        lens.append(nframes)
    feat = act.permute(0, 2, 1)                                                  # :538
    cexp = c.unsqueeze(1).expand(B, feat.shape[1], c.shape[1])
    x2 = T.cat([feat, cexp], 2).permute(1, 0, 2)                                 # :541-542
    lstm_out = _bilstm(P, x2, nframes).permute(1, 0, 2)                          # :543-544
    Tm = lstm_out.shape[1]
    hflat = lstm_out.reshape(B * Tm, -1)
    for i in range(2):                                                           # Residual :256-264
        pfx = "residual_net.module.%d.linear." % i
        hflat = F.leaky_relu(F.linear(hflat, wn(P, pfx + "weight"), wn(P, pfx + "bias")) + hflat, LRELU)
    hflat = F.leaky_relu(F.linear(hflat, wn(P, "classifier.module.0.weight"),
                                  wn(P, "classifier.module.0.bias")), LRELU)     # :508-512
    logits = F.linear(hflat, wn(P, "classifier.module.2.weight"), wn(P, "classifier.module.2.bias"))
    return logits.reshape(B, Tm), outs, lens, nframes


# ------------------------------------------------------------------- feature statistics
def fourth_moment(v):                                                            # :336-339
    return (((v - v.mean(0).unsqueeze(0)) ** 4).sum(0)) ** (1 / 4)


def calc_dists(hidden_states, hidden_state_lengths):                             # :341-359
    means_d, stds_d, fourth_d = [], [], []
    for h, l in zip(hidden_states, hidden_state_lengths):
        mask = length_mask((h.shape[0], h.shape[2]), l)
        lf = l.unsqueeze(1).float()
        m = h.sum(2) / lf
        dev = h - m.unsqueeze(2) * mask.unsqueeze(1)
        s = ((dev ** 2).sum(2) ** (1. / 2.)) / lf
        f = ((dev ** 4).sum(2) ** (1. / 4.)) / lf
        for q in (m, s, f):
            means_d.append((q.mean(0), q.std(0)))
            stds_d.append((q.std(0), q.std(0)))
            fourth_d.append((fourth_moment(q), q.std(0)))
    return means_d + stds_d + fourth_d


def feature_penalty(dists_d, dists_g, batch_size):                               # :850-855
    fp = 0
    for r, f in zip(dists_d, dists_g):
        fp = fp + T.pow(r[0] - f[0], 2).mean() / batch_size
    return fp


# --------------------------------------------------------------------- grads / optimizer
def check_grad(grads):                                                           # :232-240
    for g in grads:
        if g is None:
            continue
        assert int((g != g).long().sum()) == 0
        assert int((g.abs() > 1e5).long().sum()) == 0


def clip_grad(grads, clip_norm):                                                 # :243-253
    """Per-tensor clip, in place; returns the SUM of per-tensor norms."""
    if clip_norm == 0:
        return None
    total = 0.
    for g in grads:
        if g is None:
            continue
        n = float(g.norm())
        total += n
        if n > clip_norm:
            g /= (n / clip_norm)
    return total


def rmsprop_step(params, grads, state, lr=1e-4, alpha=0.99, eps=1e-8):           # :693-694
    with T.no_grad():
        for k, p in params.items():
            g = grads.get(k)
            if g is None:
                continue
            sq = state.setdefault(k, T.zeros_like(p))
            sq.mul_(alpha).addcmul_(g, g, value=1 - alpha)
            p.addcdiv_(g, sq.sqrt().add_(eps), value=-lr)


# ----------------------------------------------------------------------- FGSM-style moves
def adversarial_movement_d(Pd, data, data_len, embed_d, target, weight, scale=1e-3):   # :139-150
    cls, _, _, nframes = discriminator_forward(Pd, data, data_len, embed_d)
    loss = bce_with_logits_per_sample(cls, target, weight) / nframes.float()
    grad = T.autograd.grad(loss, data, grad_outputs=T.ones_like(loss))[0]
    return ((grad > 0).float() - (grad < 0).float()) * scale


def adversarially_sample_z(Pg, Pd, z, embed_g, embed_d, noise, g_optim="boundary_seeking",
                           scale=1e-2, u_stop=None):                                   # :99-137
    """z, noise are supplied (the reference draws them at :101, :104)."""
    z = z.detach().clone().requires_grad_(True)
    fake, _, _, fake_len = generator_forward(Pg, embed_g, z=z, u_stop=u_stop)
    fake = fake + noise[:, :fake.shape[1]]
    cls_g, _, _, nframes_g = discriminator_forward(Pd, fake, fake_len, embed_d)
    tgt = T.full_like(cls_g, 0.5 if g_optim == "boundary_seeking" else 0.0)
    weight = length_mask(cls_g.shape, nframes_g)
    loss = bce_with_logits_per_sample(cls_g, tgt, weight) / nframes_g.float()
    grad = T.autograd.grad(loss, z, grad_outputs=T.ones_like(loss))[0]
    advers = ((grad > 1e-9).float() - (grad < -1e-9).float()) * scale
    return (z + advers).detach()


# ------------------------------------------------------------------------ the two updates
def _grads_of(loss, params, retain_graph=False):
    keys = [k for k, p in params.items() if p.requires_grad]
    gs = T.autograd.grad(loss, [params[k] for k in keys], allow_unused=True, retain_graph=retain_graph)
    return {k: (g if g is not None else None) for k, g in zip(keys, gs)}


def d_update(Pg, Pd, st_d, batch, lr=1e-4, clip=1.0, fgsm=False, with_x_grad_norm=False):
    """One discriminator update, audiogan.py:706-788.

    batch keys: real (B,L) waveform, real_len (B,), c_real (B,E) = e_d(cs), c_g, c_d2 (B,E) =
    e_g(cs2), e_d(cs2), z (B,T,noise), noise_real, noise_fake (B,L) already scaled
    (the reference draws them at :724, :750), u_stop optional.
    ``fgsm=False`` is the even-iteration branch (:723-728, :749-751); ``fgsm=True`` the odd
    one (:729-736, :752-759).  Returns a dict of scalars / tensors for comparison.
    """
    for p in Pd.values():
        p.requires_grad_(True)
    real_len = batch["real_len"]
    if not fgsm:
        real = batch["real"] + batch["noise_real"]                               # :724-725
        cls_d, _, _, nframes_d = discriminator_forward(Pd, real, real_len, batch["c_real"])
    else:
        real = batch["real"].clone().requires_grad_(True)                        # :730-731
        cls_d, _, _, nframes_d = discriminator_forward(Pd, real, real_len, batch["c_real"])
    target = T.full_like(cls_d, 0.9)                                             # :727
    weight = length_mask(cls_d.shape, nframes_d)
    if fgsm:
        adversarial_movement_d(Pd, real, real_len, batch["c_real"], target, weight)   # :735 (result unused by the loss)
    loss_d = (bce_with_logits_per_sample(cls_d, target, weight) / nframes_d.float()).mean()   # :739-740
    correct_d = float(((cls_d.detach() > 0).float() * weight).sum())
    num_d = float(weight.sum())

    with T.no_grad():
        fake, _, _, fake_len = generator_forward(Pg, batch["c_g"], z=batch["z"],
                                                 u_stop=batch.get("u_stop"))     # :748
    if not fgsm:
        fake = (fake + batch["noise_fake"][:, :fake.shape[1]]).detach()          # :750-751
    else:
        fake = fake.detach().clone().requires_grad_(True)                        # :753-754
        cls_g, _, _, nframes_g = discriminator_forward(Pd, fake, fake_len, batch["c_d2"])
        tgt0 = T.zeros_like(cls_g)
        w0 = length_mask(cls_g.shape, nframes_g)
        adv = adversarial_movement_d(Pd, fake, fake_len, batch["c_d2"], tgt0, w0)     # :758
        fake = (fake + adv).detach()
    fake.requires_grad_(True)                                                    # :760
    cls_g, _, _, nframes_g = discriminator_forward(Pd, fake, fake_len, batch["c_d2"])  # :761
    weight_g = length_mask(cls_g.shape, nframes_g)
    loss_g_ps = bce_with_logits_per_sample(cls_g, T.zeros_like(cls_g), weight_g) / nframes_g.float()
    out = {}
    if with_x_grad_norm:                                                         # :769-775
        gx = T.autograd.grad(loss_g_ps, fake, grad_outputs=T.ones_like(loss_g_ps), retain_graph=True)[0]
        out["x_grad_norm"] = float(((gx.norm(2, 1) ** 2) / nframes_g.float()).mean())
    loss_g = loss_g_ps.mean()                                                    # :780
    correct_g = float(((cls_g.detach() < 0).float() * weight_g).sum())
    num_g = float(weight_g.sum())
    loss = loss_d + loss_g                                                       # :783
    grads = _grads_of(loss, Pd)                                                  # :784-785
    check_grad(grads.values())
    gn = clip_grad(grads.values(), clip)                                         # :787
    out.update(loss_d=float(loss_d), loss_g=float(loss_g), loss=float(loss), d_grad_norm=gn,
               acc_d=correct_d / num_d, acc_g=correct_g / num_g,
               cls_d=cls_d.detach(), cls_g=cls_g.detach(), fake=fake.detach(), grads=grads)
    rmsprop_step(Pd, grads, st_d, lr=lr)                                         # :788
    return out


def g_update(Pg, Pd, st_g, batch, lr=1e-4, clip=0.1, g_optim="boundary_seeking",
             feature_matching=False, adv_z=False, reinforce=False, baseline=None, lambda_fp=1.0):
    """One generator update, audiogan.py:816-921.

    The core step is feature_matching=adv_z=reinforce=False; the flags add
    the reference's extra passes (:836, :847-855, :873-908).  batch keys: c_g, c_d (B,E),
    z (B,T,noise), noise_fake (B,L); with feature_matching also real, real_len, noise_real;
    with adv_z also noise_adv.
    """
    for p in Pg.values():
        p.requires_grad_(True)
    for p in Pd.values():
        p.requires_grad_(False)
    z = batch["z"]
    u_stop = batch.get("u_stop")
    if adv_z:                                                                    # :836
        z = adversarially_sample_z(Pg, Pd, z, batch["c_g"], batch["c_d"], batch["noise_adv"],
                                   g_optim=g_optim, u_stop=u_stop)
    fake, fake_s, fake_stop, fake_len = generator_forward(Pg, batch["c_g"], z=z, u_stop=u_stop)   # :841
    fake = fake + batch["noise_fake"][:, :fake.shape[1]]                         # :842-843 (noise drawn at fake's size)
    cls_g, hs_g, hl_g, nframes_g = discriminator_forward(Pd, fake, fake_len, batch["c_d"])       # :845
    fp = T.zeros((), device=cls_g.device)
    if feature_matching:                                                         # :847-855
        real = batch["real"] + batch["noise_real"]
        _, hs_d, hl_d, _ = discriminator_forward(Pd, real, batch["real_len"], batch["c_d"])
        fp = feature_penalty(calc_dists(hs_d, hl_d), calc_dists(hs_g, hl_g), fake.shape[0])
    tgt = T.full_like(cls_g, 0.5 if g_optim == "boundary_seeking" else 0.0)      # :857-860
    weight = length_mask(cls_g.shape, nframes_g)
    loss_ps = bce_with_logits_per_sample(cls_g, tgt, weight) / nframes_g.float()  # :864
    _loss = loss_ps.mean()                                                       # :897
    loss = _loss + fp * lambda_fp                                                # :898
    grads = _grads_of(loss, Pg, retain_graph=reinforce)                          # :902-903 (retain_graph=True there too)
    if reinforce:                                                                # :873-908
        # Variable.reinforce() (:901) no longer exists: the stochastic node's backward is restated as the gradient of the
        # score-function surrogate -sum (reward - baseline) * w_r * log p(stop_t).  The reference runs that backward with
        # every generator parameter frozen EXCEPT g.stopper's (:904-907), so only the stop head receives it (on top of what
        # loss.backward() left in .grad): the hidden states are constants here.
        reward = -loss_ps.detach()                                               # :873
        baseline = float(reward.mean()) if baseline is None else baseline * 0.5 + float(reward.mean()) * 0.5   # :874
        nf = fake_len // 200                                                     # :862-863 fake_len / framesize
        w_r = length_mask((fake.shape[0], int(nf.max())), nf)
        adv = (reward - baseline).unsqueeze(1) * w_r                             # :885
        logp = T.where(fake_stop.bool(), F.logsigmoid(fake_s), F.logsigmoid(-fake_s))
        surrogate = -(adv * logp[:, :adv.shape[1]]).sum()
        sk = [k for k in Pg if k.startswith("stopper.")]
        for k, gr in zip(sk, T.autograd.grad(surrogate, [Pg[k] for k in sk], allow_unused=True)):
            if gr is not None:
                grads[k] = gr if grads.get(k) is None else grads[k] + gr
    check_grad(grads.values())                                                   # :909
    gn = clip_grad([g for g in grads.values() if g is not None], clip)           # :910
    out = dict(loss=float(_loss), feature_penalty=float(fp), g_grad_norm=gn, baseline=baseline,
               cls_g=cls_g.detach(), fake=fake.detach(), grads=grads)
    rmsprop_step(Pg, grads, st_g, lr=lr)                                         # :921
    for p in Pd.values():
        p.requires_grad_(True)
    return out


# --------------------------------------------------------------- init + synthetic inputs
def _wn_split(sd, name, w):
    if w.dim() == 1:
        sd[name + "_g"] = w.abs().clone()
    else:
        sd[name + "_g"] = w.reshape(w.shape[0], -1).norm(2, 1).reshape([-1] + [1] * (w.dim() - 1)).clone()
    sd[name + "_v"] = w.clone()


def _uniform(shape, bound, gen):
    return (T.rand(shape, generator=gen) * 2 - 1) * bound


def init_generator(seed, frame_size=200, embed_size=100, noise_size=100, state_size=1024, struct=G_STRUCT):
    """Random parameters with torch's default init *scales* (uniform +-1/sqrt(fan)); key names and
    shapes are the reference's (audiogan.py:362-410).  Not the same RNG stream as NN.Module init --
    parity tests load the *same* dict into both sides, so only names/shapes matter."""
    gen = T.Generator().manual_seed(seed)
    sd = {}
    H = state_size
    b = 1 / math.sqrt(H)
    for n, shp in (("weight_ih", (4 * H, frame_size + embed_size + noise_size)), ("weight_hh", (4 * H, H)),
                   ("bias_hh", (4 * H,)), ("bias_ih", (4 * H,))):
        _wn_split(sd, "rnn.0.module." + n, _uniform(shp, b, gen))
    infilters = 1
    for li, (k, st, hid, out) in enumerate(struct):
        pfx = "dense_res_gen.%d.module." % li
        bc = 1 / math.sqrt(infilters * k)
        _wn_split(sd, pfx + "conv.weight", _uniform((hid, infilters, k), bc, gen))
        _wn_split(sd, pfx + "conv.bias", _uniform((hid,), bc, gen))
        bd = 1 / math.sqrt(out * (k - 1))
        _wn_split(sd, pfx + "deconv.weight", _uniform((hid, out, k - 1), bd, gen))
        _wn_split(sd, pfx + "deconv.bias", _uniform((out,), bd, gen))
        infilters += out
    pfx = "dense_res_gen.%d.module." % len(struct)
    bc = 1 / math.sqrt(infilters * 3)
    _wn_split(sd, pfx + "weight", _uniform((1, infilters, 3), bc, gen))
    _wn_split(sd, pfx + "bias", _uniform((1,), bc, gen))
    _wn_split(sd, "proj.module.weight", _uniform((frame_size, H), b, gen))
    _wn_split(sd, "proj.module.bias", _uniform((frame_size,), b, gen))
    _wn_split(sd, "stopper.module.weight", _uniform((1, H), b, gen))
    _wn_split(sd, "stopper.module.bias", _uniform((1,), b, gen))
    return sd


def init_discriminator(seed, state_size=1024, embed_size=100, cnn_struct=D_STRUCT):
    gen = T.Generator().manual_seed(seed)
    sd = {}
    infilters = 1
    for li, (k, st, out) in enumerate(cnn_struct):
        pfx = "cnn.%d.module." % li
        bc = 1 / math.sqrt(infilters * k)
        _wn_split(sd, pfx + "weight", _uniform((out, infilters, k), bc, gen))
        _wn_split(sd, pfx + "bias", _uniform((out,), bc, gen))
        infilters = out
    Hh = state_size // 2
    b = 1 / math.sqrt(Hh)
    for sfx in ("", "_reverse"):
        sd["rnn.weight_ih_l0" + sfx] = _uniform((4 * Hh, infilters + embed_size), b, gen)
        sd["rnn.weight_hh_l0" + sfx] = _uniform((4 * Hh, Hh), b, gen)
        sd["rnn.bias_ih_l0" + sfx] = _uniform((4 * Hh,), b, gen)
        sd["rnn.bias_hh_l0" + sfx] = _uniform((4 * Hh,), b, gen)
    b = 1 / math.sqrt(state_size)
    for i in range(2):
        pfx = "residual_net.module.%d.linear." % i
        _wn_split(sd, pfx + "weight", _uniform((state_size, state_size), b, gen))
        _wn_split(sd, pfx + "bias", _uniform((state_size,), b, gen))
    _wn_split(sd, "classifier.module.0.weight", _uniform((state_size // 2, state_size), b, gen))
    _wn_split(sd, "classifier.module.0.bias", _uniform((state_size // 2,), b, gen))
    b = 1 / math.sqrt(state_size // 2)
    _wn_split(sd, "classifier.module.2.weight", _uniform((1, state_size // 2), b, gen))
    _wn_split(sd, "classifier.module.2.bias", _uniform((1,), b, gen))
    return sd


def pin_stopper(sd, value=30.0):
    """bias = g*sign(v) = -value -> the stop head never fires."""
    sd["stopper.module.bias_g"] = T.full((1,), float(value))
    sd["stopper.module.bias_v"] = T.full((1,), -1.0)
    return sd
