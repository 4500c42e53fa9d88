"""The two updates of the GAN training step (audiogan.py:706-788 and :816-921), the fused
multi-tensor optimizer with the reference's per-tensor clip, and the loop helpers
``check_grad`` / ``clip_grad`` (audiogan.py:232-253).

``d_update`` / ``g_update`` take the same batch dictionaries as ``oracle.restated`` so the parity
tests read side by side; they run the reference's call pattern on the drop-in modules
(``Generator`` / ``Discriminator`` forward, ``loss.backward()``, clip, optimizer step).
"""
import collections
import os

import torch
from torch.autograd.function import once_differentiable

from . import kernels as K
from . import _abi as A
from . import engine as E
from .plan import no_weight_grads
from .modules import (binary_cross_entropy_with_logits_per_sample, calc_dists, length_mask, cat_lengths, dev_i32)  # noqa: F401

# elements per block of the multi-tensor kernels.  8 k: ~1000 blocks for the default nets (64 k left 1.1 blocks per SM and the
# optimizer pass at 1.8 TB/s); AUDIOGAN_MT_CHUNK is the A/B knob
_CHUNK = int(os.environ.get("AUDIOGAN_MT_CHUNK", "8192"))
_NSTAGE = 4


class _MT:
    """Chunk tables for the multi-tensor kernels over a fixed parameter list."""

    def __init__(self, params):
        self.params = [p for p in params]
        dev = self.params[0].device
        ct, co = [], []
        for i, p in enumerate(self.params):
            n = p.numel()
            for off in range(0, n, _CHUNK):
                ct.append(i)
                co.append(off)
        self.nchunks = len(ct)
        self.chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=dev)
        self.chunk_off = torch.tensor(co, dtype=torch.int64, device=dev)
        self.sqnorm = torch.zeros(len(self.params), device=dev)
        self.flags = torch.zeros(2, dtype=torch.int32, device=dev)
        self.device = dev
        self._stage, self._stage_ev, self._stage_i = None, None, 0

    def table(self, state1=None, state2=None):
        ents, key = [], [id(state1), id(state2)]
        for i, p in enumerate(self.params):
            g = p.grad if p.grad is not None else None
            if g is not None and not g.is_contiguous():
                p.grad = g = g.contiguous()
            ents.append((p.data, g, state1[i] if state1 else None, state2[i] if state2 else None))
            key.append(p.data_ptr())
            key.append(g.data_ptr() if g is not None else 0)
        # in steady state the allocator hands the gradients the same addresses every step: reuse the device table
        key = tuple(key)
        if key == getattr(self, "_tab_key", None):
            return self._tab
        # parameters without a gradient are skipped by giving them n = 0
        arr = (A.MtEntry * len(ents))()
        for i, (p, g, s1, s2) in enumerate(ents):
            arr[i].p, arr[i].g, arr[i].s1, arr[i].s2 = K.addr(p), K.addr(g), K.addr(s1), K.addr(s2)
            arr[i].n = p.numel() if g is not None else 0
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        # Pinned staging buffers owned by this object (no allocation on the step path): a ring of _NSTAGE for eager steps (a
        # buffer is rewritten only after the upload issued _NSTAGE table changes ago has left the host) plus one for a table
        # built while a CUDA graph is being captured (graph.GraphedStep: the gradients then live at the graph pool's addresses;
        # the captured copy node re-reads that buffer on every replay, so eager steps in between must not overwrite it).
        cap = torch.cuda.is_current_stream_capturing()
        if self._stage is None or self._stage[0].numel() != raw.numel():
            if cap:
                raise RuntimeError("FusedRMSprop: run one eager step before capturing a CUDA graph (staging buffers)")
            self._stage = [torch.empty(raw.numel(), dtype=torch.uint8).pin_memory() for _ in range(_NSTAGE + 1)]
            self._stage_ev, self._stage_i = [None] * _NSTAGE, 0
        if cap:
            stage = self._stage[_NSTAGE]
            stage.copy_(raw)
            self._tab, self._tab_key = stage.to(self.device, non_blocking=True), key
            return self._tab
        i = self._stage_i = (self._stage_i + 1) % _NSTAGE
        if self._stage_ev[i] is not None:
            self._stage_ev[i].synchronize()
        stage = self._stage[i]
        stage.copy_(raw)
        self._tab, self._tab_key = stage.to(self.device, non_blocking=True), key
        self._stage_ev[i] = torch.cuda.Event()
        self._stage_ev[i].record()
        return self._tab

    def sqnorms(self, table, grad_scale=1.0):
        """per-tensor sum of squares + NaN / |g * grad_scale| > 1e5 flags (check_grad's test on the gradient the optimizer
        will see: after a summed all-reduce the stored values are world x the mean gradient)."""
        self.sqnorm.zero_()
        self.flags.zero_()
        A.call("ag_mt_sqnorm", K.addr(table), K.addr(self.chunk_tensor), K.addr(self.chunk_off), self.nchunks, _CHUNK,
               K.addr(self.sqnorm), K.addr(self.flags), 1e5 / float(grad_scale), A.stream())


# chunk tables for check_grad / clip_grad on caller-supplied parameter lists: a small LRU keyed by the parameters' identity
# (weak references would not do: torch Parameters hash by identity but a list of them is not weak-referenceable)
_MT_CACHE_MAX = 8
_mt_cache = collections.OrderedDict()


def _mt_for(params):
    params = list(params)
    key = tuple(id(p) for p in params)
    mt = _mt_cache.get(key)
    if mt is not None and (mt.device != params[0].device or any(a is not b for a, b in zip(mt.params, params))):
        mt = None                               # ids were recycled by other tensors
    if mt is None:
        mt = _mt_cache[key] = _MT(params)
    _mt_cache.move_to_end(key)
    while len(_mt_cache) > _MT_CACHE_MAX:
        _mt_cache.popitem(last=False)
    return mt


def check_grad(params):                                          # audiogan.py:232-240
    """assert no NaN and no |g| > 1e5 in any gradient (one fused pass + one flag read)."""
    mt = _mt_for(params)
    mt.sqnorms(mt.table())
    nan, big = mt.flags.tolist()
    assert nan == 0
    assert big == 0


def clip_grad(params, clip_norm):                                # audiogan.py:243-253
    """Per-tensor clip in place; returns the SUM of the per-tensor norms (as the reference does)."""
    if clip_norm == 0:
        return
    mt = _mt_for(params)
    tab = mt.table()
    mt.sqnorms(tab)
    A.call("ag_mt_clip", K.addr(tab), K.addr(mt.chunk_tensor), K.addr(mt.chunk_off), mt.nchunks, _CHUNK,
           K.addr(mt.sqnorm), float(clip_norm), A.stream())
    return mt.sqnorm.sqrt().sum()


class FusedRMSprop:
    """torch.optim.RMSprop(lr, alpha=0.99, eps=1e-8) (audiogan.py:693-694) as two multi-tensor launches.

    ``step(clip=c)`` fuses the reference's ``clip_grad(params, c)`` (per-tensor) into the update;
    ``step()`` after a separate ``clip_grad`` call is the literal reference sequence.  ``grad_scale``
    multiplies every gradient first (1/world_size after a summed all-reduce)."""

    def __init__(self, params, lr=1e-4, alpha=0.99, eps=1e-8, adam=False, betas=(0.9, 0.999)):
        self.params = list(params)
        self.lr, self.alpha, self.eps = lr, alpha, eps
        self.adam, self.betas, self.nstep = adam, betas, 0
        self.mt = _MT(self.params)
        self.s1 = [torch.zeros_like(p.data) for p in self.params]
        self.s2 = [torch.zeros_like(p.data) for p in self.params] if adam else None
        self.param_groups = [{"params": self.params, "lr": lr}]
        self.last_norm = None

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def step(self, clip=0.0, grad_scale=1.0, check=False):
        mt = self.mt
        tab = mt.table(self.s1, self.s2)
        sq = None
        if clip > 0 or check:
            mt.sqnorms(tab, grad_scale)
            sq = mt.sqnorm
            self.last_norm = sq.sqrt().sum() * grad_scale
        if check:
            nan, big = mt.flags.tolist()
            assert nan == 0 and big == 0, "check_grad: NaN or |g| > 1e5 (audiogan.py:239-240)"
        self.nstep += 1
        if self.adam:
            A.call("ag_mt_adam", K.addr(tab), K.addr(mt.chunk_tensor), K.addr(mt.chunk_off), mt.nchunks, _CHUNK,
                   K.addr(sq), float(clip), float(grad_scale), float(self.lr), float(self.betas[0]),
                   float(self.betas[1]), float(self.eps), self.nstep, A.stream())
        else:
            A.call("ag_mt_rmsprop", K.addr(tab), K.addr(mt.chunk_tensor), K.addr(mt.chunk_off), mt.nchunks, _CHUNK,
                   K.addr(sq), float(clip), float(grad_scale), float(self.lr), float(self.alpha), float(self.eps),
                   A.stream())
        for p in self.params:          # the kernels wrote the parameters through raw pointers: tell the packed-operand cache
            p._ag_epoch = getattr(p, "_ag_epoch", 0) + 1
        return self.last_norm


# ------------------------------------------------------------------------------- fused loss
class _BCEConstFn(torch.autograd.Function):
    """mean_b( sum_t mask * BCE(x, target) / nframes[b] ) with a constant target (audiogan.py:739-740, :766, :780,
    :864, :897) plus the accuracy counters (:741-742, :781-782), one launch; backward is a scale."""

    @staticmethod
    def forward(ctx, x, nframes_i32, target, sign):
        B, T = x.shape
        x = x.contiguous()
        loss = torch.zeros((), device=x.device)
        stats = torch.zeros(2, device=x.device)                # correct, num
        loss_ps = torch.empty(B, device=x.device)
        dlog = torch.empty_like(x)
        K.bce_const_fused(x, T, nframes_i32, target, sign, loss, loss_ps, dlog, stats, B, T)
        ctx.save_for_backward(dlog)
        ctx.mark_non_differentiable(loss_ps, stats)
        return loss, loss_ps, stats

    @staticmethod
    @once_differentiable
    def backward(ctx, gloss, _gps, _gstats):
        (dlog,) = ctx.saved_tensors
        return dlog * gloss, None, None, None


class _BCEPairFn(torch.autograd.Function):
    """The two losses of the batched D-update over ONE [2B, T] logits tensor (rows [0, Bn): target / sign 0, rows [Bn, 2B):
    target / sign 1): the same two kernel launches as two _BCEConstFn calls, but one zero-initialised accumulator, one dlogits
    buffer and a backward that scales its two halves in place of autograd's slice / zero-fill / add nodes (~9 small launches
    fewer on the critical path between the discriminator's forward and backward)."""

    @staticmethod
    def forward(ctx, x, nframes_i32, Bn, t0, s0, t1, s1):
        B2, T = x.shape
        x = x.contiguous()
        acc = torch.zeros(6, device=x.device)                 # loss_0, loss_1, correct_0, num_0, correct_1, num_1
        loss_ps = torch.empty(B2, device=x.device)
        dlog = torch.empty_like(x)
        K.bce_const_fused(x, T, nframes_i32, t0, s0, acc[0:1], loss_ps, dlog, acc[2:4], Bn, T)
        K.bce_const_fused(x[Bn:], T, nframes_i32[Bn:], t1, s1, acc[1:2], loss_ps[Bn:], dlog[Bn:], acc[4:6], B2 - Bn, T)
        ctx.save_for_backward(dlog)
        ctx.Bn = Bn
        l0, l1, st0, st1 = acc[0], acc[1], acc[2:4], acc[4:6]
        ctx.mark_non_differentiable(loss_ps, st0, st1)
        return l0, l1, loss_ps, st0, st1

    @staticmethod
    @once_differentiable
    def backward(ctx, g0, g1, _gps, _gs0, _gs1):
        (dlog,) = ctx.saved_tensors
        Bn = ctx.Bn
        out = torch.empty_like(dlog)
        if g0 is None:
            out[:Bn].zero_()
        else:
            torch.mul(dlog[:Bn], g0, out=out[:Bn])
        if g1 is None:
            out[Bn:].zero_()
        else:
            torch.mul(dlog[Bn:], g1, out=out[Bn:])
        return out, None, None, None, None, None, None


def masked_bce_pair(logits, nframes, Bn, target0, sign0, target1, sign1):
    """-> (loss_0, loss_1, per-sample losses (2B,), stats_0, stats_1) for rows [0, Bn) / [Bn, 2B) of `logits`"""
    return _BCEPairFn.apply(logits, dev_i32(nframes, logits.device), int(Bn), float(target0), float(sign0), float(target1),
                            float(sign1))


def masked_bce_mean(logits, nframes, target, sign=1.0):
    """-> (loss scalar, per-sample loss (B,), stats = [sum(mask * (sign*x > 0)), sum(mask)])."""
    return _BCEConstFn.apply(logits, dev_i32(nframes, logits.device), float(target), float(sign))


# ------------------------------------------------------------------------------- FGSM helpers
def adversarial_movement_d(d, data, data_len, embed_d, target, weight, scale=1e-3):     # audiogan.py:139-150
    cls, _, _, nframes = d(data, data_len, embed_d)
    loss = binary_cross_entropy_with_logits_per_sample(cls, target, weight) / dev_i32(nframes, cls.device).float()
    with no_weight_grads():           # autograd.grad never touches p.grad in the reference: data gradient only
        grad = torch.autograd.grad(loss, data, grad_outputs=torch.ones_like(loss))[0]
    return ((grad > 0).float() - (grad < 0).float()) * scale


def adversarially_sample_z(g, d, z, embed_g, embed_d, noise, g_optim="boundary_seeking", scale=1e-2, u_stop=None):
    """audiogan.py:99-137 with z / noise supplied by the caller (the reference draws them at :101, :104)."""
    z = z.detach().clone().requires_grad_(True)
    fake, _, _, fake_len = g(z=z, c=embed_g, u_stop=u_stop)
    fake = fake + noise[:, :fake.shape[1]]
    cls_g, _, _, nframes_g = d(fake, fake_len, embed_d)
    tgt = torch.full_like(cls_g, 0.5 if g_optim == "boundary_seeking" else 0.0)
    weight = length_mask(cls_g.shape, nframes_g)
    loss = binary_cross_entropy_with_logits_per_sample(cls_g, tgt, weight) / dev_i32(nframes_g, cls_g.device).float()
    with no_weight_grads():
        grad = torch.autograd.grad(loss, z, grad_outputs=torch.ones_like(loss))[0]
    advers = ((grad > 1e-9).float() - (grad < -1e-9).float()) * scale
    return (z + advers).detach()


# ------------------------------------------------------------------------------- the two updates
def _set_requires_grad(module, flag):
    ps = module._params() if hasattr(module, "_params") else module.parameters()
    for p in ps:
        p.requires_grad = flag


def _d_update_batched(g, d, opt_d, batch, clip, check, grad_sync, fake_pass=None):
    """Even-iteration D-update (audiogan.py:723-728, :748-751, :761-788) with the real and the fake pass of the
    discriminator run as ONE pass over the concatenated 2B minibatch.  D has no cross-sample operation, so logits,
    losses and gradients are those of the two separate calls; the sequential recurrent kernels run once, not twice.
    ``fake_pass``: (fake, fake_len) of a generator pass already run on batch["z"], batch["c_g"] (core_step)."""
    real_len = batch["real_len"]
    if fake_pass is None:
        with torch.no_grad():
            fake, _, _, fake_len = g(z=batch["z"], c=batch["c_g"], u_stop=batch.get("u_stop"))   # :748
    else:
        fake, fake_len = fake_pass
    fake = fake + batch["noise_fake"][:, :fake.shape[1]]                         # :750-751
    real = batch["real"] + batch["noise_real"]                                   # :724-725
    Bn, Lr, Lf = real.shape[0], real.shape[1], fake.shape[1]
    if Lr != Lf:                                                                # early stop: zero-extend (== conv zero padding)
        L = max(Lr, Lf)
        real = torch.nn.functional.pad(real, (0, L - Lr))
        fake = torch.nn.functional.pad(fake, (0, L - Lf))
    x = torch.cat([real, fake], 0)
    lens = cat_lengths([real_len, fake_len], x.device)
    c = torch.cat([batch["c_real"], batch["c_d2"]], 0)
    cls, _, _, nf = d(x, lens, c)
    cls_d, cls_g = cls[:Bn], cls[Bn:]
    loss_d, loss_g, _, st_d, st_g = masked_bce_pair(cls, nf, Bn, 0.9, 1.0, 0.0, -1.0)     # :739-742, :762-766, :780-782
    loss = loss_d + loss_g                                                      # :783
    opt_d.zero_grad()
    loss.backward()                                                             # :784-785
    scale = grad_sync(opt_d.params) if grad_sync is not None else 1.0
    gn = opt_d.step(clip=clip, grad_scale=scale, check=check)                   # :786-788 fused
    return dict(loss_d=loss_d.detach(), loss_g=loss_g.detach(), loss=loss.detach(), d_grad_norm=gn, stats_d=st_d,
                stats_g=st_g, cls_d=cls_d.detach(), cls_g=cls_g.detach(), fake=fake.detach())


def d_update(g, d, opt_d, batch, clip=1.0, fgsm=False, with_x_grad_norm=False, check=False, grad_sync=None,
             batched=True):
    """One discriminator update, audiogan.py:706-788.  batch keys as oracle.restated.d_update.
    ``grad_sync(params)`` (optional) is called between backward and the optimizer step (data-parallel
    all-reduce); it returns the gradient scale to apply.  ``batched`` runs the two D passes of the even-iteration
    branch as one 2B pass (same numbers); ``batched=False`` is the literal call-by-call sequence."""
    _set_requires_grad(g, False)                                                # :706-709
    _set_requires_grad(d, True)
    if batched and not fgsm and not with_x_grad_norm:
        return _d_update_batched(g, d, opt_d, batch, clip, check, grad_sync)
    real_len = batch["real_len"]
    u_stop = batch.get("u_stop")
    if not fgsm:
        real = batch["real"] + batch["noise_real"]                              # :724-725
        cls_d, _, _, nframes_d = d(real, real_len, batch["c_real"])
    else:
        real = batch["real"].clone().requires_grad_(True)                       # :730-731
        cls_d, _, _, nframes_d = d(real, real_len, batch["c_real"])
        target = torch.full_like(cls_d, 0.9)
        weight = length_mask(cls_d.shape, nframes_d)
        adversarial_movement_d(d, real, real_len, batch["c_real"], target, weight)      # :735 (unused by the loss)
    loss_d, _, st_d = masked_bce_mean(cls_d, nframes_d, 0.9, 1.0)               # :739-742
    with torch.no_grad():
        fake, _, _, fake_len = g(z=batch["z"], c=batch["c_g"], u_stop=u_stop)   # :748
    if not fgsm:
        fake = (fake + batch["noise_fake"][:, :fake.shape[1]]).detach()         # :750-751
    else:
        fake = fake.detach().clone().requires_grad_(True)                       # :753-754
        cls_g, _, _, nframes_g = d(fake, fake_len, batch["c_d2"])
        adv = adversarial_movement_d(d, fake, fake_len, batch["c_d2"], torch.zeros_like(cls_g),
                                     length_mask(cls_g.shape, nframes_g))       # :758
        fake = (fake + adv).detach()
    out = {}
    if with_x_grad_norm:
        fake.requires_grad_(True)                                               # :760
    cls_g, _, _, nframes_g = d(fake, fake_len, batch["c_d2"])                    # :761
    loss_g, loss_g_ps, st_g = masked_bce_mean(cls_g, nframes_g, 0.0, -1.0)       # :762-766, :780-782
    if with_x_grad_norm:                                                        # :769-775
        with no_weight_grads():
            gx = torch.autograd.grad(loss_g, fake, retain_graph=True)[0] * fake.shape[0]
        out["x_grad_norm"] = ((gx.norm(2, 1) ** 2) / dev_i32(nframes_g, gx.device).float()).mean()
    loss = loss_d + loss_g                                                      # :783
    opt_d.zero_grad()
    loss.backward()                                                             # :784-785
    scale = grad_sync(opt_d.params) if grad_sync is not None else 1.0
    gn = opt_d.step(clip=clip, grad_scale=scale, check=check)                   # :786-788 fused
    out.update(loss_d=loss_d.detach(), loss_g=loss_g.detach(), loss=loss.detach(), d_grad_norm=gn,
               stats_d=st_d, stats_g=st_g, cls_d=cls_d.detach(), cls_g=cls_g.detach(), fake=fake.detach())
    return out


def g_update(g, d, opt_g, batch, clip=0.1, g_optim="boundary_seeking", feature_matching=False, adv_z=False,
             check=False, lambda_fp=1.0, grad_sync=None, reinforce=False, baseline=None, fake_pass=None, gather=None):
    """One generator update, audiogan.py:816-921 (core step: feature_matching = adv_z = reinforce = False).
    batch keys as oracle.restated.g_update.

    ``reinforce=True`` adds the REINFORCE update of the stop head (:873-908): reward = -loss per sample, ``baseline`` the
    running EMA(0.5) of the mean reward (None on the first call; a 1-element device tensor afterwards -- the returned
    ``baseline`` is passed back in on the next call, no host read), score-function gradient into ``g.stopper`` only."""
    _set_requires_grad(g, True)                                                 # :816-819
    _set_requires_grad(d, False)
    z = batch["z"]
    u_stop = batch.get("u_stop")
    if adv_z:                                                                   # :836
        z = adversarially_sample_z(g, d, z, batch["c_g"], batch["c_d"], batch["noise_adv"], g_optim, u_stop=u_stop)
    if fake_pass is None:
        fake, fake_s, fake_stop, fake_len = g(z=z, c=batch["c_g"], u_stop=u_stop)   # :841
    else:                                     # core_step: this pass already ran (same parameters) beside the D-update's
        fake, fake_s, fake_stop, fake_len = fake_pass
    fake = fake + batch["noise_fake"][:, :fake.shape[1]]                        # :842-843
    cls_g, hs_g, hl_g, nframes_g = d(fake, fake_len, batch["c_d"])               # :845
    fp = None
    if feature_matching:                                                        # :847-855
        real = batch["real"] + batch["noise_real"]
        _, hs_d, hl_d, _ = d(real, batch["real_len"], batch["c_d"])
        fp = 0
        # data-parallel: `gather` (dist.gather_batch) makes the batch moments those of the global minibatch (:350-358)
        nb = fake.shape[0] * (torch.distributed.get_world_size() if gather is not None and torch.distributed.is_initialized() else 1)
        for r, f in zip(calc_dists(hs_d, hl_d, gather), calc_dists(hs_g, hl_g, gather)):
            fp = fp + torch.pow(r[0] - f[0], 2).mean() / nb
    tgt = 0.5 if g_optim == "boundary_seeking" else 0.0                          # :857-860
    _loss, loss_ps, _ = masked_bce_mean(cls_g, nframes_g, tgt, -1.0)             # :864, :897
    loss = _loss if fp is None else _loss + fp * lambda_fp                      # :898
    new_baseline = baseline
    if reinforce:                                                               # :873-885, :900-901
        Bn, Tn = fake_s.shape
        if baseline is not None and not torch.is_tensor(baseline):
            baseline = torch.full((1,), float(baseline), device=fake_s.device)
        new_baseline = torch.empty(1, device=fake_s.device)
        ds = torch.empty(Bn, Tn, device=fake_s.device)
        K.reinforce_dlogit(fake_s.detach(), fake_s.stride(0), fake_s._ag_stop, fake_s._ag_stop.stride(0), loss_ps,
                           fake_s._ag_glen, baseline, new_baseline, ds, Tn, Bn, Tn)
        g._get_plan().stopper_ds = ds            # consumed by the generator's backward (stop head only, :904-908)
    opt_g.zero_grad()
    loss.backward()                                                             # :902-903
    scale = grad_sync(opt_g.params) if grad_sync is not None else 1.0
    gn = opt_g.step(clip=clip, grad_scale=scale, check=check)                   # :909-921 fused
    _set_requires_grad(d, True)
    return dict(loss=_loss.detach(), feature_penalty=fp, g_grad_norm=gn, cls_g=cls_g.detach(), fake=fake.detach(),
                loss_ps=loss_ps, baseline=new_baseline, fake_len=fake_len)


def core_step(g, d, opt_d, opt_g, batch, clip_d=1.0, clip_g=0.1, g_optim="boundary_seeking", check=False, grad_sync=None):
    """One core training step (d_update + g_update, even-iteration branch, no extras) with the two generator
    forward passes of the step run as ONE pass over 2B samples.

    The D-update's detached generator pass (audiogan.py:748, noise batch["z"]) and the G-update's generator pass (:841, noise
    batch["g_z"]) use the same generator parameters -- the D-update only changes D -- so running them as one batched forward
    computes exactly what the two calls compute while the sequential recurrent kernel (80 frames, latency-bound) runs once
    instead of twice.  Only the second half of that pass is differentiated (Generator.forward(grad_from=B)).  Everything else
    is the literal sequence: D-update on [real | fake_1], D's optimizer step, then the G-update's discriminator pass on fake_2
    with the UPDATED discriminator, backward through it and the generator, G's optimizer step.
    Returns (d_update's dict, g_update's dict)."""
    _set_requires_grad(g, True)
    Bn = batch["z"].shape[0]
    z_all = torch.cat([batch["z"], batch["g_z"]], 0)
    c_all = torch.cat([batch["c_g"], batch["g_c_g"]], 0)
    fake_all, s_all, _, len_all = g(z=z_all, c=c_all, u_stop=None, grad_from=Bn, defer_tail=True)
    host = getattr(len_all, "_ag_host", None)

    def half(a, b):
        ln = len_all[a:b]
        if host is not None:
            ln._ag_host = host[a:b]
        return ln

    _set_requires_grad(g, False)                                                # audiogan.py:706-709
    _set_requires_grad(d, True)
    m1 = _d_update_batched(g, d, opt_d, batch, clip_d, check, grad_sync, fake_pass=(fake_all[:Bn].detach(), half(0, Bn)))
    E.shadow_join(fake_all.device)             # the second half's conv stack ran under the D-update's recurrent kernels
    gb = {"c_g": batch["g_c_g"], "c_d": batch["g_c_d"], "z": batch["g_z"], "noise_fake": batch["g_noise_fake"], "u_stop": None}
    m2 = g_update(g, d, opt_g, gb, clip=clip_g, g_optim=g_optim, check=check, grad_sync=grad_sync,
                  fake_pass=(fake_all[Bn:], s_all[Bn:], None, half(Bn, 2 * Bn)))
    return m1, m2
