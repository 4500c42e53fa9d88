"""Drop-in ``Generator`` / ``Discriminator`` and helper functions with the reference's API.

Same constructor keywords, ``forward`` signatures, return tuples and ``state_dict`` key names
(weight-norm ``_g`` / ``_v`` split *and* the ``DataParallel`` ``.module.`` infix) as
``/root/reference/audiogan.py:361-551``; all arithmetic runs in libaudiogan_b200.so.  Parameters
are fp32 master copies owned by PyTorch; gradients arrive through autograd as usual, so the
reference training loop (``loss.backward()``, ``T.autograd.grad(loss, x)``, ``requires_grad``
toggling, ``opt.step()``) works unchanged.  CUDA only: there is no CPU fallback.
"""
import collections
import math

import torch
import torch.nn as NN

from . import engine as E
from . import kernels as K
from . import plan as P

G_STRUCT = [[17, 8, 128, 16], [9, 4, 64, 32], [9, 4, 64, 32], [9, 4, 32, 32]]        # audiogan.py:368
D_STRUCT = [[7, 2, 16], [7, 2, 32], [7, 2, 64], [7, 2, 128], [7, 2, 256], [7, 2, 512]]  # audiogan.py:476


def div_roundup(x, d):                       # audiogan.py:172-173 (py2 integer division)
    return (x + d - 1) // d


# --------------------------------------------------------------------------- parameter holders
class _DP(NN.Module):
    """Stands in for NN.DataParallel in the module tree: contributes the ``.module.`` key infix only.
    Data parallelism is one process per GPU with NCCL all-reduce (audiogan_b200.dist)."""

    def __init__(self, module):
        NN.Module.__init__(self)
        self.module = module


def _uniform(shape, bound):
    return (torch.rand(shape) * 2 - 1) * bound


class _WN(NN.Module):
    """Holder of weight-normed tensors named like torch.nn.utils.weight_norm does (audiogan.py:77-80):
    ``<name>_g`` (norm over every dim but 0) and ``<name>_v``; registration order = the reference's."""

    def __init__(self, tensors):
        NN.Module.__init__(self)
        for name, w in tensors:
            if w.dim() == 1:
                g = w.abs().clone()
            else:
                g = w.reshape(w.shape[0], -1).norm(2, 1).reshape([-1] + [1] * (w.dim() - 1)).clone()
            self.register_parameter(name + "_g", NN.Parameter(g))
            self.register_parameter(name + "_v", NN.Parameter(w.clone()))


def _linear_holder(out_f, in_f):
    b = 1.0 / math.sqrt(in_f)
    return _WN([("weight", _uniform((out_f, in_f), b)), ("bias", _uniform((out_f,), b))])


def _conv_holder(cout, cin, k, transposed=False):
    # torch default init scale: U(+-1/sqrt(fan_in)); ConvTranspose1d weight is [cin, cout, k], fan_in = cout*k
    shape = (cin, cout, k) if transposed else (cout, cin, k)
    b = 1.0 / math.sqrt(shape[1] * k)
    return _WN([("weight", _uniform(shape, b)), ("bias", _uniform((cout,), b))])


class _Bottleneck(NN.Module):
    """Parameter holder for dense_res_bottleneck (audiogan.py:266-283)."""

    def __init__(self, kernel, stride, infilters, hidden_filters, outfilters):
        NN.Module.__init__(self)
        self.infilters, self.outfilters = infilters, outfilters
        self.conv = _conv_holder(hidden_filters, infilters, kernel)
        self.deconv = _conv_holder(outfilters, hidden_filters, kernel - 1, transposed=True)


class _Residual(NN.Module):
    def __init__(self, size):
        NN.Module.__init__(self)
        self.size = size
        self.linear = _linear_holder(size, size)


class _PlainLSTM(NN.Module):
    """Parameter holder named like NN.LSTM(bidirectional=True, num_layers=1) (audiogan.py:498-503)."""

    def __init__(self, in_size, hid):
        NN.Module.__init__(self)
        b = 1.0 / math.sqrt(hid)
        for sfx in ("", "_reverse"):
            self.register_parameter("weight_ih_l0" + sfx, NN.Parameter(_uniform((4 * hid, in_size), b)))
            self.register_parameter("weight_hh_l0" + sfx, NN.Parameter(_uniform((4 * hid, hid), b)))
            self.register_parameter("bias_ih_l0" + sfx, NN.Parameter(_uniform((4 * hid,), b)))
            self.register_parameter("bias_hh_l0" + sfx, NN.Parameter(_uniform((4 * hid,), b)))


class _PlanOwner(NN.Module):
    """Builds the flat parameter plan lazily and rebuilds it when parameter storage moves (.cuda(), .to())."""

    _plan = None
    _mode = "fp32"

    def set_mode(self, mode):
        """"fp32": FFMA kernels, <=1e-5 parity with the reference's fp32 path.  "bf16": tcgen05 tensor-core GEMMs on
        bf16 operands with fp32 accumulation (<=2e-2); parameters, optimizer state and recurrent state stay fp32."""
        if mode not in ("fp32", "bf16"):
            raise ValueError(mode)
        object.__setattr__(self, "_mode", mode)
        if self._plan is not None:
            self._plan.mode = mode
        return self

    def __getstate__(self):
        # T.save(module) / copy.deepcopy (audiogan.py:936-939 pickles whole modules): the plan holds device pointers
        # and is rebuilt lazily, so it never travels with the module
        st = dict(self.__dict__)
        st.pop("_plan", None)
        st.pop("_param_list", None)
        return st

    def invalidate_packed(self):
        """Call after changing parameters behind torch's back (writes through ``p.data`` do not bump the version counters the
        packed-operand cache watches; in-place ops on the parameter itself, load_state_dict and the fused optimizers do)."""
        for p in self.parameters():
            p._ag_epoch = getattr(p, "_ag_epoch", 0) + 1

    def _params(self):
        """Cached parameter list (module.parameters() walks the module tree: ~0.1 ms per call, several calls per step)."""
        pl = self.__dict__.get("_param_list")
        if pl is None:
            pl = list(self.parameters())
            object.__setattr__(self, "_param_list", pl)
        return pl

    def _get_plan(self):
        params = self._params()
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("audiogan_b200 runs on CUDA (sm_100a) only -- move the module with .cuda(); "
                               "there is no CPU fallback")
        pl = self._plan
        if pl is None or pl.device != dev or pl.signature != tuple(p.data_ptr() for p in params):
            with torch.no_grad():
                pl = self._build_plan(dev)
            object.__setattr__(self, "_plan", pl)
        pl.mode = self._mode
        return pl


# =========================================================================================
class Generator(_PlanOwner):
    """audiogan.py:361-468."""

    def __init__(self, frame_size=200, embed_size=200, noise_size=100, state_size=1024, num_layers=1,
                 struct=G_STRUCT):
        NN.Module.__init__(self)
        if num_layers != 1:
            raise NotImplementedError("num_layers > 1 (the reference default and every BASELINE config use 1)")
        self._frame_size, self._noise_size = frame_size, noise_size
        self._state_size, self._embed_size, self._num_layers = state_size, embed_size, num_layers
        self._struct = [list(s) for s in struct]
        H, nin = state_size, frame_size + embed_size + noise_size
        b = 1.0 / math.sqrt(H)
        self.rnn = NN.ModuleList([_DP(_WN([("weight_ih", _uniform((4 * H, nin), b)), ("weight_hh", _uniform((4 * H, H), b)),
                                           ("bias_hh", _uniform((4 * H,), b)), ("bias_ih", _uniform((4 * H,), b))]))])
        self.dense_res_gen = NN.ModuleList()
        infilters = 1
        for k, s, hid, out in self._struct:
            self.dense_res_gen.append(_DP(_Bottleneck(k, s, infilters, hid, out)))
            infilters += out
        self.dense_res_gen.append(_DP(_conv_holder(1, infilters, 3)))
        self.proj = _DP(_linear_holder(frame_size, H))
        self.stopper = _DP(_linear_holder(1, H))
        self.early_exit_sync = True      # one D2H read of the step count per pass (reference: one per frame)

    def _build_plan(self, dev):
        return P.build_generator_plan(self, dev)

    def forward(self, batch_size=None, length=None, z=None, c=None, u_stop="sample", grad_from=0, defer_tail=False):
        """Returns (x (B, t*frame), s (B, t) stop logits, stop_list: t LongTensors (B, 1), length (B,) samples).
        ``grad_from`` (extension, default 0): samples below this index are forward-only (train.core_step); ``defer_tail``: the
        conv stack of samples [grad_from, B) may still be in flight on the side stream when this returns -- the caller must
        call ``engine.shadow_join(device)`` before reading them (train.core_step does).

        ``u_stop``: uniforms (B, T) for the stop draw ``stop = u < sigmoid(logit)`` (audiogan.py:445-450 draws
        the same Bernoulli through ``multinomial``); "sample" draws them with torch.rand, None never stops."""
        plan = self._get_plan()
        dev = plan.device
        if z is None:
            nframes = div_roundup(length, self._frame_size)
            z = torch.randn(batch_size, nframes, self._noise_size, device=dev)
        else:
            batch_size, nframes, _ = z.shape
        cexp = c.unsqueeze(1).expand(batch_size, nframes, self._embed_size)                 # audiogan.py:425-426
        zc1 = torch.cat([z, cexp, torch.ones(batch_size, nframes, 2, device=dev)], 2)
        if isinstance(u_stop, str):
            u_stop = torch.rand(batch_size, nframes, device=dev)
        token = P.pack(plan)
        x, s, stop, glen = E._GenFn.apply(plan, self._struct, token, zc1, u_stop, self.early_exit_sync, grad_from, defer_tail)
        stop_list = list(stop.long().unsqueeze(2).unbind(1))
        s._ag_stop, s._ag_glen = stop, glen          # raw int32 device copies for the REINFORCE kernel (train.g_update)
        out_len = glen.long() * self._frame_size
        if u_stop is None:
            # no stop is ever drawn: every sample runs all frames.  The host copy rides along so that a following
            # Discriminator.forward does not have to read the lengths back (the reference's tonumpy(length), :516)
            out_len._ag_host = torch.full((batch_size,), nframes * self._frame_size, dtype=torch.int64)
        return x, s, stop_list, out_len


# =========================================================================================
class Discriminator(_PlanOwner):
    """audiogan.py:471-551."""

    def __init__(self, state_size=1024, embed_size=200, num_layers=1, cnn_struct=D_STRUCT):
        NN.Module.__init__(self)
        if num_layers != 1:
            raise NotImplementedError("num_layers > 1 (the reference default and every BASELINE config use 1)")
        self._state_size, self._embed_size, self._num_layers = state_size, embed_size, num_layers
        self._cnn_struct = [list(s) for s in cnn_struct]
        self.cnn_struct = self._cnn_struct
        self.cnn = NN.ModuleList()
        infilters = 1
        for k, s, out in self._cnn_struct:
            self.cnn.append(_DP(_conv_holder(out, infilters, k)))
            infilters = out
        self.frame_size = self._frame_size = infilters
        self.rnn = _PlainLSTM(infilters + embed_size, state_size // 2)
        self.residual_net = _DP(NN.Sequential(_Residual(state_size), _Residual(state_size)))
        self.classifier = _DP(NN.Sequential(_linear_holder(state_size // 2, state_size), NN.Identity(),
                                            _linear_holder(1, state_size // 2)))

    def _build_plan(self, dev):
        return P.build_discriminator_plan(self, dev)

    def forward(self, x, length, c, percent_used=0.1):
        """x (B, L) waveform, length LongTensor (B) in samples, c (B, embed).  Returns
        (logits (B, T'), cnn_outputs [6 x (B, C_i, T_i)], cnn_output_lengths [6 x LongTensor (B)], nframes)."""
        plan = self._get_plan()
        dev = plan.device
        B, L = x.shape
        lens_h, lens_d, Tm, Tmin, lens_out = _length_tables(length, tuple(s for _, s, _ in self._cnn_struct), dev)
        token = P.pack(plan)
        outs = E._DiscCNNFn.apply(plan, self._cnn_struct, token, x, lens_d)
        logits = E._DiscTailFn.apply(plan, token, outs[-1], c, lens_d[-1], Tm, Tmin)
        return logits, list(outs), lens_out, lens_out[-1]


# =========================================================================================
# helper functions the training loop calls (audiogan.py:172-253, :336-359)
# =========================================================================================
# per-layer frame counts of a lengths vector (audiogan.py:533: nframes = (nframes + stride - 1) / stride per conv layer): host
# values (shapes, Tm) and int32 device copies for the kernels' masks.  Keyed by the lengths' VALUES: a training loop sees the same
# few vectors over and over (full-length batches; the D-update's [real | fake] concatenation), so in steady state a forward
# pass issues no host<->device copy for its lengths at all (which also keeps it capturable in a CUDA graph).
_LENS_CACHE = collections.OrderedDict()
_LENS_CACHE_MAX = 32


def _length_tables(length, strides, dev):
    length_h = host_lengths(length)
    key = (dev, strides, length.device, tuple(length_h.tolist()))
    ent = _LENS_CACHE.get(key)
    if ent is None:
        lens_h, nf = [], length_h
        for s in strides:
            nf = (nf + s - 1) // s
            lens_h.append(nf)
        all_d = torch.stack(lens_h, 0).to(torch.int32).to(dev)
        lens_d = [all_d[i] for i in range(len(lens_h))]
        lens_out = []
        for lh, ld in zip(lens_h, lens_d):
            lo = lh.to(length.device)                        # what the reference returns: LongTensors beside the input lengths
            lo._ag_host, lo._ag_dev_i32 = lh, ld
            lens_out.append(lo)
        ent = (lens_h, lens_d, int(lens_h[-1].max()), int(lens_h[-1].min()), lens_out)
        _LENS_CACHE[key] = ent
        while len(_LENS_CACHE) > _LENS_CACHE_MAX:
            _LENS_CACHE.popitem(last=False)
    else:
        _LENS_CACHE.move_to_end(key)
    return ent


def dev_i32(length, device):
    """int32 device copy of a lengths tensor; free when it came out of Discriminator.forward (cached beside it)."""
    d = getattr(length, "_ag_dev_i32", None)
    if d is not None and d.device == torch.device(device):
        return d
    return length.to(device, torch.int32)


def host_lengths(length):
    """int64 CPU copy of a lengths tensor.  CPU tensors and tensors that carry a host copy (Generator.forward without stop
    sampling, cat_lengths) cost nothing; a bare CUDA tensor costs the device-to-host read the reference also pays."""
    h = getattr(length, "_ag_host", None)
    if h is not None:
        return h
    return length.detach().to("cpu", torch.int64)


def cat_lengths(parts, device):
    """torch.cat of lengths tensors.  When every part is known on the host (CPU tensors, Generator.forward without stop
    sampling) the result is a CPU tensor -- Discriminator.forward needs the host values and caches its device tables by
    them, so nothing is copied; otherwise the parts are concatenated on `device`."""
    hosts = [p if not p.is_cuda else getattr(p, "_ag_host", None) for p in parts]
    if all(h is not None for h in hosts):
        out = torch.cat([h.to(torch.int64) for h in hosts], 0)
        out._ag_host = out
        return out
    return torch.cat([p.to(device, non_blocking=True) for p in parts], 0)


def length_mask(size, length):                                   # audiogan.py:204-211
    """1 where t < length[b]; built on the device (the reference fills it in a host loop and uploads it)."""
    dev = length.device if length.is_cuda else torch.device("cuda", torch.cuda.current_device())
    ar = torch.arange(size[1], device=dev).unsqueeze(0)
    return (ar < dev_i32(length, dev).reshape(-1, 1)).float()


def binary_cross_entropy_with_logits_per_sample(input, target, weight=None):    # audiogan.py:187-197
    if not (target.size() == input.size()):
        raise ValueError("Target size ({}) must be the same as input size ({})".format(target.size(), input.size()))
    return E._BCEFn.apply(input, target, weight)


def fourth_moment(v):                                            # audiogan.py:336-339
    return (((v - v.mean(0).unsqueeze(0)) ** 4).sum(0)) ** (1 / 4)


def calc_dists(hidden_states, hidden_state_lengths, gather=None):            # audiogan.py:341-359
    """Feature-matching statistics of the discriminator's conv activations.  The scans over the activations (197 MB per
    discriminator pass at configs[1]: per (sample, channel) mean / centred 2nd / centred 4th moment over time, and their
    gradient back into every activation) are the library's streaming kernels (engine._TimeMomentsFn -> ag_time_moments_*);
    what is left here works on (B, C) arrays: the batch statistics and the reference's list layout
    (means + stds + fourths, three (value, std) pairs per layer each).

    ``gather``: data-parallel hook ``q (3, B_local, C) -> q (3, B_global, C)`` (dist.gather_batch) so that the batch
    statistics are those of the GLOBAL minibatch, as in the reference's single process."""
    means_d, stds_d, fourth_d = [], [], []
    for h, l in zip(hidden_states, hidden_state_lengths):
        nfr = getattr(l, "_ag_dev_i32", None)
        if nfr is None or nfr.device != h.device:
            nfr = l.to(h.device, torch.int32)
        q = E._TimeMomentsFn.apply(h, nfr)                           # (3, B, C): m, s, f of :345-347
        if gather is not None:
            q = gather(q)
        mean, std = q.mean(1), q.std(1)
        fourth = ((q - mean.unsqueeze(1)) ** 4).sum(1) ** (1 / 4)
        for i in range(3):
            means_d.append((mean[i], std[i]))
            stds_d.append((std[i], std[i]))
            fourth_d.append((fourth[i], std[i]))
    return means_d + stds_d + fourth_d


class Embedder(_PlanOwner):
    """Character embedder producing the conditioning vector `c` (audiogan.py:302-334): Embedding(256, 50) ->
    BiLSTM(50 -> 2 x output/2) over the character sequence with per-sample lengths (the reference's dynamic_rnn sorts, packs
    and unpacks, :214-229) -> last hidden states (B, output).  Same constructor, forward signature and state_dict keys
    (`embed.module.weight`, `rnn.weight_ih_l0`, ..., `rnn.bias_hh_l0_reverse`) as the reference.  The recurrence, its input
    projection and every gradient run on the library's fp32 kernels (engine._EmbedFn: hidden size padded 50 -> 64 with
    structural zeros, per-sample lengths in the kernel -- no sort / pack / unpack); the table lookup is torch indexing."""

    def __init__(self, output_size=100, char_embed_size=50, num_layers=1, num_chars=256):
        NN.Module.__init__(self)
        if num_layers != 1:
            raise NotImplementedError("num_layers > 1 (the reference default is 1)")
        self._output_size, self._char_embed_size, self._num_layers = output_size, char_embed_size, num_layers
        self.embed = _DP(NN.Embedding(num_chars, char_embed_size))
        self.rnn = _PlainLSTM(char_embed_size, output_size // 2)

    def _build_plan(self, dev):
        return P.build_embedder_plan(self, dev)

    def set_mode(self, mode):
        """53.6 k parameters, <= ~20 steps: always the fp32 kernels (the conditioning vector is an fp32 quantity)."""
        return self

    def forward(self, chars, length):
        plan = self._get_plan()
        dev = plan.device
        x = self.embed.module(chars.to(dev))                                     # :322 (B, T, E)
        len_long = length.to(dev, torch.int64)
        Tm = int(host_lengths(length).max())
        token = P.pack(plan)
        return E._EmbedFn.apply(plan, token, x[:, :Tm].contiguous(), len_long.to(torch.int32), len_long)


def pin_stopper(g, value=30.0):
    """stop bias = g*sign(v) = -value -> logits ~ -value -> the generator never stops early."""
    with torch.no_grad():
        g.stopper.module.bias_g.fill_(value)
        g.stopper.module.bias_v.fill_(-1.0)
    return g
