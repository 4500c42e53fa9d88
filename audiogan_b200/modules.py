"""Drop-in ``Generator`` / ``Discriminator`` and helper functions with the reference's API.

Same constructor keywords, ``forward`` signatures, return tuples and ``state_dict`` key names
(weight-norm ``_g`` / ``_v`` split *and* the ``DataParallel`` ``.module.`` infix) as
``/root/reference/audiogan.py:361-551``; all arithmetic runs in libaudiogan_b200.so.  Parameters
are fp32 master copies owned by PyTorch; gradients arrive through autograd as usual, so the
reference training loop (``loss.backward()``, ``T.autograd.grad(loss, x)``, ``requires_grad``
toggling, ``opt.step()``) works unchanged.  CUDA only: there is no CPU fallback.
"""
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

    def forward(self, batch_size=None, length=None, z=None, c=None, u_stop="sample"):
        """Returns (x (B, t*frame), s (B, t) stop logits, stop_list: t LongTensors (B, 1), length (B,) samples).

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
        x, s, stop, glen = E._GenFn.apply(plan, self._struct, token, zc1, u_stop, self.early_exit_sync)
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
        length_h = host_lengths(length)                              # host copy of the lengths (reference: tonumpy)
        lens_h, lens_d = [], []
        nf = length_h
        for _, s, _ in self._cnn_struct:                              # audiogan.py:533
            nf = (nf + s - 1) // s
            lens_h.append(nf)
        all_d = torch.stack(lens_h, 0).to(torch.int32).to(dev, non_blocking=True)
        lens_d = [all_d[i] for i in range(len(lens_h))]
        Tm = int(lens_h[-1].max())
        token = P.pack(plan)
        outs = E._DiscCNNFn.apply(plan, self._cnn_struct, token, x, lens_d)
        logits = E._DiscTailFn.apply(plan, token, outs[-1], c, lens_d[-1], Tm, int(lens_h[-1].min()))
        lens_out = [l.to(length.device) for l in lens_h]
        return logits, list(outs), lens_out, lens_out[-1]


# =========================================================================================
# helper functions the training loop calls (audiogan.py:172-253, :336-359)
# =========================================================================================
def host_lengths(length):
    """int64 CPU copy of a lengths tensor.  CPU tensors and tensors that carry a host copy (Generator.forward without stop
    sampling, cat_lengths) cost nothing; a bare CUDA tensor costs the device-to-host read the reference also pays."""
    h = getattr(length, "_ag_host", None)
    if h is not None:
        return h
    return length.detach().to("cpu", torch.int64)


def cat_lengths(parts, device):
    """torch.cat of lengths tensors on `device`, keeping a host copy when every part has one."""
    hosts = [p if not p.is_cuda else getattr(p, "_ag_host", None) for p in parts]
    out = torch.cat([p.to(device, non_blocking=True) for p in parts], 0)
    if all(h is not None for h in hosts):
        out._ag_host = torch.cat([h.to(torch.int64) for h in hosts], 0)
    return out


def length_mask(size, length):                                   # audiogan.py:204-211
    """1 where t < length[b]; built on the device (the reference fills it in a host loop and uploads it)."""
    dev = length.device if length.is_cuda else torch.device("cuda")
    ar = torch.arange(size[1], device=dev).unsqueeze(0)
    return (ar < length.to(dev).reshape(-1, 1)).float()


def binary_cross_entropy_with_logits_per_sample(input, target, weight=None):    # audiogan.py:187-197
    if not (target.size() == input.size()):
        raise ValueError("Target size ({}) must be the same as input size ({})".format(target.size(), input.size()))
    return E._BCEFn.apply(input, target, weight)


def fourth_moment(v):                                            # audiogan.py:336-339
    return (((v - v.mean(0).unsqueeze(0)) ** 4).sum(0)) ** (1 / 4)


def calc_dists(hidden_states, hidden_state_lengths):             # audiogan.py:341-359 ("next" row)
    means_d, stds_d, fourth_d = [], [], []
    for h, l in zip(hidden_states, hidden_state_lengths):
        h = h.float()            # bf16 mode hands the conv activations back in their bf16 storage type
        l = l.to(h.device)
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


class Embedder(NN.Module):
    """Character embedder producing the conditioning vector `c` (audiogan.py:302-334): Embedding(256, 50) ->
    BiLSTM(50 -> 2 x output/2) over the packed character sequence -> last hidden states (B, output).  A "next" row of
    SURVEY 8(f): 53.6 k parameters and <= ~20 character steps, far off the hot path -- it runs on stock torch modules
    (cuDNN); same constructor, forward signature, state_dict keys (`embed.module.weight`, `rnn.*`) as the reference."""

    def __init__(self, output_size=100, char_embed_size=50, num_layers=1, num_chars=256):
        NN.Module.__init__(self)
        self._output_size, self._char_embed_size, self._num_layers = output_size, char_embed_size, num_layers
        self.embed = _DP(NN.Embedding(num_chars, char_embed_size))
        self.rnn = NN.LSTM(char_embed_size, output_size // 2, num_layers, bidirectional=True)

    def forward(self, chars, length):
        from torch.nn.utils.rnn import pack_padded_sequence
        batch_size = chars.size(0)
        seq = self.embed.module(chars).permute(1, 0, 2)                          # :322-324
        packed = pack_padded_sequence(seq, length.detach().cpu(), enforce_sorted=False)   # dynamic_rnn :214-229
        # cuDNN's RNN would otherwise run its GEMMs in TF32 on sm_100 (4e-4 relative on `c`): the conditioning vector is an
        # fp32 quantity in the reference
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            _, (h, _) = self.rnn(packed)
        h = h.permute(1, 0, 2)                                                   # :333
        return h[:, -2:].reshape(batch_size, self._output_size)                   # :334


def pin_stopper(g, value=30.0):
    """stop bias = g*sign(v) = -value -> logits ~ -value -> the generator never stops early."""
    with torch.no_grad():
        g.stopper.module.bias_g.fill_(value)
        g.stopper.module.bias_v.fill_(-1.0)
    return g
