"""Flat parameter plans: weight-norm table, packed GEMM operand layouts and gradient un-packing.

The reference re-computes ``w = g * v / ||v||`` inside every module call through forward
pre-hooks (audiogan.py:77-80 -- 736 ``_weight_norm_interface`` calls per step).  Here
one multi-tensor kernel writes every effective weight of a network into one flat fp32 buffer
(``wflat``) and one gather kernel re-lays them out into every operand layout the GEMM / recurrent
kernels consume (``pflat``): im2col-ordered conv filters, phase-decomposed transposed-conv and
data-gradient filters, concatenated / transposed recurrent matrices.  Weight gradients are
accumulated by the kernels into ``gpflat`` (one region per weight, bias gradient in an extra
column) and come back through the inverse index map and the weight-norm backward kernel.

All layouts are described with ordinary torch indexing on *index tensors* (``-1`` = structural
zero), so a layout is declared exactly the way one would write the corresponding reshape.
"""
import os

import torch
from torch.autograd.function import once_differentiable

from . import kernels as K

_ALIGN = 64     # elements: every region starts 256-byte aligned

# Weight-gradient work is decided at BACKWARD time.  ctx.needs_input_grad is fixed when the forward runs, so a
# ``torch.autograd.grad(loss, x)`` pass over a graph whose parameters require grad (x_grad_norm audiogan.py:769-770, the FGSM
# moves :145 and :131) would otherwise accumulate weight gradients into plan.gpflat that no _PackFn.backward consumes -- the
# reference never lets autograd.grad touch p.grad.  The loop helpers wrap those calls in ``no_weight_grads()``.
_wgrad_off = 0


class no_weight_grads:
    """Context manager: backward passes run inside it produce data gradients only (no weight gradients, no early all-reduce)."""

    def __enter__(self):
        global _wgrad_off
        _wgrad_off += 1
        return self

    def __exit__(self, *exc):
        global _wgrad_off
        _wgrad_off -= 1
        return False


def wgrad_enabled():
    return _wgrad_off == 0


def _round(n, a=_ALIGN):
    return (n + a - 1) // a * a


class Regions:
    """Named regions in one flat buffer."""

    def __init__(self):
        self.off = {}
        self.shape = {}
        self.size = 0

    def add(self, name, shape):
        shape = tuple(int(s) for s in shape)
        n = 1
        for s in shape:
            n *= s
        self.off[name] = self.size
        self.shape[name] = shape
        self.size = _round(self.size + n)
        return self.index(name)

    def index(self, name):
        n = 1
        for s in self.shape[name]:
            n *= s
        return torch.arange(self.off[name], self.off[name] + n, dtype=torch.int64).view(self.shape[name])

    def view(self, flat, name):
        n = 1
        for s in self.shape[name]:
            n *= s
        return flat[self.off[name]:self.off[name] + n].view(self.shape[name])


class WRef(tuple):
    """(flat buffer, element offset) naming a packed weight layout; lets the GEMM wrappers swap in the bf16
    copy (row stride padded to 16 bytes for TMA) when the plan runs in bf16 mode."""

    def __new__(cls, plan, name, extra, flat, off):
        obj = tuple.__new__(cls, (flat, off + extra))
        obj.plan, obj.name, obj.extra = plan, name, extra
        return obj

    def bf16(self, ldb):
        """-> ((pflat16, offset), ld16) if `ldb` is the natural row stride of the layout, else None."""
        pl = self.plan
        if self.name not in pl.off16 or pl.ldnat[self.name] != ldb:
            return None
        row, col = divmod(self.extra, ldb)
        return (pl.pflat16, pl.off16[self.name] + row * pl.ld16[self.name] + col), pl.ld16[self.name]


class NetPlan:
    """Everything that depends only on a network's parameter shapes, built once per (module, device)."""

    def __init__(self, device):
        self.device = device
        self.canon = Regions()          # effective weights (after weight-norm), canonical torch layouts
        self.pack = Regions()           # GEMM / recurrent operand layouts
        self.gpack = Regions()          # packed weight gradients
        self._wn = []                   # (canon name, kind, v param, g param)
        self._pack_idx = {}
        self._unpack_idx = {}
        self.params = []                # parameters in module.parameters() order
        self.mode = "fp32"
        # recurrent products in bf16 mode: 1 = mma.sync path (default: measured faster end to end, see
        # profiles/r1_lstm_phase_cycles.txt), 2 = tcgen05 / TMEM path (AUDIOGAN_LSTM=tcgen05)
        self.lstm_prec = 2 if os.environ.get("AUDIOGAN_LSTM", "") == "tcgen05" else 1
        # AUDIOGAN_LSTM=grid keeps the grid-barrier kernels (no cluster / TMEM-resident variants): A/B timing aid
        self.lstm_flags = 1 if os.environ.get("AUDIOGAN_LSTM", "") == "grid" else 0
        # data-parallel: a dist.GradSync that all-reduces packed weight-gradient regions while backward is still running
        self.early_sync = None
        # kernel family (+ decline reason) the recurrent calls of the last pass ran on: {"g_fwd" | "g_bwd" | "d_fwd" | "d_bwd": str}
        self.last_path = {}

    # -- declaration ----------------------------------------------------------------------
    def weight(self, name, v, g=None):
        """Declare an effective weight: weight-normed (v, g) pair or plain parameter."""
        idx = self.canon.add(name, v.shape)
        self._wn.append((name, 0 if g is not None else 1, v, g))
        return idx

    def layout(self, name, idx):
        """Declare a packed operand layout from an index tensor into the canonical buffer."""
        self.pack.add(name, idx.shape)
        self._pack_idx[name] = idx.contiguous()

    def grad_region(self, name, shape):
        return self.gpack.add(name, shape)

    def grad_of(self, canon_name, idx):
        """Canonical gradient of `canon_name` = gpflat[idx] (idx has the canonical shape)."""
        assert tuple(idx.shape) == self.canon.shape[canon_name], (canon_name, idx.shape, self.canon.shape[canon_name])
        self._unpack_idx[canon_name] = idx.contiguous()

    # -- materialisation ------------------------------------------------------------------
    def finalize(self, params):
        dev = self.device
        self.params = list(params)
        self.wflat = torch.zeros(self.canon.size, device=dev)
        self.dwflat = torch.zeros(self.canon.size, device=dev)
        self.pflat = torch.zeros(self.pack.size, device=dev)
        self.gpflat = torch.zeros(self.gpack.size, device=dev)
        pidx = torch.full((self.pack.size,), -1, dtype=torch.int64)
        for name, idx in self._pack_idx.items():
            o = self.pack.off[name]
            pidx[o:o + idx.numel()] = idx.reshape(-1)
        self.idx_pack = pidx.to(torch.int32).to(dev)
        # bf16 copies of the 2-D+ layouts, rows padded to a multiple of 8 elements (16 bytes: TMA global stride rule)
        self.off16, self.ld16, self.ldnat, n16, parts = {}, {}, {}, 0, []
        for name, idx in self._pack_idx.items():
            if idx.dim() < 2:
                continue
            kk = idx.shape[-1]
            kp = _round(kk, 8)
            flat2 = idx.reshape(-1, kk)
            padded = torch.full((flat2.shape[0], kp), -1, dtype=torch.int64)
            padded[:, :kk] = flat2
            self.off16[name], self.ld16[name], self.ldnat[name] = n16, kp, kk
            parts.append((n16, padded.reshape(-1)))
            n16 = _round(n16 + padded.numel())
        pidx16 = torch.full((max(n16, 1),), -1, dtype=torch.int64)
        for o, v in parts:
            pidx16[o:o + v.numel()] = v
        self.idx_pack16 = pidx16.to(torch.int32).to(dev)
        self.pflat16 = torch.zeros(max(n16, 1), device=dev, dtype=torch.bfloat16)
        uidx = torch.full((self.canon.size,), -1, dtype=torch.int64)
        for name, _, _, _ in self._wn:
            idx = self._unpack_idx[name]
            o = self.canon.off[name]
            uidx[o:o + idx.numel()] = idx.reshape(-1)
        self.idx_unpack = uidx.to(torch.int32).to(dev)
        # flat gradient work buffer in parameter order
        self.poff, n = {}, 0
        for p in self.params:
            self.poff[id(p)] = n
            n = _round(n + p.numel(), 4)
        self.ngrad = n
        self.gpersist = None
        self._bind_grad_buffer(torch.zeros(n, device=dev))
        self.signature = tuple(p.data_ptr() for p in self.params)
        self._const_len = {}
        return self

    def _bind_grad_buffer(self, buf):
        """(re)build the weight-norm table so that the backward kernel writes the parameter gradients into `buf`"""
        dev = self.device
        self.gwork = buf
        rows = sum(v.shape[0] for _, _, v, _ in self._wn)
        self.norms = torch.zeros(rows, device=dev)
        entries, r0 = [], 0
        for name, kind, v, g in self._wn:
            nrow = v.shape[0]
            ncol = v.numel() // nrow
            o = self.canon.off[name]
            e = dict(v=v.data, g=(g.data if g is not None else None), w=(self.wflat, o), norm=(self.norms, r0),
                     dw=(self.dwflat, o), dv=(self.gwork, self.poff[id(v)]),
                     dg=((self.gwork, self.poff[id(g)]) if g is not None else None),
                     rows=nrow, cols=ncol, kind=kind)
            entries.append(e)
            r0 += nrow
        self.wn_tab, self.wn_rows, self.wn_total = K.wn_table(entries, dev)
        self.wn_n = len(entries)
        self._packed_key = None                                  # the norms buffer is new: re-run weight-norm forward

    def set_grad_buffer(self, buf):
        """Data-parallel over peer memory (dist.PeerGradSync): the flat gradient buffer lives in symmetric memory that every
        rank maps, and ``p.grad`` become PERSISTENT views of it (no per-step copy) -- each backward pass overwrites them, which
        is what a training loop that calls ``zero_grad()`` before ``backward()`` expects."""
        assert buf.numel() >= self.ngrad and buf.dtype == torch.float32 and buf.is_contiguous()
        buf.zero_()
        self._bind_grad_buffer(buf)
        self.gpersist = buf

    def P(self, name):
        return self.pack.view(self.pflat, name)

    def Poff(self, name, extra=0):
        return WRef(self, name, extra, self.pflat, self.pack.off[name])

    def GP(self, name):
        return self.gpack.view(self.gpflat, name)

    def GPoff(self, name, extra=0):
        return WRef(self, name, extra, self.gpflat, self.gpack.off[name])

    def W(self, name):
        return self.canon.view(self.wflat, name)

    def const_len(self, B, value):
        """int32 [B] filled with `value` (bounds-only masks for the view-GEMM epilogue)."""
        key = (int(B), int(value))
        t = self._const_len.get(key)
        if t is None:
            t = torch.full((int(B),), int(value), dtype=torch.int32, device=self.device)
            self._const_len[key] = t
        return t

    # -- per-step work --------------------------------------------------------------------
    def pack_key(self):
        """Changes whenever a parameter may have changed: torch's version counters (in-place torch ops, load_state_dict)
        plus the epoch the fused optimizers bump (their kernels write through raw pointers, invisible to torch)."""
        return (self.mode, sum(p._version for p in self.params), sum(getattr(p, "_ag_epoch", 0) for p in self.params))

    def pack_forward(self):
        # The training step touches each net twice per iteration with the same values (G: D-update then G-update; D: G-update
        # then the next D-update): weight-norm + operand packing run once per parameter change, not once per call.
        key = self.pack_key()
        if key == getattr(self, "_packed_key", None):
            return
        self._packed_key = key
        K.wn_fwd(self.wn_tab, self.wn_rows, self.wn_n, self.wn_total)
        K.gather(self.pflat, self.wflat, self.idx_pack)
        if self.mode == "bf16":
            K.gather(self.pflat16, self.wflat, self.idx_pack16)

    def reduce_span(self, first, last):
        """The weight-gradient regions first..last (contiguous in gpflat) are complete: start their all-reduce now, on NCCL's
        stream, so it overlaps the rest of backward.  Un-packing and weight-norm backward are linear in the packed gradients,
        so reducing before them gives the same sums as reducing p.grad afterwards."""
        es = self.early_sync
        if es is None or es.world == 1:
            return
        n = 1
        for d in self.gpack.shape[last]:
            n *= d
        es.reduce_async(self, self.gpack.off[first], self.gpack.off[last] + n)

    def pack_backward(self):
        """Consume gpflat -> fresh flat gradient buffer in parameter order (views per parameter)."""
        if self.early_sync is not None and self.early_sync.world > 1:
            self.early_sync.finish(self)                 # the remaining regions, then wait for every bucket; GradSync then
                                                         # skips these parameters (their gradients are already global sums)
        K.gather(self.dwflat, self.gpflat, self.idx_unpack)
        self.gpflat.zero_()
        K.wn_bwd(self.wn_tab, self.wn_rows, self.wn_n, self.wn_total)
        out = self.gwork.clone() if self.gpersist is None else self.gwork
        return [out[self.poff[id(p)]:self.poff[id(p)] + p.numel()].view(p.shape) for p in self.params]


class _PackFn(torch.autograd.Function):
    """wflat/pflat <- params.  Returns a 1-element token every consumer takes as an input, so that
    autograd runs this node's backward after all of them (they accumulate into plan.gpflat)."""

    @staticmethod
    def forward(ctx, plan, *params):
        ctx.plan = plan
        plan.pack_forward()
        return torch.zeros(1, device=plan.device)

    @staticmethod
    @once_differentiable
    def backward(ctx, gtoken):
        plan = ctx.plan
        grads = plan.pack_backward()
        out = [g if ctx.needs_input_grad[i + 1] else None for i, g in enumerate(grads)]
        return (None, *out)


def pack(plan):
    return _PackFn.apply(plan, *plan.params)


# =========================================================================================
# Generator plan (audiogan.py:362-410)
# =========================================================================================
def build_generator_plan(mod, device):
    H, F, NZ = mod._state_size, mod._frame_size, mod._noise_size + mod._embed_size
    FP = (F + 1 + 7) // 8 * 8
    pl = NetPlan(device)
    cell = mod.rnn[0].module
    wih = pl.weight("rnn.wih", cell.weight_ih_v, cell.weight_ih_g)
    whh = pl.weight("rnn.whh", cell.weight_hh_v, cell.weight_hh_g)
    bhh = pl.weight("rnn.bhh", cell.bias_hh_v, cell.bias_hh_g)
    bih = pl.weight("rnn.bih", cell.bias_ih_v, cell.bias_ih_g)
    convs = []
    for li, (k, s, hid, out) in enumerate(mod._struct):
        blk = mod.dense_res_gen[li].module
        cw = pl.weight("c%d.w" % li, blk.conv.weight_v, blk.conv.weight_g)
        cb = pl.weight("c%d.b" % li, blk.conv.bias_v, blk.conv.bias_g)
        dw = pl.weight("d%d.w" % li, blk.deconv.weight_v, blk.deconv.weight_g)
        db = pl.weight("d%d.b" % li, blk.deconv.bias_v, blk.deconv.bias_g)
        convs.append((cw, cb, dw, db))
    fin = mod.dense_res_gen[len(mod._struct)].module
    fw = pl.weight("f.w", fin.weight_v, fin.weight_g)
    fb = pl.weight("f.b", fin.bias_v, fin.bias_g)
    pw = pl.weight("proj.w", mod.proj.module.weight_v, mod.proj.module.weight_g)
    pb = pl.weight("proj.b", mod.proj.module.bias_v, mod.proj.module.bias_g)
    sw = pl.weight("stop.w", mod.stopper.module.weight_v, mod.stopper.module.weight_g)
    sb = pl.weight("stop.b", mod.stopper.module.bias_v, mod.stopper.module.bias_g)

    neg = lambda *shape: torch.full(shape, -1, dtype=torch.int64)
    # recurrent operands
    pl.layout("w1", torch.cat([whh, wih[:, :F]], 1))                                   # [4H, H+F]
    pl.layout("wz", torch.cat([wih[:, F:], bih[:, None], bhh[:, None]], 1))            # [4H, NZ+2]
    pl.layout("w2", torch.cat([pw, sw], 0))                                            # [F+1, H]
    pl.layout("b2", torch.cat([pb, sb], 0))
    pl.layout("w1t", torch.cat([whh.t(), pw.t(), sw.t(), neg(H, FP - F - 1)], 1))       # [H, 4H+FP]
    pl.layout("wxt", wih[:, :F].t())                                                   # [F, 4H]
    pl.layout("wzt", wih[:, F:].t())                                                   # [NZ, 4H]
    g1 = pl.grad_region("w1", (4 * H, H + F))
    pl.ZP = (NZ + 2 + 7) // 8 * 8                  # row pitch of the wz gradient region (and of the bf16 [z|c|1|1] operand)
    gz = pl.grad_region("wz", (4 * H, pl.ZP))
    g2 = pl.grad_region("w2", (FP, H + 1))
    pl.grad_of("rnn.wih", torch.cat([g1[:, H:], gz[:, :NZ]], 1))
    pl.grad_of("rnn.whh", g1[:, :H])
    pl.grad_of("rnn.bih", gz[:, NZ])
    pl.grad_of("rnn.bhh", gz[:, NZ + 1])
    pl.grad_of("proj.w", g2[:F, :H])
    pl.grad_of("proj.b", g2[:F, H])
    pl.grad_of("stop.w", g2[F:F + 1, :H])
    pl.grad_of("stop.b", g2[F:F + 1, H])
    # conv stack.  The dense channel-last buffer keeps every channel group in a slot padded to a multiple of 8
    # channels (x: 1 -> 8, then each block's `out`), so channel prefixes, slot offsets and the row pitch are all
    # 16/32-byte aligned: vector loads / stores in the GEMM kernels.  pos[c] = padded position of canonical channel c.
    # AUDIOGAN_SLOT (A/B knob, default 16): slot granularity in channels.  16 bf16 channels = 32 bytes = one DRAM sector, and the
    # default net's slots (16, 16, 32, 32, 32) then add up to 128 channels = a 256-byte row pitch, so every slot of every row is
    # sector-aligned: a block's output write touches exactly its own sectors and a prefix read exactly the prefix's.  With
    # 8-channel slots the pitch was 240 bytes (rows alternately misaligned by half a sector) and the transposed-conv launches
    # moved 2.1x their algorithmic DRAM bytes (profiles/r1_bf16_gemm_nt_tma_metrics.txt).
    SL = int(os.environ.get("AUDIOGAN_SLOT", "16"))
    r8 = lambda n: (n + SL - 1) // SL * SL
    pos, slot_off, nxt = [0], [], r8(1)
    for (_, _, _, out) in mod._struct:
        slot_off.append(nxt)
        pos += list(range(nxt, nxt + out))
        nxt += r8(out)
    CTp = nxt
    pos_t = torch.tensor(pos, dtype=torch.int64)

    def pad_ch(idx, dim, cin, cp):
        """scatter the `cin` canonical channels of `idx` along `dim` into their `cp` padded positions (-1 elsewhere)"""
        shape = list(idx.shape)
        shape[dim] = cp
        out_ = torch.full(shape, -1, dtype=torch.int64)
        out_.index_copy_(dim, pos_t[:cin], idx)
        return out_

    def unpad_ch(idx, dim, cin):
        return idx.index_select(dim, pos_t[:cin])

    cin = 1
    pl.cinp, pl.coff, pl.skip_off, pl.q2c, pl.Kq, pl.alay = [], [], [], [], [], []
    PFX64 = int(os.environ.get("AUDIOGAN_PFX64", "32"))       # A/B knob: narrowest channel prefix on the 128-byte-box layout
    for li, (k, s, hid, out) in enumerate(mod._struct):
        cw, cb, dw, db = convs[li]
        kd = k - 1
        assert kd == 2 * s and (k - 1) // 2 == s, "dense_res_bottleneck layer must have kernel = 2*stride + 1"
        cp = slot_off[li]                                   # padded prefix read by this block == slot it writes to
        pl.cinp.append(cp)
        pl.coff.append(slot_off[li])
        if cin >= out:                                      # skip = last `out` canonical channels: must be one slot run
            sk = pos[cin - out:cin]
            assert sk == list(range(sk[0], sk[0] + out)), "dense skip must cover one contiguous padded slot"
            pl.skip_off.append(sk[0])
        else:
            pl.skip_off.append(-1)
        cwp = pad_ch(cw, 1, cin, cp)                                                     # [hid, cp, k]
        pl.layout("c%d.w" % li, cwp.permute(0, 2, 1).reshape(hid, k * cp))              # (j, ci_p)
        pl.layout("c%d.b" % li, cb)
        # transposed conv as a GEMM over 2 taps: [(r', co), (u, ci)] = Wd[ci, co, s*(1-u) + r']
        d4 = dw.view(hid, out, 2, s).flip(2)                                            # [ci, co, u, r']
        pl.layout("d%d.w" % li, d4.permute(3, 1, 2, 0).reshape(s * out, 2 * hid))
        pl.layout("d%d.b" % li, db)
        pl.layout("d%d.wg" % li, dw.permute(0, 2, 1).reshape(hid, kd * out))            # [ci, (j, co)]
        # conv data-gradient over 3 taps: [(r', ci_p), (u, h)] = Wc[h, ci, s*(2-u) + r']
        cpad = torch.cat([cwp, neg(hid, cp, 3 * s - k)], 2).view(hid, cp, 3, s).flip(2)   # [h, ci_p, u, r']
        pl.layout("c%d.wg" % li, cpad.permute(3, 1, 2, 0).reshape(s * cp, 3 * hid))
        # the same filter for the TMA-fed kernels' channel-prefix view (include/audiogan_b200.h: a_layout 1 / 2); its gradient region +
        # the map back to "c%d.w"'s.  Narrow prefixes: columns ordered (channel group of 8, tap padded to a multiple of 8, channel),
        # 16-byte TMA boxes.  Prefixes of PFX64 channels and more: columns ordered (tap, channel group of 64, channel), 128-byte boxes
        # (the TMA request rate bounds the 16-byte variant: 0.6x the producer-warp kernel at 88 channels).
        if cp >= PFX64:
            G_ = (cp + 63) // 64
            Kq = k * G_ * 64
            cq = torch.cat([cwp, neg(hid, G_ * 64 - cp, k)], 1)                                  # [h, G*64, tap]
            pl.layout("c%d.wq" % li, cq.permute(0, 2, 1).reshape(hid, Kq))
            gq = pl.grad_region("c%d.wq" % li, (hid, Kq + 1))
            gqv = gq[:, :Kq].view(hid, k, G_ * 64)[:, :, :cp].reshape(hid, k * cp)               # -> (tap, ci_p)
            pl.alay.append(2)
        else:
            G_, KT = cp // 8, (k + 7) // 8
            Kq = G_ * KT * 64
            cq = torch.cat([cwp, neg(hid, cp, KT * 8 - k)], 2).view(hid, G_, 8, KT * 8)          # [h, g, c8, tap]
            pl.layout("c%d.wq" % li, cq.permute(0, 1, 3, 2).reshape(hid, Kq))
            gq = pl.grad_region("c%d.wq" % li, (hid, Kq + 1))
            gqv = gq[:, :Kq].view(hid, G_, KT * 8, 8)[:, :, :k].permute(0, 2, 1, 3).reshape(hid, k * cp)      # -> (tap, ci_p)
            pl.alay.append(1)
        pl.q2c.append((torch.cat([gqv, gq[:, Kq:]], 1) - pl.gpack.off["c%d.wq" % li]).to(torch.int32).contiguous().to(device))
        pl.Kq.append(Kq)
        gc = pl.grad_region("c%d.w" % li, (hid, k * cp + 1))
        gd = pl.grad_region("d%d.w" % li, (hid, kd * out))
        gb = pl.grad_region("d%d.b" % li, (out,))
        pl.grad_of("c%d.w" % li, unpad_ch(gc[:, :k * cp].reshape(hid, k, cp).permute(0, 2, 1), 1, cin))
        pl.grad_of("c%d.b" % li, gc[:, k * cp])
        pl.grad_of("d%d.w" % li, gd.view(hid, kd, out).permute(0, 2, 1))
        pl.grad_of("d%d.b" % li, gb)
        cin += out
    CT = CTp
    fwp = pad_ch(fw, 1, cin, CTp)                                                      # [1, CTp, 3]
    pl.layout("f.w", fwp.permute(0, 2, 1).reshape(1, 3 * CT))
    pl.layout("f.b", fb)
    pl.layout("f.wg", fwp[0].flip(1))                                                  # [CTp, 3]: W[0, ci, 2-kk]
    gf = pl.grad_region("f.w", (1, 3 * CT + 1))
    pl.grad_of("f.w", unpad_ch(gf[:, :3 * CT].reshape(1, 3, CT).permute(0, 2, 1), 1, cin))
    pl.grad_of("f.b", gf[:, 3 * CT])
    pl.CT, pl.H, pl.F, pl.NZ, pl.FP = CT, H, F, NZ, FP
    return pl.finalize(mod.parameters())


# =========================================================================================
# Discriminator plan (audiogan.py:472-512)
# =========================================================================================
def build_discriminator_plan(mod, device):
    S, E = mod._state_size, mod._embed_size
    H = S // 2
    pl = NetPlan(device)
    neg = lambda *shape: torch.full(shape, -1, dtype=torch.int64)
    cin = 1
    pl.d_ntap = []
    for i, (k, s, cout) in enumerate(mod._cnn_struct):
        assert k == 7 and s in (1, 2), "discriminator conv layers: kernel 7, stride 1 or 2 (audiogan.py:476; cfg 5 adds stride-1 layers)"
        ntap = (k + s - 1) // s                             # taps per output phase of the data gradient
        pl.d_ntap.append(ntap)
        cm = mod.cnn[i].module
        w = pl.weight("c%d.w" % i, cm.weight_v, cm.weight_g)
        b = pl.weight("c%d.b" % i, cm.bias_v, cm.bias_g)
        pl.layout("c%d.w" % i, w.permute(0, 2, 1).reshape(cout, k * cin))
        pl.layout("c%d.b" % i, b)
        # data gradient over ntap taps per phase: [(r', ci), (u, co)] = W[co, ci, s*(ntap-1-u) + r']
        wp = torch.cat([w, neg(cout, cin, ntap * s - k)], 2).view(cout, cin, ntap, s).flip(2)   # [co, ci, u, r']
        pl.layout("c%d.wg" % i, wp.permute(3, 1, 2, 0).reshape(s * cin, ntap * cout))
        g = pl.grad_region("c%d.w" % i, (cout, k * cin + 1))
        pl.grad_of("c%d.w" % i, g[:, :k * cin].reshape(cout, k, cin).permute(0, 2, 1))
        pl.grad_of("c%d.b" % i, g[:, k * cin])
        cin = cout
    Cf = cin
    r = mod.rnn
    wih, whh, bih, bhh = [], [], [], []
    for d, sfx in enumerate(("", "_reverse")):
        wih.append(pl.weight("rnn.wih%d" % d, getattr(r, "weight_ih_l0" + sfx)))
        whh.append(pl.weight("rnn.whh%d" % d, getattr(r, "weight_hh_l0" + sfx)))
        bih.append(pl.weight("rnn.bih%d" % d, getattr(r, "bias_ih_l0" + sfx)))
        bhh.append(pl.weight("rnn.bhh%d" % d, getattr(r, "bias_hh_l0" + sfx)))
    pl.layout("wih", torch.cat([torch.cat([wih[d], bih[d][:, None], bhh[d][:, None]], 1) for d in range(2)], 0))
    pl.layout("w1", torch.stack(whh, 0))                                               # [2, 4H, H]
    pl.layout("w1t", torch.stack([w.t() for w in whh], 0))                             # [2, H, 4H]
    pl.layout("wiht", torch.cat([w.t() for w in wih], 1))                              # [Cf+E, 8H]
    gi = pl.grad_region("wih", (8 * H, Cf + E + 2))
    gh = pl.grad_region("whh", (2, 4 * H, H))
    for d in range(2):
        rows = slice(d * 4 * H, (d + 1) * 4 * H)
        pl.grad_of("rnn.wih%d" % d, gi[rows, :Cf + E])
        pl.grad_of("rnn.bih%d" % d, gi[rows, Cf + E])
        pl.grad_of("rnn.bhh%d" % d, gi[rows, Cf + E + 1])
        pl.grad_of("rnn.whh%d" % d, gh[d])
    lin = [("r0", mod.residual_net.module[0].linear), ("r1", mod.residual_net.module[1].linear),
           ("k0", mod.classifier.module[0]), ("k2", mod.classifier.module[2])]
    for name, m in lin:
        w = pl.weight(name + ".w", m.weight_v, m.weight_g)
        b = pl.weight(name + ".b", m.bias_v, m.bias_g)
        pl.layout(name + ".w", w)
        pl.layout(name + ".wt", w.t())
        pl.layout(name + ".b", b)
        n, kk = w.shape
        g = pl.grad_region(name + ".w", (n, kk + 1))
        pl.grad_of(name + ".w", g[:, :kk])
        pl.grad_of(name + ".b", g[:, kk])
    pl.S, pl.H, pl.E, pl.Cf = S, H, E, Cf
    return pl.finalize(mod.parameters())


# =========================================================================================
# Embedder plan (audiogan.py:302-334): BiLSTM(50 -> 2 x 50) over <= ~20 characters
# =========================================================================================
def build_embedder_plan(mod, device):
    """The recurrent kernels want a hidden size that tiles onto CTAs: H = output/2 (50 by default) is padded to HP = 64 with
    structural zeros.  A padded unit has zero weights and zero bias, so its cell stays c = 0.5 c + 0.5 * 0 = 0 and its output
    h = 0.5 tanh(0) = 0 at every step: the padded network computes the reference's numbers exactly."""
    H, E = mod._output_size // 2, mod._char_embed_size
    HP = (H + 15) // 16 * 16
    pl = NetPlan(device)
    neg = lambda *shape: torch.full(shape, -1, dtype=torch.int64)
    r = mod.rnn

    def pad_rows(idx):                       # [4H, X] -> [4HP, X], gate-major
        out = neg(4 * HP, idx.shape[1])
        for g in range(4):
            out[g * HP:g * HP + H] = idx[g * H:(g + 1) * H]
        return out

    def pad_cols(idx):                       # [R, H] -> [R, HP]
        return torch.cat([idx, neg(idx.shape[0], HP - H)], 1)

    wih, whh, bih, bhh = [], [], [], []
    for d, sfx in enumerate(("", "_reverse")):
        wih.append(pl.weight("rnn.wih%d" % d, getattr(r, "weight_ih_l0" + sfx)))
        whh.append(pl.weight("rnn.whh%d" % d, getattr(r, "weight_hh_l0" + sfx)))
        bih.append(pl.weight("rnn.bih%d" % d, getattr(r, "bias_ih_l0" + sfx)))
        bhh.append(pl.weight("rnn.bhh%d" % d, getattr(r, "bias_hh_l0" + sfx)))
    pl.layout("wih", torch.cat([pad_rows(torch.cat([wih[d], bih[d][:, None], bhh[d][:, None]], 1)) for d in range(2)], 0))
    w1 = [pad_rows(pad_cols(whh[d])) for d in range(2)]
    pl.layout("w1", torch.stack(w1, 0))                                                # [2, 4HP, HP]
    pl.layout("w1t", torch.stack([w.t() for w in w1], 0))                              # [2, HP, 4HP]
    pl.layout("wiht", torch.cat([pad_rows(wih[d]).t() for d in range(2)], 1))          # [E, 8HP]
    gi = pl.grad_region("wih", (8 * HP, E + 2))
    gh = pl.grad_region("whh", (2, 4 * HP, HP))
    for d in range(2):
        rows = torch.cat([torch.arange(d * 4 * HP + g * HP, d * 4 * HP + g * HP + H) for g in range(4)])
        pl.grad_of("rnn.wih%d" % d, gi[rows, :E])
        pl.grad_of("rnn.bih%d" % d, gi[rows, E])
        pl.grad_of("rnn.bhh%d" % d, gi[rows, E + 1])
        pl.grad_of("rnn.whh%d" % d, gh[d][rows - d * 4 * HP, :H])
    pl.H, pl.HP, pl.E = H, HP, E
    return pl.finalize(list(mod.rnn.parameters()))
