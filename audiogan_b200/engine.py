"""autograd Functions that drive the sm_100a kernels for the generator and the discriminator.

Three coarse Functions cover the whole hot path (plus plan._PackFn for the weights):
  _GenFn       Generator.forward body  audiogan.py:428-468 (recurrence + frame assembly + conv stack)
  _DiscCNNFn   Discriminator conv loop audiogan.py:527-536
  _DiscTailFn  BiLSTM + residual + classifier audiogan.py:537-549
Every arithmetic step is a call into libaudiogan_b200.so (kernels.py); torch is used for device
memory, views and autograd bookkeeping only.
"""
import os

import torch
from torch.autograd.function import once_differentiable

from . import kernels as K
from . import _abi as A
from .plan import pack, wgrad_enabled  # noqa: F401  (pack re-exported)

GPAD = 8        # zero rows either side of the generator's dense channel-last buffer
DPAD = 3        # zero rows either side of the discriminator's channel-last activations


# bf16 mode keeps the conv stacks' activations AND their gradients in HBM as bf16: they only feed tensor-core GEMMs (which
# round their operands to bf16 anyway) and HBM-bound element-wise kernels, so the storage type sets the step's memory
# traffic, not its accuracy.  AUDIOGAN_ACT=fp32 keeps fp32 storage (the A/B switch used for profiles/).
_ACT_BF16 = os.environ.get("AUDIOGAN_ACT", "bf16") != "fp32"


# ---------------------------------------------------------------------------------------------------------------------
# "Shadow" work.  The recurrent kernels are latency-bound and occupy 64-128 of the 148 SMs for 0.6-1.1 ms each; throughput-bound
# work that nothing on the critical path waits for (the tail's weight gradients; the conv stack of the generator samples that
# only the G-update will read) is queued here and launched on a side stream right AFTER such a kernel has been issued, so it
# runs on the idle SMs underneath it.  AUDIOGAN_OVERLAP=0 runs everything in line.
# ---------------------------------------------------------------------------------------------------------------------
_OVERLAP = os.environ.get("AUDIOGAN_OVERLAP", "1") != "0"
_STEPWISE = os.environ.get("AUDIOGAN_STEPWISE", "1") != "0"   # per-frame GEMM recurrence for hidden sizes that cannot be resident
_TMA_CIN = int(os.environ.get("AUDIOGAN_TMA_CIN", "32"))     # widest channel prefix whose conv runs on the 4-D-map TMA kernel
_side_streams = {}
_shadow_jobs = {}          # device -> [callable]


def _side_stream(dev):
    s = _side_streams.get(dev)
    if s is None:
        s = _side_streams[dev] = torch.cuda.Stream(dev)
    return s


def shadow_submit(dev, fn):
    """queue `fn` (kernel launches whose operands the caller keeps alive until shadow_join)"""
    _shadow_jobs.setdefault(dev, []).append(fn)


def shadow_ready(dev):
    """call BEFORE issuing a latency-bound kernel: marks the point queued jobs may start after.  Returns a token or None."""
    if not _shadow_jobs.get(dev):
        return None
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))
    return ev


def shadow_launch(dev, token):
    """call right AFTER the latency-bound kernel has been issued on the current stream"""
    jobs = _shadow_jobs.get(dev)
    if token is None or not jobs:
        return
    side = _side_stream(dev)
    side.wait_event(token)
    with torch.cuda.stream(side):
        for fn in jobs:
            fn()
    jobs.clear()
    _side_streams[(dev, "busy")] = True


def shadow_join(dev):
    """everything queued so far has been issued and the current stream waits for it"""
    jobs = _shadow_jobs.get(dev)
    if jobs:                                                   # no launch point came by: run them now, in line
        for fn in jobs:
            fn()
        jobs.clear()
    if _side_streams.pop((dev, "busy"), False):
        torch.cuda.current_stream(dev).wait_stream(_side_stream(dev))

# ---------------------------------------------------------------------------------------------------------------------
# Step-wise generator recurrence (csrc/lstm_step.cu): hidden sizes whose recurrent weights cannot stay resident on chip
# (--gstatesize 2048: 36.8 MB in bf16).  One small-M tensor-core GEMM per frame over the bf16 weights (which stay in the L2
# between frames) + the point-wise kernels between them; 4 launches per frame, all inside the step's CUDA graph.
# ---------------------------------------------------------------------------------------------------------------------
def _gen_stepwise_fwd(plan, B, Tcap, H, F, FP, pre, hbuf16, xbuf, xbuf16, gates, cbuf, sbuf):
    dev = plan.device
    KP = (H + F + 7) // 8 * 8
    hx = torch.zeros(B, KP, device=dev, dtype=torch.bfloat16)          # operand rows [h_{t-1} | x_{t-1}]; zeros = initial state
    gpre = _empty(B, 4 * H, device=dev)
    px = _empty(B, FP, device=dev)
    st = A.stream
    for t in range(Tcap):
        K.gemm_nt(B, 4 * H, H + F, hx, (B, 0, KP), plan.Poff("w1"), H + F, gpre, (B, 0, 4 * H))
        A.call("ag_lstm_step_cell_fwd", K.addr(gpre), K.addr(pre[:, t]), Tcap * 4 * H, K.addr(cbuf[:, t - 1]) if t else None, Tcap * H,
               K.addr(gates[:, t]) if gates is not None else None, Tcap * 4 * H, K.addr(cbuf[:, t]), None, (Tcap + 2) * H,
               K.addr(hbuf16[:, t + 1]), K.addr(hx), KP, B, H, st())
        K.gemm_nt(B, F + 1, H, hx, (B, 0, KP), plan.Poff("w2"), H, px, (B, 0, FP), bias=plan.Poff("b2"))
        A.call("ag_gen_step_proj_finish", K.addr(px), FP, K.addr(xbuf[:, t + 1]), K.addr(xbuf16[:, t + 1]), (Tcap + 1) * F, K.addr(hx), KP, H,
               K.addr(sbuf[:, t]), Tcap, B, F, st())


def _gen_stepwise_bwd(plan, B, T, Tcap, H, F, FP, gates, cbuf, xbuf, dx_ext, ds_ext, dpx, dpx16, dgates16):
    dev = plan.device
    KW = 4 * H + FP
    dgp = torch.zeros(B, KW, device=dev, dtype=torch.bfloat16)         # operand rows [dgates_{t+1} | dpx_t]
    dc = torch.zeros(B, H, device=dev)
    dh = _empty(B, H, device=dev)
    dxpre = _empty(B, F, device=dev)
    st = A.stream
    for t in range(T - 1, -1, -1):
        last = t == T - 1
        if not last:
            K.gemm_nt(B, F, 4 * H, dgp, (B, 0, KW), plan.Poff("wxt"), 4 * H, dxpre, (B, 0, F))
        A.call("ag_gen_step_dpx", None if last else K.addr(dxpre), F, K.addr(dx_ext[:, t]) if dx_ext is not None else None, Tcap * F,
               K.addr(ds_ext[:, t]) if ds_ext is not None else None, Tcap, K.addr(xbuf[:, t + 1]), (Tcap + 1) * F, K.addr(dpx[:, t]),
               K.addr(dpx16[:, t]), Tcap * FP, K.addr(dgp), KW, 4 * H, B, F, FP, st())
        K.gemm_nt(B, H, KW, dgp, (B, 0, KW), plan.Poff("w1t"), KW, dh, (B, 0, H))
        A.call("ag_lstm_step_cell_bwd", K.addr(dh), K.addr(gates[:, t]), Tcap * 4 * H, K.addr(cbuf[:, t]), K.addr(cbuf[:, t - 1]) if t else None,
               Tcap * H, K.addr(dc), None, K.addr(dgates16[:, t]), Tcap * 4 * H, K.addr(dgp), KW, B, H, st())


def _adt(plan):
    return torch.bfloat16 if (plan.mode == "bf16" and _ACT_BF16) else torch.float32


def _zeros(*shape, device, dtype=torch.float32):
    return torch.zeros(*shape, device=device, dtype=dtype)


def _empty(*shape, device, dtype=torch.float32):
    return torch.empty(*shape, device=device, dtype=dtype)


# =========================================================================================
# Generator
# =========================================================================================
class _GenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, struct, token, zc1, u_stop, early_exit_sync, grad_from=0, defer_tail=False):
        """zc1: (B, T, NZ+2) = [noise | conditioning | 1 | 1].  Returns (x (B, t*F), s (B, t), stop (B, t) int32,
        glen (B,) int32).  ``grad_from``: only samples [grad_from, B) take part in the backward pass (train.core_step runs the
        D-update's detached generator pass and the G-update's pass as ONE forward; the first half is never differentiated)."""
        dev = plan.device
        B, Tcap, _ = zc1.shape
        H, F, NZ, CT, FP = plan.H, plan.F, plan.NZ, plan.CT, plan.FP
        zc1 = zc1.contiguous()
        save = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        # hoisted input projection: pre = [z|c|1|1] . [W_z | b_ih | b_hh]^T      (audiogan.py:425-426, :439-440)
        pre = _empty(B, Tcap, 4 * H, device=dev)
        if _adt(plan) == torch.bfloat16 and B * Tcap * 4 * H * plan.ZP >= K.TC_MIN_MACS:
            # bf16 copy with a 16-byte row pitch: the projection and its weight gradient run on the TMA-fed kernels
            # (K = the padded width: the pad columns are zeros on both sides)
            zld = plan.ZP
            zin = torch.zeros(B, Tcap, zld, device=dev, dtype=torch.bfloat16)
            zin[:, :, :NZ + 2] = zc1
        else:
            zld, zin = NZ + 2, zc1
        K.gemm_nt(B * Tcap, 4 * H, zld, zin, (Tcap, Tcap * zld, zld), plan.Poff("wz"), NZ + 2,
                  pre, (Tcap, Tcap * 4 * H, 4 * H))
        bf = plan.mode == "bf16"
        if u_stop is None:
            # no stop sampling: every frame of every sample is written by the kernel; only the initial state (row 0 = h_{-1} /
            # x_{-1}) and the row past the end have to be zero -- not the whole buffers (76 MB of fills at B = 128)
            hbuf = _empty(B, Tcap + 2, H, device=dev)
            xbuf = _empty(B, Tcap + 1, F, device=dev)
            hbuf16 = xbuf16 = None
            if bf:
                hbuf16 = torch.empty(B, Tcap + 2, H, device=dev, dtype=torch.bfloat16)
                xbuf16 = torch.empty(B, Tcap + 1, F, device=dev, dtype=torch.bfloat16)
            K.zero_pads_multi([(t, 1, Tcap + 1) for t in (hbuf, xbuf, hbuf16, xbuf16) if t is not None])    # one launch
        else:
            hbuf = _zeros(B, Tcap + 2, H, device=dev)
            xbuf = _zeros(B, Tcap + 1, F, device=dev)
            hbuf16 = torch.zeros(B, Tcap + 2, H, device=dev, dtype=torch.bfloat16) if bf else None
            xbuf16 = torch.zeros(B, Tcap + 1, F, device=dev, dtype=torch.bfloat16) if bf else None
        gates = _empty(B, Tcap, 4 * H, device=dev) if save else None
        cbuf = _empty(B, Tcap, H, device=dev) if save else None
        sbuf = _zeros(B, Tcap, device=dev)
        stop = torch.zeros(B, Tcap, device=dev, dtype=torch.int32)
        glen = torch.zeros(B, device=dev, dtype=torch.int32)
        misc = torch.zeros(1024, device=dev, dtype=torch.int32)        # [0:8] barrier, [8] t_end, [16:] per-CTA flags
        u = u_stop.contiguous() if u_stop is not None else None
        # Samples are independent: a batch larger than one launch of the TMEM-resident kernel takes (whole co-resident batch
        # groups: 128 samples for the default net) runs as consecutive launches over row slices of the batch-major buffers
        # instead of dropping to the grid-barrier kernels (with stop sampling the early exit is a cross-batch decision: one launch)
        cap = K.lstm_batch_cap(H, F, False) if (bf and u is None and not (plan.lstm_flags & 1)) else 0
        chunk = cap if (cap and B > cap) else B
        # a hidden size the resident kernel cannot hold (H = 2048): per-frame tensor-core GEMMs over the bf16 weights
        stepwise = (bf and u is None and cap == 0 and not (plan.lstm_flags & 1) and H % 8 == 0 and F % 8 == 0 and H >= 1024 and _STEPWISE)
        if stepwise:
            if cbuf is None:
                cbuf = _empty(B, Tcap, H, device=dev)
            _gen_stepwise_fwd(plan, B, Tcap, H, F, FP, pre, hbuf16, xbuf, xbuf16, gates, cbuf, sbuf)
            glen.fill_(Tcap)
        # exchange workspace of the TMEM-resident recurrence (include/audiogan_b200.h: ag_lstm_desc.ll_ws)
        ll_ws = K.lstm_workspace(chunk, H, F, False, dev) if bf else None
        for b0 in (() if stepwise else range(0, B, chunk)):
            sl = slice(b0, min(B, b0 + chunk))
            cut = lambda t: t[sl] if t is not None else None
            K.lstm_fwd(B=sl.stop - sl.start, T=Tcap, Tcap=Tcap, H=H, ndir=1, F=F, pre=pre[sl], w1=plan.Poff("w1"), w2=plan.Poff("w2"),
                       b2=plan.Poff("b2"), hbuf=hbuf[sl], gates=cut(gates), cbuf=cut(cbuf), xbuf=xbuf[sl], sbuf=sbuf[sl], u=u,
                       stop=stop[sl], glen=glen[sl], t_end=(misc, 8), barrier=misc, prec=plan.lstm_prec if bf else 0,
                       hbuf16=cut(hbuf16), xbuf16=cut(xbuf16), flags=plan.lstm_flags | 2 | (8 if u is None else 0), ll_ws=ll_ws,
                       ll_ws_bytes=ll_ws.numel() if ll_ws is not None else 0)
        plan.last_path["g_fwd"] = "stepwise (per-frame tcgen05 GEMMs: H=%d is too large for the resident kernel)" % H if stepwise else K.lstm_last_path()
        if u is not None and early_exit_sync:
            T = int(misc[8].item())          # the one host sync per generator pass (audiogan.py:459-460)
        else:
            T = Tcap
        L = T * F
        Lp = L + 2 * GPAD
        # frame assembly into channel 0 of the dense buffer (audiogan.py:462-464)
        adt = _adt(plan)
        Xd = _empty(B, Lp, CT, device=dev, dtype=adt)       # CT = padded channel count (slots of 8, plan.py)
        hh = [_empty(B, L // s + 2, hid, device=dev, dtype=adt) for (k, s, hid, out) in struct]
        xout = _empty(B, L, device=dev)

        def conv_stack(b0, b1):
            """frames -> dense buffer -> the four bottleneck blocks -> final conv, for samples [b0, b1) (batch-major buffers)"""
            n = b1 - b0
            Xs, xb = Xd[b0:b1], xbuf[b0:b1]
            # the pad rows of the dense buffer and of the four hidden activations: one launch
            K.zero_pads_multi([(Xs, GPAD, GPAD + L)] + [(hh[li][b0:b1], 1, L // s_ + 1) for li, (_, s_, _, _) in enumerate(struct)])
            # frames -> channel 0 of the waveform slot, its pad channels zeroed in the same pass
            K.frames_to_slot((Xs, GPAD * CT), Lp * CT, CT, plan.coff[0], (xb, F), (Tcap + 1) * F, n, L)
            lenL = plan.const_len(n, L)
            for li, (k, s, hid, out) in enumerate(struct):                       # audiogan.py:278-283, :465-467
                p, pd, Lh = (k - 1) // 2, s // 2, L // s
                cin = plan.cinp[li]                             # padded channel prefix this block reads == slot it writes
                Hh = hh[li][b0:b1]
                # TMA-fed kernel over a 4-D tensor map (channel prefix, row, tap, batch).  Its boxes have 16-byte inner rows, so
                # for wide prefixes the producer-warp kernel is faster (measured: cin 8 / 24 -> 1.9x / 1.5x faster, 56 equal, 88 0.6x)
                # (plan.alay 1).  Prefixes of 32 channels and more use 128-byte boxes (plan.alay 2: one tap's 64-channel group per box).
                if adt == torch.bfloat16 and (cin <= _TMA_CIN or plan.alay[li] == 2):
                    K.gemm_nt(n * Lh, hid, k * cin, (Xs, (GPAD - p) * CT), (Lh, Lp * CT, s * CT, cin, CT),
                              plan.Poff("c%d.wq" % li), plan.Kq[li], (Hh, hid), (Lh, (Lh + 2) * hid, hid),
                              bias=plan.Poff("c%d.b" % li), act=1, a_layout=plan.alay[li])
                else:
                    K.gemm_nt(n * Lh, hid, k * cin, (Xs, (GPAD - p) * CT), (Lh, Lp * CT, s * CT, cin, CT),
                              plan.Poff("c%d.w" % li), k * cin, (Hh, hid), (Lh, (Lh + 2) * hid, hid),
                              bias=plan.Poff("c%d.b" % li), act=1)
                skip = (Xs, (GPAD - pd) * CT + plan.skip_off[li]) if plan.skip_off[li] >= 0 else None
                K.gemm_nt(n * (Lh + 1), s * out, 2 * hid, Hh, (Lh + 1, (Lh + 2) * hid, hid),
                          plan.Poff("d%d.w" % li), 2 * hid, (Xs, (GPAD - pd) * CT + cin), (Lh + 1, Lp * CT, s * CT, out, CT),
                          bias=plan.Poff("d%d.b" % li), bias_mod=out, skip=skip, act=1,
                          mask_len=lenL, mask=(s, 1, -pd))
            K.conv1out_fwd((Xs, (GPAD - 1) * CT), Lp * CT, CT, 3, plan.Poff("f.w"), plan.Poff("f.b"), xout[b0:b1], n, L)

        gf0 = int(grad_from)
        if defer_tail and _OVERLAP and bf and 0 < gf0 < B:
            # train.core_step: samples [grad_from, B) are only read by the G-update -- their conv stack runs in the shadow of
            # the D-update's recurrent kernels (shadow_join in core_step before the G-update touches them)
            conv_stack(0, gf0)
            plan.const_len(B - gf0, L)                           # cached on this stream, not inside the side stream's job
            shadow_submit(dev, lambda: conv_stack(gf0, B))
        else:
            conv_stack(0, B)
        s_out = sbuf[:, :T]
        if save:
            gf = int(grad_from)
            ctx.plan, ctx.struct, ctx.dims, ctx.gf = plan, struct, (B - gf, T, Tcap, L), gf
            cut = (lambda t: t[gf:] if t is not None else None) if gf else (lambda t: t)   # batch-major buffers: a row slice
            ctx.bufs = (cut(zin), cut(hbuf), cut(xbuf), cut(gates), cut(cbuf), cut(Xd), [cut(h) for h in hh], cut(hbuf16), cut(xbuf16))
        ctx.mark_non_differentiable(stop, glen)
        ctx.set_materialize_grads(False)
        return xout, s_out, stop[:, :T], glen

    @staticmethod
    @once_differentiable
    def backward(ctx, gx, gs, _gstop, _glen):
        plan, struct = ctx.plan, ctx.struct
        B, T, Tcap, L = ctx.dims
        gf = ctx.gf
        if gf:
            gx = gx[gf:] if gx is not None else None
            gs = gs[gf:] if gs is not None else None
        zc1, hbuf, xbuf, gates, cbuf, Xd, hh, hbuf16, xbuf16 = ctx.bufs
        bf = hbuf16 is not None
        dev = plan.device
        H, F, NZ, CT, FP = plan.H, plan.F, plan.NZ, plan.CT, plan.FP
        Lp = L + 2 * GPAD
        wgrad = ctx.needs_input_grad[2] and wgrad_enabled()
        dx_ext = None
        overlap, keep = False, []
        if gx is not None:
            gx = gx.contiguous()
            # ---- final conv (audiogan.py:403-407): data gradient into all CT channels, weight gradient
            adt = Xd.dtype
            dXd = _empty(B, Lp, CT, device=dev, dtype=adt)
            K.conv1out_dgrad(gx, plan.Poff("f.w"), (dXd, (GPAD - 1) * CT), Lp * CT, CT, 3, B, L)
            # the conv stack's weight gradients are off the critical path: in bf16 mode they are queued and run on the side stream
            # under the BPTT kernel below (whatever does not fit there finishes after it; joined at the end of this backward)
            overlap = wgrad and bf and _OVERLAP and plan.early_sync is None
            keep = []                                       # operands of queued jobs stay alive until the join

            def wjob(fn):
                if overlap:
                    shadow_submit(dev, fn)
                else:
                    fn()

            if wgrad:
                wjob(lambda: K.conv1out_wgrad(gx, (Xd, (GPAD - 1) * CT), Lp * CT, CT, 3, plan.GPoff("f.w"), B, L))
            dHs = [_empty(B, L // s_ + 2, hid_, device=dev, dtype=adt) for (_, s_, hid_, _) in struct]
            K.zero_pads_multi([(t, 1, t.shape[1] - 1) for t in dHs])            # pad rows of the four hidden gradients: one launch
            for li in range(len(struct) - 1, -1, -1):
                k, s, hid, out = struct[li]
                cin = plan.cinp[li]
                p, pd, Lh, kd = (k - 1) // 2, s // 2, L // s, k - 1
                Hh = hh[li]
                # dyl = d(block output) * lrelu'(output); also feeds the dense skip (audiogan.py:281-283)
                dyl = _empty(B, L + 2 * pd, out, device=dev, dtype=adt)
                slice_off = GPAD * CT + cin
                fuse_b = wgrad and out % 4 == 0 and (out // 4) & (out // 4 - 1) == 0 and out <= 1024      # bias gradient in the same pass
                K.ew_grad(B, L, out, out=dyl, pad=(pd, pd), g1=(dXd, slice_off), g1_str=(Lp * CT, CT, 1),
                          act=(Xd, slice_off), act_str=(Lp * CT, CT),
                          acc=((dXd, GPAD * CT + plan.skip_off[li]) if plan.skip_off[li] >= 0 else None),
                          acc_str=(Lp * CT, CT), colsum=plan.GPoff("d%d.b" % li) if fuse_b else None)
                if wgrad and not fuse_b:
                    K.colsum((dyl, pd * out), (L + 2 * pd) * out, out, B, L, out, plan.GPoff("d%d.b" % li))
                # transposed-conv data gradient = strided conv over dyl, times lrelu'(hidden)
                dH = dHs[li]
                K.gemm_nt(B * Lh, hid, kd * out, dyl, (Lh, (L + 2 * pd) * out, s * out), plan.Poff("d%d.wg" % li), kd * out,
                          (dH, hid), (Lh, (Lh + 2) * hid, hid), dact=(Hh, hid))
                if wgrad:
                    keep.append((dyl, dH))

                    def block_wgrad(li=li, k=k, s=s, hid=hid, out=out, cin=cin, p=p, pd=pd, Lh=Lh, kd=kd, Hh=Hh, dyl=dyl, dH=dH):
                        K.gemm_tn(B * Lh, hid, kd * out, (Hh, hid), (Lh, (Lh + 2) * hid, hid), dyl,
                                  (Lh, (L + 2 * pd) * out, s * out), plan.GPoff("d%d.w" % li), kd * out)
                        if adt == torch.bfloat16:
                            Kq = plan.Kq[li]
                            K.gemm_tn(B * Lh, hid, k * cin, (dH, hid), (Lh, (Lh + 2) * hid, hid), (Xd, (GPAD - p) * CT),
                                      (Lh, Lp * CT, s * CT, cin, CT), plan.GPoff("c%d.wq" % li), Kq + 1, ones_col=True, a_layout=plan.alay[li])
                            K.gather(plan.GPoff("c%d.w" % li), plan.GPoff("c%d.wq" % li), plan.q2c[li])
                        else:
                            K.gemm_tn(B * Lh, hid, k * cin, (dH, hid), (Lh, (Lh + 2) * hid, hid), (Xd, (GPAD - p) * CT),
                                      (Lh, Lp * CT, s * CT, cin, CT), plan.GPoff("c%d.w" % li), k * cin + 1, ones_col=True)

                    wjob(block_wgrad)
                # conv data gradient (3 taps over dH), accumulated into channels [0, cin) of dXd
                cdst = (dXd, (GPAD + s - p) * CT)
                K.gemm_nt(B * Lh, s * cin, 3 * hid, dH, (Lh, (Lh + 2) * hid, hid), plan.Poff("c%d.wg" % li), 3 * hid,
                          cdst, (Lh, Lp * CT, s * CT, cin, CT), skip=cdst)
            dx_ext = _zeros(B, Tcap, F, device=dev)
            K.copy3d(dx_ext, (Tcap * F, 1, 0), (dXd, GPAD * CT), (Lp * CT, CT, 0), B, L, 1)
        ds_ext = None
        if gs is not None:
            ds_ext = _zeros(B, Tcap, device=dev)
            ds_ext[:, :T] = gs
        if wgrad and gx is not None:
            plan.reduce_span("c0.wq", "f.w")          # data-parallel: the conv stack's gradients travel under the BPTT
        # ---- BPTT through the recurrence (audiogan.py:437-444)
        dgates = _empty(B, Tcap, 4 * H, device=dev)
        dpx = _empty(B, Tcap, FP, device=dev)
        if T < Tcap:
            dgates[:, T:].zero_()
            dpx[:, T:].zero_()
        misc = torch.zeros(16, device=dev, dtype=torch.int32)
        dgates16 = torch.empty(B, Tcap, 4 * H, device=dev, dtype=torch.bfloat16) if bf else None
        dpx16 = torch.empty(B, Tcap, FP, device=dev, dtype=torch.bfloat16) if bf else None
        # batch chunks as in the forward pass (64 samples per launch of the TMEM-resident BPTT kernel for the default net)
        cap = K.lstm_batch_cap(H, F, True) if (bf and not (plan.lstm_flags & 1)) else 0
        chunk = cap if (cap and B > cap) else B
        stepwise = (bf and K.lstm_batch_cap(H, F, False) == 0 and not (plan.lstm_flags & 1) and H % 8 == 0 and F % 8 == 0 and H >= 1024
                    and _STEPWISE and T == Tcap)
        if stepwise:
            _gen_stepwise_bwd(plan, B, T, Tcap, H, F, FP, gates, cbuf, xbuf, dx_ext, ds_ext, dpx, dpx16, dgates16)
        # reduce-scatter workspace of the TMEM-resident BPTT kernel (include/audiogan_b200.h: ag_lstm_desc.ll_ws)
        ll_ws = K.lstm_workspace(chunk, H, F, True, dev) if bf else None
        ready = shadow_ready(dev) if bf else None
        for b0 in (() if stepwise else range(0, B, chunk)):
            sl = slice(b0, min(B, b0 + chunk))
            cut = lambda t: t[sl] if t is not None else None
            K.lstm_bwd(B=sl.stop - sl.start, T=T, Tcap=Tcap, H=H, ndir=1, F=F, gates=gates[sl], cbuf=cbuf[sl], xbuf=xbuf[sl],
                       dx_ext=cut(dx_ext), ds_ext=cut(ds_ext), dgates=dgates[sl], dpx=dpx[sl], w1t=plan.Poff("w1t"),
                       wxt=plan.Poff("wxt"), barrier=misc, prec=plan.lstm_prec if bf else 0, dgates16=cut(dgates16),
                       dpx16=cut(dpx16), flags=plan.lstm_flags | 2 | (4 if T == Tcap else 0), ll_ws=ll_ws,
                       ll_ws_bytes=ll_ws.numel() if ll_ws is not None else 0)
            if chunk < B:
                misc.zero_()
            if b0 == 0:
                shadow_launch(dev, ready)                # queued weight gradients run on the SMs the BPTT kernel leaves idle
        plan.last_path["g_bwd"] = "stepwise" if stepwise else K.lstm_last_path()
        # REINFORCE (audiogan.py:900-908): the score-function gradient of the stop logits reaches the stop head's weight and
        # bias ONLY (the reference freezes every other generator parameter for that backward), so it joins column F of dpx
        # after the BPTT has run -- it feeds row F of the [wp; ws] weight-gradient GEMM below and nothing else.
        r = plan.__dict__.pop("stopper_ds", None)
        if r is not None and wgrad:
            K.copy3d((dpx, F), (Tcap * FP, FP, 0), r, (r.stride(0), r.stride(1), 0), B, T, 1, accumulate=True)
            if bf:
                K.copy3d((dpx16, F), (Tcap * FP, FP, 0), r, (r.stride(0), r.stride(1), 0), B, T, 1, accumulate=True)
        # in bf16 mode the batched GEMMs read the kernels' bf16 shadow copies (half the operand traffic)
        dgo, hbo, xbo, dpo = (dgates16, hbuf16, xbuf16, dpx16) if bf else (dgates, hbuf, xbuf, dpx)
        if wgrad:
            yv = (T, Tcap * 4 * H, 4 * H)
            K.gemm_tn(B * T, 4 * H, H, dgo, yv, hbo, (T, (Tcap + 2) * H, H), plan.GPoff("w1"), H + F)
            K.gemm_tn(B * T, 4 * H, F, dgo, yv, xbo, (T, (Tcap + 1) * F, F), plan.GPoff("w1", H), H + F)
            zld = zc1.shape[2]        # fp32 [.., NZ+2] or the bf16 copy with its padded row pitch
            K.gemm_tn(B * T, 4 * H, zld, dgo, yv, zc1, (T, Tcap * zld, zld), plan.GPoff("wz"), plan.ZP)
            K.gemm_tn(B * T, FP, H, dpo, (T, Tcap * FP, FP), (hbo, H), (T, (Tcap + 2) * H, H), plan.GPoff("w2"), H + 1,
                      ones_col=True)
        dzc1 = None
        if ctx.needs_input_grad[3]:
            dzc1_all = _zeros(B + gf, Tcap, NZ + 2, device=dev)
            dzc1 = dzc1_all[gf:]
            K.gemm_nt(B * T, NZ, 4 * H, dgo, (T, Tcap * 4 * H, 4 * H), plan.Poff("wzt"), 4 * H,
                      dzc1, (T, Tcap * (NZ + 2), NZ + 2))
            dzc1 = dzc1_all
        if overlap:
            shadow_join(dev)
            keep.clear()
        gtok = torch.zeros(1, device=dev) if wgrad else None
        return None, None, gtok, dzc1, None, None, None, None


# =========================================================================================
# Discriminator: conv stack
# =========================================================================================
class _DiscCNNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, struct, token, x, lens):
        """x (B, L) waveform; lens: list of per-layer int32 [B] frame counts (device).  Returns the 6
        activations as (B, C_i, T_i) views of zero-padded channel-last buffers."""
        dev = plan.device
        B, L = x.shape
        x = x.contiguous()
        a0 = _empty(B, L + 2 * DPAD, device=dev)
        K.frame_noise(a0, L + 2 * DPAD, DPAD, x, L, None, 0.0, B, L)
        acts, Ts = [a0], [L]
        cin, Tin = 1, L
        outs, Tq = [], L
        for (k, s, cout) in struct:
            Tq = (Tq + s - 1) // s
            outs.append(_empty(B, Tq + 2 * DPAD, cout, device=dev, dtype=_adt(plan)))
        K.zero_pads_multi([(o, DPAD, o.shape[1] - DPAD) for o in outs])         # every layer's pad rows: one launch per 8 layers
        for i, (k, s, cout) in enumerate(struct):
            Tout = (Tin + s - 1) // s
            a = outs[i]
            if cin == 1 and k <= 8 and cout % 4 == 0 and (k - 1) // 2 == DPAD:
                # first layer on the raw waveform: K = k, a stream over the output (direct HBM kernel, fp32 arithmetic)
                K.conv1in_fwd(acts[-1], Tin + 2 * DPAD, plan.Poff("c%d.w" % i), plan.Poff("c%d.b" % i), (a, DPAD * cout),
                              (Tout + 2 * DPAD) * cout, k, s, cout, B, Tout, lens[i])
            else:
                K.gemm_nt(B * Tout, cout, k * cin, acts[-1], (Tout, (Tin + 2 * DPAD) * cin, s * cin),
                          plan.Poff("c%d.w" % i), k * cin, (a, DPAD * cout), (Tout, (Tout + 2 * DPAD) * cout, cout),
                          bias=plan.Poff("c%d.b" % i), act=1, mask_len=lens[i], mask=(1, 0, 0))
            acts.append(a)
            Ts.append(Tout)
            cin, Tin = cout, Tout
        ctx.plan, ctx.struct, ctx.acts, ctx.Ts, ctx.lens = plan, struct, acts, Ts, lens
        ctx.set_materialize_grads(False)
        return tuple(a[:, DPAD:DPAD + t].permute(0, 2, 1) for a, t in zip(acts[1:], Ts[1:]))

    @staticmethod
    @once_differentiable
    def backward(ctx, *gouts):
        plan, struct, acts, Ts, lens = ctx.plan, ctx.struct, ctx.acts, ctx.Ts, ctx.lens
        dev = plan.device
        B = acts[0].shape[0]
        wgrad = ctx.needs_input_grad[2] and wgrad_enabled()
        need_dx = ctx.needs_input_grad[3]
        chans = [1] + [c for (_, _, c) in struct]
        dX = None                   # internal gradient wrt acts[i+1], padded geometry
        fused_dy = False            # dX already is dy of the layer below (LeakyReLU' and mask applied by the GEMM above)
        for i in range(len(struct) - 1, -1, -1):
            k, s, cout = struct[i]
            cin, Tin, Tout = chans[i], Ts[i], Ts[i + 1]
            g = gouts[i]
            if g is None and dX is None:
                continue
            a_out = acts[i + 1]
            ntap = plan.d_ntap[i]
            PL = ntap - 1                                  # left pad of dy: the data-gradient window reaches ntap-1 rows back
            geo = ((Tout + 2 * DPAD) * cout, cout)         # geometry of the forward activation buffers
            gdy = ((PL + Tout + DPAD) * cout, cout)        # geometry of dy
            adt = a_out.dtype
            if g is None and fused_dy:
                # the data-gradient GEMM of the layer above already applied LeakyReLU'(a_out) and this layer's length mask in
                # its epilogue and wrote into the dy geometry (PL == DPAD): no activation-gradient pass
                dy = dX
            else:
                dy = _empty(B, PL + Tout + DPAD, cout, device=dev, dtype=adt)
                kw = {}
                if g is not None:                  # (B, C, T) tensor with arbitrary strides
                    kw.update(g1=g, g1_str=(g.stride(0), g.stride(2), g.stride(1)))
                if dX is not None:
                    kw.update(g2=(dX, DPAD * cout), g2_str=(geo[0], geo[1], 1))
                K.ew_grad(B, Tout, cout, out=dy, pad=(PL, DPAD), act=(a_out, DPAD * cout), act_str=geo, length=lens[i], **kw)
            fused_dy = False
            a_view = (Tout, (Tin + 2 * DPAD) * cin, s * cin)
            if wgrad and cin == 1 and k <= 8 and cout % 4 == 0 and (k - 1) // 2 == DPAD:
                K.conv1in_wgrad((dy, PL * cout), gdy[0], acts[i], Tin + 2 * DPAD, plan.GPoff("c%d.w" % i), k, s, cout, B, Tout)
            elif wgrad:
                K.gemm_tn(B * Tout, cout, k * cin, (dy, PL * cout), (Tout, gdy[0], cout), acts[i], a_view,
                          plan.GPoff("c%d.w" % i), k * cin + 1, ones_col=True)
            if i == 0 and need_dx and cin == 1 and cout % 4 == 0 and (k - 1) // 2 == DPAD:
                # gradient of the raw waveform: a GEMM with N = s columns -> direct kernel (one thread per sample)
                dX = _empty(B, Tin + 2 * DPAD, 1, device=dev)
                K.conv1in_dgrad((dy, PL * cout), gdy[0], plan.Poff("c%d.w" % i), dX, Tin + 2 * DPAD, k, s, DPAD, cout, B, Tout, Tin)
            elif i > 0 or need_dx:
                Mp = (Tin + DPAD + s - 1) // s
                # the gradient of the raw waveform (layer 0) is a caller-visible fp32 tensor
                dXn = _empty(B, Tin + 2 * DPAD, cin, device=dev, dtype=adt if i > 0 else torch.float32)
                if s * Mp < Tin + 2 * DPAD:
                    dXn[:, s * Mp:].zero_()
                # When nothing else flows into the activation below (no external gradient for cnn_outputs[i-1]) and its dy buffer
                # has this geometry, fold its LeakyReLU' and length mask into this epilogue.
                fuse = (i > 0 and gouts[i - 1] is None and plan.d_ntap[i - 1] - 1 == DPAD and adt == torch.bfloat16)
                if fuse:
                    K.gemm_nt(B * Mp, s * cin, ntap * cout, dy, (Mp, gdy[0], cout), plan.Poff("c%d.wg" % i), ntap * cout,
                              dXn, (Mp, (Tin + 2 * DPAD) * cin, s * cin, cin, cin), dact=acts[i],
                              mask_len=lens[i - 1], mask=(s, 1, -DPAD))
                else:
                    K.gemm_nt(B * Mp, s * cin, ntap * cout, dy, (Mp, gdy[0], cout), plan.Poff("c%d.wg" % i), ntap * cout,
                              dXn, (Mp, (Tin + 2 * DPAD) * cin, s * cin, cin, cin),
                              mask_len=plan.const_len(B, Tin), mask=(s, 1, -DPAD))
                fused_dy = fuse
                dX = dXn
            else:
                dX = None
        gx = dX[:, DPAD:DPAD + Ts[0], 0] if (need_dx and dX is not None) else None
        gtok = torch.zeros(1, device=dev) if wgrad else None
        return None, None, gtok, gx, None


# =========================================================================================
# Discriminator: BiLSTM + residual net + classifier
# =========================================================================================
class _DiscTailFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, token, feat, c, nfr, Tm, Tmin=0):
        """feat: (B, Cf, T6) channel-last view; c (B, E); nfr int32 [B] frames per sample; Tm = max(nfr), Tmin = min(nfr)
        (host values: rows 1..Tmin of the recurrent state are written by the kernel for every sample, so only row 0 and
        rows > Tmin are zeroed -- packed-sequence semantics, audiogan.py:214-229: outputs past a sample's length are zero)."""
        dev = plan.device
        B, Cf, T6 = feat.shape
        if feat.stride(1) != 1:
            feat = feat.permute(0, 2, 1).contiguous().permute(0, 2, 1)
        S, H, E = plan.S, plan.H, plan.E
        ldi = Cf + E + 2
        c1 = torch.cat([c, torch.ones(B, 2, device=dev)], 1).contiguous()
        rb = _empty(B, 8 * H, device=dev)
        K.gemm_nt(B, 8 * H, E + 2, c1, (B, 0, E + 2), plan.Poff("wih", Cf), ldi,
                  rb, (B, 0, 8 * H))
        pre = _empty(B, Tm, 8 * H, device=dev)
        K.gemm_nt(B * Tm, 8 * H, Cf, feat, (Tm, feat.stride(0), feat.stride(2)), plan.Poff("wih"), ldi,
                  pre, (Tm, Tm * 8 * H, 8 * H), rowbias=rb, rowbias_ld=8 * H)
        bf = plan.mode == "bf16"
        Tmin = max(0, min(int(Tmin), Tm))
        hbuf = _empty(B, Tm + 2, 2 * H, device=dev)
        hbuf16 = torch.empty(B, Tm + 2, 2 * H, device=dev, dtype=torch.bfloat16) if bf else None
        K.zero_pads_multi([(t, 1, Tmin + 1) for t in (hbuf, hbuf16) if t is not None])
        gates = _empty(B, Tm, 8 * H, device=dev)
        cbuf = _empty(B, Tm, 2 * H, device=dev)
        misc = torch.zeros(16, device=dev, dtype=torch.int32)
        ready = shadow_ready(dev) if bf else None
        K.lstm_fwd(B=B, T=Tm, Tcap=Tm, H=H, ndir=2, F=0, pre=pre, w1=plan.Poff("w1"), hbuf=hbuf, gates=gates, cbuf=cbuf,
                   len=nfr, barrier=misc, prec=plan.lstm_prec if bf else 0, hbuf16=hbuf16,
                   flags=plan.lstm_flags | (8 if bf else 0))       # bf16 mode: h is consumed as bf16 only (fallback kernels ignore the bit)
        shadow_launch(dev, ready)
        plan.last_path["d_fwd"] = K.lstm_last_path()
        # residual_net + classifier on the (B*Tm) rows (audiogan.py:547-549); same row geometry as hbuf
        geo = (Tm, (Tm + 2) * S, S)
        # bf16 mode: the tail's activations live in HBM as bf16 (they only feed tensor-core GEMMs, whose throughput is
        # set by operand bytes per FLOP: profiles/r1_gemm_notes.txt); fp32 accumulation and fp32 logits
        adt = torch.bfloat16 if bf else torch.float32
        hin = hbuf16 if bf else hbuf
        r1 = torch.empty(B, Tm + 2, S, device=dev, dtype=adt)
        r2 = torch.empty(B, Tm + 2, S, device=dev, dtype=adt)
        K.gemm_nt(B * Tm, S, S, (hin, S), geo, plan.Poff("r0.w"), S, (r1, S), geo, bias=plan.Poff("r0.b"),
                  skip=(hin, S), act=1)
        K.gemm_nt(B * Tm, S, S, (r1, S), geo, plan.Poff("r1.w"), S, (r2, S), geo, bias=plan.Poff("r1.b"),
                  skip=(r1, S), act=1)
        h3 = torch.empty(B * Tm, S // 2, device=dev, dtype=adt)
        K.gemm_nt(B * Tm, S // 2, S, (r2, S), geo, plan.Poff("k0.w"), S, h3, (B * Tm, 0, S // 2), bias=plan.Poff("k0.b"), act=1)
        logits = _empty(B, Tm, device=dev)
        if bf and (S // 2) % 4 == 0 and S // 2 <= 1024:
            # Linear(S/2 -> 1): one warp per row over the packed activation (a GEMM with one output column wastes the tile)
            K.rowdot(h3, plan.Poff("k2.w"), plan.Poff("k2.b"), logits, B * Tm, S // 2)
        else:
            K.gemm_nt(B * Tm, 1, S // 2, h3, (B * Tm, 0, S // 2), plan.Poff("k2.w"), S // 2, logits, (B * Tm, 0, 1),
                      bias=plan.Poff("k2.b"))
        ctx.plan, ctx.dims = plan, (B, Cf, T6, Tm)
        ctx.bufs = (feat, c1, nfr, hbuf, gates, cbuf, r1, r2, h3, hbuf16)
        return logits

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        plan = ctx.plan
        B, Cf, T6, Tm = ctx.dims
        feat, c1, nfr, hbuf, gates, cbuf, r1, r2, h3, hbuf16 = ctx.bufs
        bf = hbuf16 is not None
        dev = plan.device
        S, H, E = plan.S, plan.H, plan.E
        ldi = Cf + E + 2
        wgrad = ctx.needs_input_grad[1] and wgrad_enabled()
        M = B * Tm
        g = g.contiguous()
        geo = (Tm, (Tm + 2) * S, S)
        flat = lambda n: (M, 0, n)
        # The tail's weight gradients are off the critical path (nothing downstream reads them before _PackFn.backward), and the
        # BPTT kernel that follows is latency-bound on 64-96 of the 148 SMs: in bf16 mode they are issued on a side stream right
        # AFTER the BPTT launch and run on the idle SMs underneath it (AUDIOGAN_OVERLAP=0: in line).  Not with the early
        # all-reduce (GradSync.attach), which wants these gradients final before the BPTT starts.
        overlap = wgrad and bf and _OVERLAP and plan.early_sync is None

        def wjob(fn):
            if overlap:
                shadow_submit(dev, fn)
            else:
                fn()

        if wgrad and (S // 2) % 4 == 0 and S // 2 <= 1024:
            wjob(lambda: K.wcolsum(g, h3, M, S // 2, plan.GPoff("k2.w")))   # Linear(S/2 -> 1): weighted column sum, not a GEMM
        elif wgrad:
            wjob(lambda: K.gemm_tn(M, 1, S // 2, g, flat(1), h3, flat(S // 2), plan.GPoff("k2.w"), S // 2 + 1, ones_col=True))
        adt = torch.bfloat16 if bf else torch.float32
        hin = hbuf16 if bf else hbuf
        dh3 = torch.empty(M, S // 2, device=dev, dtype=adt)
        K.outer_dact(g, plan.Poff("k2.w"), h3, dh3, M, S // 2)          # rank-1: dh3 = g (x) k2.w * lrelu'(h3)
        if wgrad:
            wjob(lambda: K.gemm_tn(M, S // 2, S, dh3, flat(S // 2), (r2, S), geo, plan.GPoff("k0.w"), S + 1, ones_col=True))
        dz2 = torch.empty(B, Tm + 2, S, device=dev, dtype=adt)   # grads below keep the padded-row geometry of r1 / r2 / hbuf
        K.gemm_nt(M, S, S // 2, dh3, flat(S // 2), plan.Poff("k0.wt"), S // 2, (dz2, S), geo, dact=(r2, S))
        if wgrad:
            wjob(lambda: K.gemm_tn(M, S, S, (dz2, S), geo, (r1, S), geo, plan.GPoff("r1.w"), S + 1, ones_col=True))
        dz1 = torch.empty(B, Tm + 2, S, device=dev, dtype=adt)
        K.gemm_nt(M, S, S, (dz2, S), geo, plan.Poff("r1.wt"), S, (dz1, S), geo, skip=(dz2, S), dact=(r1, S))
        if wgrad:
            wjob(lambda: K.gemm_tn(M, S, S, (dz1, S), geo, (hin, S), geo, plan.GPoff("r0.w"), S + 1, ones_col=True))
        if wgrad:
            plan.reduce_span("r0.w", "k2.w")          # data-parallel: these gradients are final, reduce them under the BPTT
        dh_ext = _empty(B, Tm + 2, S, device=dev)
        K.gemm_nt(M, S, S, (dz1, S), geo, plan.Poff("r0.wt"), S, (dh_ext, S), geo, skip=(dz1, S))
        # BPTT through both directions
        dgates = _empty(B, Tm, 8 * H, device=dev)
        misc = torch.zeros(16, device=dev, dtype=torch.int32)
        dgates16 = torch.empty(B, Tm, 8 * H, device=dev, dtype=torch.bfloat16) if bf else None
        ready = shadow_ready(dev)                               # the queued jobs' operands are complete here
        K.lstm_bwd(B=B, T=Tm, Tcap=Tm, H=H, ndir=2, F=0, gates=gates, cbuf=cbuf, len=nfr, dh_ext=(dh_ext, S),
                   dh_ext_bs=(Tm + 2) * S, dgates=dgates, w1t=plan.Poff("w1t"), barrier=misc,
                   prec=plan.lstm_prec if bf else 0, dgates16=dgates16, flags=plan.lstm_flags | 4)
        shadow_launch(dev, ready)
        plan.last_path["d_bwd"] = K.lstm_last_path()
        dgo, hbo = (dgates16, hbuf16) if bf else (dgates, hbuf)
        dgsum = None
        if wgrad or ctx.needs_input_grad[3]:
            dgsum = _empty(B, 8 * H, device=dev)
            K.rowgroup_sum(dgo, dgsum, B, Tm, 8 * H)          # bf16 mode: from the bf16 shadow (half the bytes)
        if wgrad:
            for d in range(2):
                K.gemm_tn(M, 4 * H, H, (dgo, d * 4 * H), (Tm, Tm * 8 * H, 8 * H),
                          (hbo, (2 * d) * 2 * H + d * H), (Tm, (Tm + 2) * 2 * H, 2 * H),
                          plan.GPoff("whh", d * 4 * H * H), H)
            K.gemm_tn(M, 8 * H, Cf, dgo, flat(8 * H), feat, (Tm, feat.stride(0), feat.stride(2)), plan.GPoff("wih"), ldi)
            K.gemm_tn(B, 8 * H, E + 2, dgsum, (B, 0, 8 * H), c1, (B, 0, E + 2), plan.GPoff("wih", Cf), ldi)
            plan.reduce_span("wih", "whh")            # ... and the recurrent weights' under the conv stack's backward
        dfeat = None
        if ctx.needs_input_grad[2]:
            dfeat = _zeros(B, T6, Cf, device=dev, dtype=feat.dtype) if Tm < T6 else _empty(B, T6, Cf, device=dev, dtype=feat.dtype)
            K.gemm_nt(M, Cf, 8 * H, dgo, flat(8 * H), plan.Poff("wiht"), 8 * H, dfeat, (Tm, T6 * Cf, Cf))
            dfeat = dfeat.permute(0, 2, 1)
        dc = None
        if ctx.needs_input_grad[3]:
            dc = _empty(B, E, device=dev)
            K.gemm_nt(B, E, 8 * H, dgsum, (B, 0, 8 * H), plan.Poff("wiht", Cf * 8 * H), 8 * H,
                      dc, (B, 0, E))
        if overlap:
            shadow_join(dev)                                      # the side stream's weight gradients join before anything reads them
        gtok = torch.zeros(1, device=dev) if wgrad else None
        return None, gtok, dfeat, dc, None, None, None


# =========================================================================================
# Embedder: dynamic_rnn + BiLSTM over the character sequence, last hidden states (audiogan.py:214-229, :325-334)
# =========================================================================================
class _EmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, token, x, nfr, len_long):
        """x (B, T, E) embedded characters; nfr int32 [B] / len_long int64 [B] lengths on the device.  Returns (B, 2H):
        [h of the forward direction at its last valid step | h of the reverse direction at step 0] (audiogan.py:333-334)."""
        dev = plan.device
        B, Tm, E = x.shape
        H, HP = plan.H, plan.HP
        x1 = torch.cat([x, torch.ones(B, Tm, 2, device=dev)], 2).contiguous()          # [x | 1 | 1]: both biases ride along
        pre = _empty(B, Tm, 8 * HP, device=dev)
        K.gemm_nt(B * Tm, 8 * HP, E + 2, x1, (Tm, Tm * (E + 2), E + 2), plan.Poff("wih"), E + 2, pre, (Tm, Tm * 8 * HP, 8 * HP))
        hbuf = _zeros(B, Tm + 2, 2 * HP, device=dev)
        gates = _empty(B, Tm, 8 * HP, device=dev)
        cbuf = _empty(B, Tm, 2 * HP, device=dev)
        misc = torch.zeros(16, device=dev, dtype=torch.int32)
        K.lstm_fwd(B=B, T=Tm, Tcap=Tm, H=HP, ndir=2, F=0, pre=pre, w1=plan.Poff("w1"), hbuf=hbuf, gates=gates, cbuf=cbuf,
                   len=nfr, barrier=misc, prec=0, flags=1)
        ar = torch.arange(B, device=dev)
        out = torch.cat([hbuf[ar, len_long, :H], hbuf[:, 1, HP:HP + H]], 1)            # row t+1 = h_t
        ctx.plan, ctx.bufs = plan, (x1, nfr, len_long, hbuf, gates, cbuf)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        plan = ctx.plan
        x1, nfr, len_long, hbuf, gates, cbuf = ctx.bufs
        dev = plan.device
        B, Tm, E2 = x1.shape
        E, H, HP = E2 - 2, plan.H, plan.HP
        M = B * Tm
        wgrad = ctx.needs_input_grad[1] and wgrad_enabled()
        dh_ext = _zeros(B, Tm + 2, 2 * HP, device=dev)
        ar = torch.arange(B, device=dev)
        dh_ext[ar, len_long, :H] = g[:, :H]
        dh_ext[:, 1, HP:HP + H] = g[:, H:]
        dgates = _empty(B, Tm, 8 * HP, device=dev)
        misc = torch.zeros(16, device=dev, dtype=torch.int32)
        K.lstm_bwd(B=B, T=Tm, Tcap=Tm, H=HP, ndir=2, F=0, gates=gates, cbuf=cbuf, len=nfr, dh_ext=(dh_ext, 2 * HP),
                   dh_ext_bs=(Tm + 2) * 2 * HP, dgates=dgates, w1t=plan.Poff("w1t"), barrier=misc, prec=0, flags=1)
        flat = lambda n: (M, 0, n)
        if wgrad:
            for d in range(2):
                K.gemm_tn(M, 4 * HP, HP, (dgates, d * 4 * HP), (Tm, Tm * 8 * HP, 8 * HP),
                          (hbuf, (2 * d) * 2 * HP + d * HP), (Tm, (Tm + 2) * 2 * HP, 2 * HP),
                          plan.GPoff("whh", d * 4 * HP * HP), HP)
            K.gemm_tn(M, 8 * HP, E + 2, dgates, flat(8 * HP), x1, flat(E + 2), plan.GPoff("wih"), E + 2)
        dx = None
        if ctx.needs_input_grad[2]:
            dx = _empty(B, Tm, E, device=dev)
            K.gemm_nt(M, E, 8 * HP, dgates, flat(8 * HP), plan.Poff("wiht"), 8 * HP, dx, flat(E))
        gtok = torch.zeros(1, device=dev) if wgrad else None
        return None, gtok, dx, None, None


# =========================================================================================
# calc_dists: per-(sample, channel) time moments of a conv activation (audiogan.py:341-348)
# =========================================================================================
class _TimeMomentsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, nfr):
        """h (B, C, T): a cnn_outputs entry of Discriminator.forward (a permuted view of the kernels' channel-last buffer; any
        other layout is re-laid-out once); nfr int32 [B] on the device.  Returns q (3, B, C) = (m, s, f)."""
        B, Cn, Tn = h.shape
        if h.stride(1) != 1 or h.dtype not in (torch.float32, torch.bfloat16):
            h = h.float().permute(0, 2, 1).contiguous().permute(0, 2, 1)
        dev = h.device
        S1 = torch.empty(B, Cn, device=dev)
        Q = torch.empty(3, B, Cn, device=dev)
        K.time_moments_fwd(h, h.stride(0), h.stride(2), nfr, B, Tn, Cn, S1, Q)
        lf = nfr.to(torch.float32).unsqueeze(1)
        q = torch.stack([S1 / lf, Q[0].sqrt() / lf, Q[2].pow(0.25) / lf], 0)
        ctx.save_for_backward(h, nfr, S1, Q)
        return q

    @staticmethod
    @once_differentiable
    def backward(ctx, gq):
        h, nfr, S1, Q = ctx.saved_tensors
        B, Cn, Tn = h.shape
        gq = gq.contiguous().float()
        dh = torch.empty(B, Tn, Cn, device=h.device, dtype=h.dtype)
        K.time_moments_bwd(h, h.stride(0), h.stride(2), nfr, B, Tn, Cn, S1, Q, gq[0], gq[1], gq[2], dh)
        return dh.permute(0, 2, 1), None


# =========================================================================================
# BCE with logits per sample (audiogan.py:187-197)
# =========================================================================================
class _BCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target, weight):
        B, T = x.shape
        x, target = x.contiguous(), target.contiguous()
        w = weight.contiguous() if weight is not None else None
        loss = torch.empty(B, device=x.device, dtype=torch.float32)
        K.bce_fwd(x, target, w, loss, B, T)
        ctx.save_for_backward(x, target, w)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        x, target, w = ctx.saved_tensors
        B, T = x.shape
        dx = torch.empty_like(x)
        K.bce_bwd(x, target, w, gout.contiguous(), dx, B, T)
        return dx, None, None
