"""Thin tensor-level wrappers over the C ABI (one Python function per exported kernel).

Operands are torch CUDA tensors used purely as device memory; an operand may be given as
``(tensor, element_offset)``.  No arithmetic happens here: every function fills a descriptor
and launches on torch's current stream.
"""
import ctypes as C

import torch

from . import _abi as A

LRELU_SLOPE = 0.01      # NN.LeakyReLU() / F.leaky_relu defaults (audiogan.py:261, :277, :532)
_DT = {torch.float32: 0, torch.bfloat16: 1}


def addr(x):
    """x: None | tensor | (tensor, element offset) -> integer device address (None -> 0)."""
    if x is None:
        return None
    if isinstance(x, tuple):
        t, off = x
        return t.data_ptr() + off * t.element_size()
    return x.data_ptr()


def _dtype_of(x):
    if x is None:
        return 0
    t = x[0] if isinstance(x, tuple) else x
    return _DT[t.dtype]


def gemm_desc(M, N, K, A_, a_view, B_, ldb, C_, c_view, alpha=0.0, bias=None, bias_mod=0, rowbias=None,
              rowbias_ld=0, skip=None, act=0, dact=None, slope=LRELU_SLOPE, mask_len=None, mask=(0, 0, 0), a_layout=0):
    """a_view = (rows_per_batch, batch_stride, row_stride[, k_inner, k_outer_stride]);
    c_view = (rows_per_batch, batch_stride, row_stride[, n_inner, n_outer_stride]).
    (One positional constructor call in field order: ~90 descriptors are built per training step.)"""
    a = a_view if len(a_view) == 5 else (a_view[0], a_view[1], a_view[2], K, 0)
    c = c_view if len(c_view) == 5 else (c_view[0], c_view[1], c_view[2], N, 0)
    return A.GemmDesc(M, N, K,
                      addr(A_), a[0], a[1], a[2], a[3], a[4],
                      addr(B_), ldb,
                      addr(C_), c[0], c[1], c[2], c[3], c[4],
                      alpha, addr(bias), bias_mod, addr(rowbias), rowbias_ld, addr(skip), act, addr(dact), slope,
                      addr(mask_len), mask[0], mask[1], mask[2],
                      _dtype_of(A_), _dtype_of(B_), _dtype_of(C_), _dtype_of(skip if skip is not None else dact), a_layout)


TC_MIN_MACS = 1 << 20      # below this a GEMM stays on the fp32 FFMA kernel even in bf16 mode (launch-bound anyway)


def _plan_mode(ref):
    pl = getattr(ref, "plan", None)
    return pl.mode if pl is not None else "fp32"


def gemm_nt(M, N, K, A_, a_view, B_, ldb, C_, c_view, tc=False, **kw):
    """C = epilogue(A . B^T).  B_ given as plan.Poff(...) runs on the tcgen05 kernel with the plan's bf16 weight copy
    when the plan is in bf16 mode; tc=True forces the tensor-core kernel on explicit bf16 B operands."""
    if not tc and _plan_mode(B_) == "bf16" and M * N * K >= TC_MIN_MACS:
        sub = B_.bf16(ldb)
        if sub is not None:
            B_, ldb = sub
            tc = True
    d = gemm_desc(M, N, K, A_, a_view, B_, ldb, C_, c_view, **kw)
    A.call("ag_gemm_nt_tc" if tc else "ag_gemm_nt_f32", C.byref(d), A.stream())


def gemm_tn(M, N, K, Y, y_view, A_, a_view, dw, ldw, ones_col=False, tc=False, a_layout=0):
    """dw[n, k] += sum_m Y(m, n) * A(m, k); optional bias column K."""
    if not tc and _plan_mode(dw) == "bf16" and M * N * K >= TC_MIN_MACS:
        tc = True
    d = gemm_desc(M, N, K, A_, a_view, None, 0, Y, y_view, a_layout=a_layout)
    A.call("ag_gemm_tn_tc" if tc else "ag_gemm_tn_f32", C.byref(d), addr(dw), ldw, 1 if ones_col else 0, A.stream())


def lstm_desc(**kw):
    d = A.LstmDesc()
    for k, v in kw.items():
        if isinstance(v, (torch.Tensor, tuple)):
            v = addr(v)
        setattr(d, k, v)
    return d


def lstm_fwd(**kw):
    d = lstm_desc(**kw)
    A.call("ag_lstm_fwd", C.byref(d), A.stream())


def lstm_bwd(**kw):
    d = lstm_desc(**kw)
    A.call("ag_lstm_bwd", C.byref(d), A.stream())


def lstm_workspace(B, H, F, bwd, device):
    """`ll_ws` for the TMEM-resident generator recurrence, sized by the library (ag_lstm_workspace_bytes); None when the shape
    does not run there (the call then uses a kernel family that needs no workspace)."""
    d = A.LstmDesc()
    d.B, d.H, d.F, d.ndir, d.prec = B, H, F, 1, 1
    n = A.lib().ag_lstm_workspace_bytes(C.byref(d), 1 if bwd else 0)
    if n < 0:
        raise A.AudioganError("ag_lstm_workspace_bytes failed (%d)" % n)
    return torch.empty(n, device=device, dtype=torch.uint8) if n > 0 else None


_batch_cap = {}


def lstm_batch_cap(H, F, bwd):
    """samples one launch of the TMEM-resident generator recurrence takes (0: the shape does not run there)"""
    key = (H, F, bool(bwd), torch.cuda.current_device())
    if key not in _batch_cap:
        d = A.LstmDesc()
        d.B, d.H, d.F, d.ndir, d.prec = 1, H, F, 1, 1
        _batch_cap[key] = int(A.lib().ag_lstm_batch_cap(C.byref(d), 1 if bwd else 0))
    return _batch_cap[key]


def lstm_last_path():
    """Kernel family (and decline reason, if any) of the last lstm_fwd / lstm_bwd call on this thread."""
    return A.lib().ag_lstm_last_path().decode("utf-8", "replace")


def gather(dst, src, idx):
    A.call("ag_gather", addr(dst), addr(src), addr(idx), idx.numel(), _dtype_of(dst), A.stream())


def frame_noise(dst, dst_ld, pad_l, src, src_ld, noise, noise_scale, B, L):
    A.call("ag_frame_noise", addr(dst), dst_ld, pad_l, addr(src), src_ld, addr(noise), float(noise_scale), B, L,
           _dtype_of(dst), A.stream())


def ew_grad(B, T, Cn, out=None, pad=(0, 0), g1=None, g1_str=(0, 0, 0), g2=None, g2_str=(0, 0, 0), act=None,
            act_str=(0, 0), length=None, acc=None, acc_str=(0, 0), slope=LRELU_SLOPE, colsum=None):
    d = A.EwDesc()
    d.B, d.T, d.C = B, T, Cn
    d.g1 = addr(g1)
    d.g1_bs, d.g1_rs, d.g1_cs = g1_str
    d.g2 = addr(g2)
    d.g2_bs, d.g2_rs, d.g2_cs = g2_str
    d.act = addr(act)
    d.a_bs, d.a_rs = act_str
    d.slope = slope
    d.len = addr(length)
    d.out = addr(out)
    d.pad_l, d.pad_r = pad
    d.acc = addr(acc)
    d.acc_bs, d.acc_rs = acc_str
    d.g1_dtype, d.g2_dtype, d.act_dtype = _dtype_of(g1), _dtype_of(g2), _dtype_of(act)
    d.acc_dtype, d.out_dtype = _dtype_of(acc), _dtype_of(out)
    d.colsum = addr(colsum)
    A.call("ag_ew_grad", C.byref(d), A.stream())


def conv1in_fwd(x, x_ld, w, bias, out, out_bs, k, s, Cn, B, T, length, slope=LRELU_SLOPE):
    A.call("ag_conv1in_fwd", addr(x), x_ld, addr(w), addr(bias), addr(out), _dtype_of(out), out_bs, k, s, Cn, B, T, addr(length), float(slope),
           A.stream())


def conv1in_wgrad(dy, dy_bs, x, x_ld, dw, k, s, Cn, B, T):
    A.call("ag_conv1in_wgrad", addr(dy), _dtype_of(dy), dy_bs, addr(x), x_ld, addr(dw), k, s, Cn, B, T, A.stream())


def conv1in_dgrad(dy, dy_bs, w, dx, dx_ld, k, s, p, Cn, B, T, Tin):
    A.call("ag_conv1in_dgrad", addr(dy), _dtype_of(dy), dy_bs, addr(w), addr(dx), dx_ld, k, s, p, Cn, B, T, Tin, A.stream())


def wcolsum(g, X, M, Kn, out):
    """out[k] += sum_m g[m] X[m, k], out[Kn] += sum_m g[m] (X packed [M, Kn], fp32 or bf16)."""
    A.call("ag_wcolsum", addr(g), addr(X), _dtype_of(X), M, Kn, addr(out), A.stream())


def rowdot(X, w, bias, out, M, Kn):
    """out[m] = sum_k X[m, k] * w[k] + bias[0] (Linear(K -> 1) forward)"""
    A.call("ag_rowdot", addr(X), _dtype_of(X), addr(w), addr(bias), M, Kn, addr(out), A.stream())


def outer_dact(g, w, act, out, M, N, slope=LRELU_SLOPE):
    """out[m, n] = g[m] * w[n] * lrelu'(act[m, n]) (packed [M, N]; act / out fp32 or bf16)."""
    A.call("ag_outer_dact", addr(g), addr(w), addr(act), _dtype_of(act), addr(out), _dtype_of(out), M, N, float(slope), A.stream())


def colsum(src, bs, rs, B, T, Cn, out):
    A.call("ag_colsum", addr(src), _dtype_of(src), bs, rs, B, T, Cn, addr(out), A.stream())


def copy3d(dst, d_str, src, s_str, B, T, Cn, accumulate=False):
    A.call("ag_copy3d", addr(dst), d_str[0], d_str[1], d_str[2], addr(src), s_str[0], s_str[1], s_str[2], B, T, Cn,
           1 if accumulate else 0, _dtype_of(src), _dtype_of(dst), A.stream())


def conv1out_fwd(X, x_bs, Cn, k, w, bias, out, B, T):
    A.call("ag_conv1out_fwd", addr(X), _dtype_of(X), x_bs, Cn, k, addr(w), addr(bias), addr(out), B, T, A.stream())


def conv1out_dgrad(g, w, dX, dx_bs, Cn, k, B, T):
    A.call("ag_conv1out_dgrad", addr(g), addr(w), addr(dX), _dtype_of(dX), dx_bs, Cn, k, B, T, A.stream())


def conv1out_wgrad(g, X, x_bs, Cn, k, dw, B, T):
    A.call("ag_conv1out_wgrad", addr(g), addr(X), _dtype_of(X), x_bs, Cn, k, addr(dw), B, T, A.stream())


def frames_to_slot(dst, d_bs, d_rs, slot, src, s_bs, B, L):
    A.call("ag_frames_to_slot", addr(dst), _dtype_of(dst), d_bs, d_rs, slot, addr(src), s_bs, B, L, A.stream())


def zero_pads(buf, head, tail0):
    """zero rows [0, head) and [tail0, rows) of every batch of a contiguous [B, rows, C] tensor."""
    Bn, rows = buf.shape[0], buf.shape[1]
    rb = buf[0, 0].numel() * buf.element_size()
    if not buf.is_contiguous():
        raise ValueError("zero_pads needs a contiguous [B, rows, ...] buffer")
    A.call("ag_zero_pads", addr(buf), Bn, rows, rb, head, tail0, A.stream())


def zero_pads_multi(items):
    """items: [(buf, head, tail0), ...] as for zero_pads -- one launch per 8 buffers."""
    for i0 in range(0, len(items), 8):
        chunk = items[i0:i0 + 8]
        arr = (A.PadEntry * len(chunk))()
        for e, (buf, head, tail0) in zip(arr, chunk):
            if not buf.is_contiguous():
                raise ValueError("zero_pads needs a contiguous [B, rows, ...] buffer")
            e.buf, e.B, e.rows = buf.data_ptr(), buf.shape[0], buf.shape[1]
            e.row_bytes, e.head, e.tail0 = buf[0, 0].numel() * buf.element_size(), head, tail0
        A.call("ag_zero_pads_multi", C.cast(arr, C.c_void_p), len(chunk), A.stream())


def rowgroup_sum(src, out, B, T, N):
    A.call("ag_rowgroup_sum", addr(src), _dtype_of(src), addr(out), B, T, N, A.stream())


def bce_fwd(x, tgt, w, loss, B, T):
    A.call("ag_bce_fwd", addr(x), addr(tgt), addr(w), addr(loss), B, T, A.stream())


def bce_bwd(x, tgt, w, gout, dx, B, T):
    A.call("ag_bce_bwd", addr(x), addr(tgt), addr(w), addr(gout), addr(dx), B, T, A.stream())


def bce_const_fused(x, ld, length, target, sign, loss_mean, loss_ps, dlogits, stats, B, T):
    A.call("ag_bce_const_fused", addr(x), ld, addr(length), float(target), float(sign), addr(loss_mean), addr(loss_ps),
           addr(dlogits), addr(stats), B, T, A.stream())


def reinforce_dlogit(s, s_ld, stop, stop_ld, loss_ps, glen, baseline_in, baseline_out, out, out_ld, B, T):
    A.call("ag_reinforce_dlogit", addr(s), s_ld, addr(stop), stop_ld, addr(loss_ps), addr(glen), addr(baseline_in),
           addr(baseline_out), addr(out), out_ld, B, T, A.stream())


def time_moments_fwd(h, h_bs, h_rs, length, B, T, Cn, S1, Q):
    A.call("ag_time_moments_fwd", addr(h), _dtype_of(h), h_bs, h_rs, addr(length), B, T, Cn, addr(S1), addr(Q), A.stream())


def time_moments_bwd(h, h_bs, h_rs, length, B, T, Cn, S1, Q, gm, gs, gf, dh):
    A.call("ag_time_moments_bwd", addr(h), _dtype_of(h), h_bs, h_rs, addr(length), B, T, Cn, addr(S1), addr(Q), addr(gm), addr(gs),
           addr(gf), addr(dh), _dtype_of(dh), A.stream())


def wn_table(entries, device):
    """entries: list of dicts(v,g,w,norm,dw,dv,dg,rows,cols,kind) with tensors/None -> device table + row_start."""
    arr = (A.WnEntry * len(entries))()
    starts, total = [], 0
    for i, e in enumerate(entries):
        for f in ("v", "g", "w", "norm", "dw", "dv", "dg"):
            setattr(arr[i], f, addr(e.get(f)))
        arr[i].rows, arr[i].cols, arr[i].kind = e["rows"], e["cols"], e["kind"]
        starts.append(total)
        total += e["rows"]
    raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
    rs = torch.tensor(starts, dtype=torch.int32).to(device)
    return raw, rs, total


def wn_fwd(table, row_start, nt, total_rows):
    A.call("ag_wn_fwd_multi", addr(table), addr(row_start), nt, total_rows, A.stream())


def wn_bwd(table, row_start, nt, total_rows):
    A.call("ag_wn_bwd_multi", addr(table), addr(row_start), nt, total_rows, A.stream())
