"""ctypes binding of libaudiogan_b200.so (the C ABI declared in include/audiogan_b200.h).

The library is the product path: if it is missing or a call fails this module raises --
there is no CPU / eager fallback.  Every wrapper launches on torch's current CUDA stream.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libaudiogan_b200.so")

AG_OK = 0
_ERRNAMES = {-1: "AG_EINVAL", -2: "AG_ECUDA", -3: "AG_ENOTSUP"}

i32, i64, f32, f64, vp = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_void_p


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", i64), ("N", i64), ("K", i64),
        ("A", vp), ("a_rpb", i64), ("a_bs", i64), ("a_rs", i64), ("a_kin", i64), ("a_k1s", i64),
        ("B", vp), ("ldb", i64),
        ("C", vp), ("c_rpb", i64), ("c_bs", i64), ("c_rs", i64), ("c_nin", i64), ("c_n1s", i64),
        ("alpha", f32),
        ("bias", vp), ("bias_mod", i64),
        ("rowbias", vp), ("rowbias_ld", i64),
        ("skip", vp),
        ("act", i32),
        ("dact", vp),
        ("slope", f32),
        ("mask_len", vp),
        ("mask_tmul", i64), ("mask_n1mul", i64), ("mask_toff", i64),
        ("a_dtype", i32), ("b_dtype", i32), ("c_dtype", i32), ("aux_dtype", i32),
        ("a_layout", i32),
    ]


class PadEntry(C.Structure):
    _fields_ = [("buf", C.c_void_p), ("B", C.c_int64), ("rows", C.c_int64), ("row_bytes", C.c_int64), ("head", C.c_int64),
                ("tail0", C.c_int64)]


class LstmDesc(C.Structure):
    _fields_ = [
        ("B", i32), ("T", i32), ("Tcap", i32), ("H", i32), ("ndir", i32), ("F", i32),
        ("pre", vp), ("w1", vp), ("w2", vp), ("b2", vp),
        ("hbuf", vp), ("gates", vp), ("cbuf", vp),
        ("len", vp),
        ("xbuf", vp), ("sbuf", vp),
        ("u", vp), ("stop", vp), ("glen", vp), ("t_end", vp),
        ("dh_ext", vp), ("dx_ext", vp), ("ds_ext", vp),
        ("dgates", vp), ("dpx", vp),
        ("w1t", vp), ("wxt", vp),
        ("barrier", vp),
        ("dh_ext_bs", i64),
        ("prec", i32), ("flags", i32),
        ("hbuf16", vp), ("xbuf16", vp), ("dgates16", vp), ("dpx16", vp),
        ("dbg", vp),
        ("ll_ws", vp), ("ll_ws_bytes", i64),
    ]


class WnEntry(C.Structure):
    _fields_ = [("v", vp), ("g", vp), ("w", vp), ("norm", vp), ("dw", vp), ("dv", vp), ("dg", vp),
                ("rows", i32), ("cols", i32), ("kind", i32), ("reserved", i32)]


class MtEntry(C.Structure):
    _fields_ = [("p", vp), ("g", vp), ("s1", vp), ("s2", vp), ("n", i64)]


class EwDesc(C.Structure):
    _fields_ = [
        ("B", i64), ("T", i64), ("C", i64),
        ("g1", vp), ("g1_bs", i64), ("g1_rs", i64), ("g1_cs", i64),
        ("g2", vp), ("g2_bs", i64), ("g2_rs", i64), ("g2_cs", i64),
        ("act", vp), ("a_bs", i64), ("a_rs", i64),
        ("slope", f32), ("reserved", i32),
        ("len", vp),
        ("out", vp), ("pad_l", i64), ("pad_r", i64),
        ("acc", vp), ("acc_bs", i64), ("acc_rs", i64),
        ("g1_dtype", i32), ("g2_dtype", i32), ("act_dtype", i32), ("acc_dtype", i32), ("out_dtype", i32), ("reserved2", i32),
        ("colsum", vp),
    ]


# name -> argtypes (restype is int for every entry but the error string)
_PROTOS = {
    "ag_version": [],
    "ag_sync_check": [vp],
    "ag_device_info": [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "ag_gemm_nt_f32": [C.POINTER(GemmDesc), vp],
    "ag_gemm_tn_f32": [C.POINTER(GemmDesc), vp, i64, i32, vp],
    "ag_lstm_fwd": [C.POINTER(LstmDesc), vp],
    "ag_lstm_bwd": [C.POINTER(LstmDesc), vp],
    "ag_lstm_cluster_max_active": [i32, i32],
    "ag_lstm_batch_cap": [C.POINTER(LstmDesc), i32],
    "ag_wn_fwd_multi": [vp, vp, i32, i32, vp],
    "ag_wn_bwd_multi": [vp, vp, i32, i32, vp],
    "ag_gather": [vp, vp, vp, i64, i32, vp],
    "ag_frame_noise": [vp, i64, i64, vp, i64, vp, f32, i64, i64, i32, vp],
    "ag_bce_fwd": [vp, vp, vp, vp, i64, i64, vp],
    "ag_bce_bwd": [vp, vp, vp, vp, vp, i64, i64, vp],
    "ag_bce_const_fused": [vp, i64, vp, f32, f32, vp, vp, vp, vp, i64, i64, vp],
    "ag_reinforce_dlogit": [vp, i64, vp, i64, vp, vp, vp, vp, vp, i64, i64, i64, vp],
    "ag_time_moments_fwd": [vp, i32, i64, i64, vp, i64, i64, i64, vp, vp, vp],
    "ag_time_moments_bwd": [vp, i32, i64, i64, vp, i64, i64, i64, vp, vp, vp, vp, vp, vp, i32, vp],
    "ag_ew_grad": [C.POINTER(EwDesc), vp],
    "ag_colsum": [vp, i32, i64, i64, i64, i64, i64, vp, vp],
    "ag_outer_dact": [vp, vp, vp, i32, vp, i32, i64, i64, f32, vp],
    "ag_conv1in_fwd": [vp, i64, vp, vp, vp, i32, i64, i32, i32, i64, i64, i64, vp, f32, vp],
    "ag_conv1in_wgrad": [vp, i32, i64, vp, i64, vp, i32, i32, i64, i64, i64, vp],
    "ag_conv1in_dgrad": [vp, i32, i64, vp, vp, i64, i32, i32, i32, i64, i64, i64, i64, vp],
    "ag_wcolsum": [vp, vp, i32, i64, i64, vp, vp],
    "ag_rowdot": [vp, i32, vp, vp, i64, i64, vp, vp],
    "ag_conv1out_fwd": [vp, i32, i64, i64, i32, vp, vp, vp, i64, i64, vp],
    "ag_conv1out_dgrad": [vp, vp, vp, i32, i64, i64, i32, i64, i64, vp],
    "ag_conv1out_wgrad": [vp, vp, i32, i64, i64, i32, vp, i64, i64, vp],
    "ag_copy3d": [vp, i64, i64, i64, vp, i64, i64, i64, i64, i64, i64, i32, i32, i32, vp],
    "ag_zero_pads": [vp, i64, i64, i64, i64, i64, vp],
    "ag_zero_pads_multi": [vp, i32, vp],
    "ag_frames_to_slot": [vp, i32, i64, i64, i32, vp, i64, i64, i64, vp],
    "ag_rowgroup_sum": [vp, i32, vp, i64, i64, i64, vp],
    "ag_transpose_bct": [vp, vp, i64, i64, i64, i64, i64, i32, vp],
    "ag_lstm_step_cell_fwd": [vp, vp, i64, vp, i64, vp, i64, vp, vp, i64, vp, vp, i64, i32, i32, vp],
    "ag_gen_step_proj_finish": [vp, i32, vp, vp, i64, vp, i64, i32, vp, i64, i32, i32, vp],
    "ag_gen_step_dpx": [vp, i64, vp, i64, vp, i64, vp, i64, vp, vp, i64, vp, i64, i32, i32, i32, i32, vp],
    "ag_lstm_step_cell_bwd": [vp, vp, i64, vp, vp, i64, vp, vp, vp, i64, vp, i64, i32, i32, vp],
    "ag_peer_barrier": [vp, i32, i32, vp],
    "ag_peer_allreduce": [vp, vp, i32, i32, i64, i32, vp],
    "ag_mt_sqnorm": [vp, vp, vp, i32, i32, vp, vp, f32, vp],
    "ag_mt_clip": [vp, vp, vp, i32, i32, vp, f32, vp],
    "ag_mt_rmsprop": [vp, vp, vp, i32, i32, vp, f32, f32, f64, f64, f64, vp],
    "ag_mt_adam": [vp, vp, vp, i32, i32, vp, f32, f32, f64, f64, f64, f64, i32, vp],
}
_PROTOS.update({
    "ag_gemm_nt_tc": [C.POINTER(GemmDesc), vp],
    "ag_gemm_tn_tc": [C.POINTER(GemmDesc), vp, i64, i32, vp],
    "ag_gemm_dbg_enable": [i32],
    "ag_gemm_dbg_read": [vp],
})
_OPTIONAL = {}

_lib = None


def lib():
    """Load the shared library once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "audiogan_b200: %s is missing -- run `python -m audiogan_b200._build` "
            "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, args in _PROTOS.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    for name, args in _OPTIONAL.items():
        if hasattr(L, name):
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
    L.ag_last_error_string.argtypes = []
    L.ag_last_error_string.restype = C.c_char_p
    L.ag_lstm_last_path.argtypes = []
    L.ag_lstm_last_path.restype = C.c_char_p
    L.ag_lstm_workspace_bytes.argtypes = [C.POINTER(LstmDesc), i32]
    L.ag_lstm_workspace_bytes.restype = C.c_int64
    _lib = L
    return L


def exported_symbols():
    return list(_PROTOS) + ["ag_last_error_string", "ag_lstm_last_path", "ag_lstm_workspace_bytes"]


class AudioganError(RuntimeError):
    pass


def check(rc, who):
    if rc != AG_OK:
        msg = lib().ag_last_error_string().decode("utf-8", "replace")
        raise AudioganError("%s failed: %s (%s)" % (who, _ERRNAMES.get(rc, rc), msg))


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_get_device = getattr(torch._C, "_cuda_getDevice", None)


def stream():
    """torch's current CUDA stream as a raw handle.  Called once per kernel launch (~170 per step): the private fast path
    costs ~0.3 us, torch.cuda.current_stream().cuda_stream ~15 us (it builds a Stream object); falls back if the private
    accessors are absent."""
    if _raw_stream is not None and _get_device is not None:
        return C.c_void_p(_raw_stream(_get_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Keeps no reference: the caller does."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


launches = 0      # calls into kernel-launching entry points (bench.py reports it as gpu_launches)
_hook = None      # optional profiler hook: hook(name, args) -> callable(done) or None


def set_hook(h):
    global _hook
    _hook = h


def call(name, *args):
    global launches
    launches += 1
    if _hook is not None:
        fin = _hook(name, args)
        check(getattr(lib(), name)(*args), name)
        if fin is not None:
            fin()
        return
    check(getattr(lib(), name)(*args), name)


def device_info():
    sm, smem, cc = C.c_int(), C.c_int(), C.c_int()
    call("ag_device_info", C.byref(sm), C.byref(smem), C.byref(cc))
    return sm.value, smem.value, cc.value
