"""Phase cycle counters of the persistent TMA-fed NT GEMM (ag_gemm_dbg_*): per tile, how long the epilogue warps wait for an
accumulator vs work on it, and how long the MMA thread waits for operands (TMA-bound) vs for a drained accumulator
(epilogue-bound).  Shapes: the discriminator tail (plain and with the skip / LeakyReLU' operands), a thin transposed-conv
GEMM writing a channel slot of the generator's dense buffer."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as T
from audiogan_b200 import kernels as Kn, _abi as A

L = A.lib()
L.ag_gemm_dbg_enable.argtypes = [C.c_int]
L.ag_gemm_dbg_read.argtypes = [C.POINTER(C.c_ulonglong)]


def run(name, fn, flops):
    for _ in range(3):
        fn()
    T.cuda.synchronize()
    e0, e1 = T.cuda.Event(enable_timing=True), T.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    T.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    L.ag_gemm_dbg_enable(MODE)
    fn()
    out = (C.c_ulonglong * 16)()
    L.ag_gemm_dbg_read(out)
    L.ag_gemm_dbg_enable(0)
    ctas, tiles, kbs = max(out[7], 1), max(out[0], 1), max(out[6], 1)
    print("%-46s %.3f ms %6.0f TF/s | %d CTAs %d tiles | per tile cycles: epilogue wait-acc %.0f work %.0f ; MMA thread total %.0f "
          "wait-operands %.0f wait-drained-acc %.0f | per k-block wait-operands %.0f" % (
              name, ms, flops / ms / 1e9, ctas, tiles, out[1] / tiles, out[2] / tiles, out[5] / tiles, out[3] / tiles,
              out[4] / tiles, out[3] / kbs), flush=True)


MODE = int(os.environ.get("DBG_MODE", "1"))
bf = T.bfloat16
Bn, Tm, S = 128, 250, 1024
M = Bn * Tm
geo = (Tm, (Tm + 2) * S, S)
h = T.randn(Bn, Tm + 2, S, device="cuda").to(bf)
r1 = T.empty_like(h)
W = (T.randn(S, S, device="cuda") / 32).to(bf)
bias = T.randn(S, device="cuda")
run("tail fwd: bias+skip+lrelu, bf16 out", lambda: Kn.gemm_nt(M, S, S, (h, S), geo, W, S, (r1, S), geo, bias=bias, skip=(h, S), act=1, tc=True), 2.0 * M * S * S)
run("tail dgrad: skip+dact, bf16 out", lambda: Kn.gemm_nt(M, S, S, (h, S), geo, W, S, (r1, S), geo, skip=(h, S), dact=(h, S), tc=True), 2.0 * M * S * S)
run("tail plain, bf16 out", lambda: Kn.gemm_nt(M, S, S, (h, S), geo, W, S, (r1, S), geo, tc=True), 2.0 * M * S * S)
o32 = T.empty(Bn, Tm + 2, S, device="cuda")
run("tail plain, fp32 out", lambda: Kn.gemm_nt(M, S, S, (h, S), geo, W, S, (o32, S), geo, tc=True), 2.0 * M * S * S)
W4 = (T.randn(4096, 512, device="cuda") / 32).to(bf)
f5 = T.randn(Bn, Tm, 512, device="cuda").to(bf)
pre = T.empty(Bn, Tm, 4096, device="cuda")
run("lstm input proj N4096 K512, fp32 out", lambda: Kn.gemm_nt(M, 4096, 512, f5, (Tm, Tm * 512, 512), W4, 512, pre, (Tm, Tm * 4096, 4096), tc=True), 2.0 * M * 4096 * 512)
Wk = (T.randn(512, 4096, device="cuda") / 64).to(bf)
dg = T.randn(M, 4096, device="cuda").to(bf)
df = T.empty(Bn, Tm, 512, device="cuda", dtype=bf)
run("lstm dfeat N512 K4096, bf16 out", lambda: Kn.gemm_nt(M, 512, 4096, dg, (M, 0, 4096), Wk, 4096, df, (Tm, Tm * 512, 512), tc=True), 2.0 * M * 512 * 4096)
# generator block 1 transposed conv: Hh [B, Lh+2, 64] -> slot of 32 channels, 4 phases (N = 128), K = 2*64
B_, Lx, CT, s_, hid, out, cin = 64, 16000, 120, 4, 64, 32, 24
Lh, Lp = Lx // s_, Lx + 16
Xd = T.zeros(B_, Lp, CT, device="cuda", dtype=bf)
Hh = T.randn(B_, Lh + 2, hid, device="cuda").to(bf)
Wd = (T.randn(s_ * out, 2 * hid, device="cuda") / 11).to(bf)
bd = T.randn(out, device="cuda")
lenL = T.full((B_,), Lx, device="cuda", dtype=T.int32)
Md = B_ * (Lh + 1)
run("G deconv1 into dense slot (N128 K128)", lambda: Kn.gemm_nt(Md, s_ * out, 2 * hid, Hh, (Lh + 1, (Lh + 2) * hid, hid), Wd, 2 * hid,
    (Xd, 6 * CT + cin), (Lh + 1, Lp * CT, s_ * CT, out, CT), bias=bd, bias_mod=out, skip=(Xd, 6 * CT + 8), act=1, mask_len=lenL, mask=(s_, 1, -2), tc=True),
    2.0 * Md * s_ * out * 2 * hid)
pk = T.empty(B_, Lh + 1, s_ * out, device="cuda", dtype=bf)
run("same GEMM, packed bf16 output, no skip", lambda: Kn.gemm_nt(Md, s_ * out, 2 * hid, Hh, (Lh + 1, (Lh + 2) * hid, hid), Wd, 2 * hid,
    pk, (Lh + 1, (Lh + 1) * s_ * out, s_ * out), bias=bd, bias_mod=out, act=1, tc=True), 2.0 * Md * s_ * out * 2 * hid)
