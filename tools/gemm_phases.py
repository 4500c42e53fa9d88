"""Phase cycle counters of the NT tensor-core GEMM (ag_gemm_dbg_*): where a CTA's time goes, per k-block."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as T
from audiogan_b200 import kernels as Kn, _abi as A

L = A.lib()
L.ag_gemm_dbg_enable.argtypes = [C.c_int]
L.ag_gemm_dbg_read.argtypes = [C.POINTER(C.c_ulonglong)]
for (M, N, K, dt) in ((16000, 1024, 1024, "fp32"), (16000, 1024, 1024, "bf16"), (32000, 4096, 512, "bf16"), (256064, 128, 128, "fp32")):
    A_ = T.randn(M, K, device="cuda")
    Cc = T.empty(M, N, device="cuda")
    if dt == "bf16":
        A_, Cc = A_.bfloat16(), Cc.bfloat16()
    B16 = T.randn(N, K, device="cuda").bfloat16()
    for _ in range(3):
        Kn.gemm_nt(M, N, K, A_, (M, 0, K), B16, K, Cc, (M, 0, N), tc=True)
    T.cuda.synchronize()
    L.ag_gemm_dbg_enable(1)
    e0, e1 = T.cuda.Event(enable_timing=True), T.cuda.Event(enable_timing=True)
    e0.record()
    Kn.gemm_nt(M, N, K, A_, (M, 0, K), B16, K, Cc, (M, 0, N), tc=True)
    e1.record()
    out = (C.c_ulonglong * 16)()
    L.ag_gemm_dbg_read(out)
    L.ag_gemm_dbg_enable(0)
    n, nkb = out[0], out[6] / max(out[0], 1)
    ms = e0.elapsed_time(e1)
    print("NT %s M%d N%d K%d: %.3f ms (%.0f TF/s), %d CTAs, %d k-blocks; per CTA cycles: main loop %.0f, epilogue %.0f | per k-block: "
          "producer wait-for-loads+convert %.0f, wait-slot-free %.0f, store+fence+arrive %.0f ; MMA thread wait-full %.0f" % (
              dt, M, N, K, ms, 2.0 * M * N * K / ms / 1e9, n, nkb, out[4] / n, out[5] / n, out[1] / n / nkb, out[2] / n / nkb,
              out[3] / n / nkb, out[7] / n / nkb))

# ---- TN (weight gradient) kernel: D tail shape (bf16 operands) and a generator conv wgrad with its real im2col view (fp32)
def tn_case(name, M, N, K, Y, yv, A_, av, ones):
    dw = T.zeros(N, K + 1, device="cuda")
    for _ in range(3):
        Kn.gemm_tn(M, N, K, Y, yv, A_, av, dw, K + 1, ones_col=ones, tc=True)
    T.cuda.synchronize()
    L.ag_gemm_dbg_enable(1)
    e0, e1 = T.cuda.Event(enable_timing=True), T.cuda.Event(enable_timing=True)
    e0.record()
    Kn.gemm_tn(M, N, K, Y, yv, A_, av, dw, K + 1, ones_col=ones, tc=True)
    e1.record()
    out = (C.c_ulonglong * 16)()
    L.ag_gemm_dbg_read(out)
    L.ag_gemm_dbg_enable(0)
    n, nst = max(out[8], 1), out[14] / max(out[8], 1)
    ms = e0.elapsed_time(e1)
    print("TN %s M%d N%d K%d ones=%d: %.3f ms (%.0f TF/s), %d CTAs, %.0f stages each; per CTA cycles: main loop %.0f, epilogue %.0f | per stage: "
          "wait-slot-free %.0f, row offsets+bar %.0f, loads+convert+store+arrive %.0f" % (
              name, M, N, K, ones, ms, 2.0 * M * N * K / ms / 1e9, n, nst, out[12] / n, out[13] / n, out[9] / n / nst, out[10] / n / nst, out[11] / n / nst))

M, N, K = 32000, 1024, 1024
tn_case("bf16 tail", M, N, K, T.randn(M, N, device="cuda").bfloat16(), (M, 0, N), T.randn(M, K, device="cuda").bfloat16(), (M, 0, K), True)
tn_case("fp32 tail", M, N, K, T.randn(M, N, device="cuda"), (M, 0, N), T.randn(M, K, device="cuda"), (M, 0, K), True)
# generator block 1 conv wgrad: B=64, L=16000, k=9, s=4 -> Lh=4000; dense buffer [B, 8+L+8, 120], prefix cin=24..: use k*cin = 17*48 of block 0 style
B_, Lx, CT, k, s_, cin, hid = 64, 16000, 120, 17, 8, 8, 128
Lh, Lp = Lx // s_, Lx + 16
Xd = T.randn(B_, Lp, CT, device="cuda")
dH = T.randn(B_, Lh + 2, hid, device="cuda")
tn_case("G conv0 wgrad (k17 s8 cin8 -> 128)", B_ * Lh, hid, k * cin, (dH, hid), (Lh, (Lh + 2) * hid, hid), (Xd, 0), (Lh, Lp * CT, s_ * CT, cin, CT), True)
k, s_, cin, hid = 9, 4, 88, 32
Lh = Lx // s_
dH = T.randn(B_, Lh + 2, hid, device="cuda")
tn_case("G conv3 wgrad (k9 s4 cin88 -> 32)", B_ * Lh, hid, k * cin, (dH, hid), (Lh, (Lh + 2) * hid, hid), (Xd, 0), (Lh, Lp * CT, s_ * CT, cin, CT), True)

# L2-resident variant of the tail shape: if the per-stage cost drops, the large case is bound by L2 misses / HBM re-reads
for M2 in (4000, 8000, 16000):
    tn_case("bf16 tail, small M", M2, N, K, T.randn(M2, N, device="cuda").bfloat16(), (M2, 0, N), T.randn(M2, K, device="cuda").bfloat16(), (M2, 0, K), True)
