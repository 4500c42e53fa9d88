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
