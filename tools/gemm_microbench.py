"""Isolated timing of the view-GEMM kernels on a few shapes (CUDA events, 20 reps after 3 warm-ups)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as T
from audiogan_b200 import kernels as Kn

def bench(fn, reps=20):
    for _ in range(3): fn()
    T.cuda.synchronize()
    e0, e1 = T.cuda.Event(enable_timing=True), T.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); T.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

shapes = [(16000, 1024, 1024), (16000, 512, 1792), (16000, 4096, 512), (256064, 128, 128), (16000, 256, 1024)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in sys.argv[1].split(","))]
for (M, N, K) in shapes:
    A32 = T.randn(M, K, device="cuda"); A16 = A32.bfloat16()
    B16 = T.randn(N, K, device="cuda").bfloat16(); B32 = B16.float()
    C = T.empty(M, N, device="cuda"); C16 = T.empty(M, N, device="cuda", dtype=T.bfloat16)
    fl = 2.0 * M * N * K
    t = bench(lambda: Kn.gemm_nt(M, N, K, A32, (M, 0, K), B16, K, C, (M, 0, N), tc=True))
    print("NT tc  A fp32 C fp32  M%d N%d K%d: %.3f ms %.1f TF/s" % (M, N, K, t, fl / t / 1e9))
    t = bench(lambda: Kn.gemm_nt(M, N, K, A16, (M, 0, K), B16, K, C16, (M, 0, N), tc=True))
    print("NT tc  A bf16 C bf16  M%d N%d K%d: %.3f ms %.1f TF/s" % (M, N, K, t, fl / t / 1e9))
    t = bench(lambda: Kn.gemm_nt(M, N, K, A32, (M, 0, K), B32, K, C, (M, 0, N)))
    print("NT f32                M%d N%d K%d: %.3f ms %.1f TF/s" % (M, N, K, t, fl / t / 1e9))
    t = bench(lambda: T.matmul(A16, B16.t()))
    print("torch bf16 matmul     M%d N%d K%d: %.3f ms %.1f TF/s" % (M, N, K, t, fl / t / 1e9))
    dw = T.zeros(N, K, device="cuda")
    Y = T.randn(M, N, device="cuda")
    t = bench(lambda: Kn.gemm_tn(M, N, K, Y, (M, 0, N), A32, (M, 0, K), dw, K, tc=True))
    print("TN tc                 M%d N%d K%d: %.3f ms %.1f TF/s" % (M, N, K, t, fl / t / 1e9))
