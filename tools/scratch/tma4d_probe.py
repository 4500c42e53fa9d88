import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch as T
from audiogan_b200 import _abi as A
L = A.lib()
f = L.ag_dbg_tma4d
f.argtypes = [C.c_void_p] + [C.c_int64] * 7 + [C.c_int] * 4 + [C.c_void_p] + [C.c_int] * 4
f.restype = C.c_int
cin, taps, rows, nb, CT, s = [int(a) for a in sys.argv[1:7]]
c = [int(a) for a in sys.argv[7:11]]
bx = [int(a) for a in sys.argv[11:15]]
ts_ = int(sys.argv[15]) if len(sys.argv) > 15 else None
Lp = rows * s + 64
# element value encodes (batch, buffer row, channel): v = row * 128 + ch  (exact in bf16? no -> use int16 bit patterns)
buf = T.zeros(nb, Lp, CT, dtype=T.int16)
buf += (T.arange(Lp, dtype=T.int16)[None, :, None] * 128 + T.arange(CT, dtype=T.int16)[None, None, :])
buf += (T.arange(nb, dtype=T.int16)[:, None, None] * 16384)
g = buf.cuda()
out = T.full((8192,), -1, dtype=T.int16, device="cuda")
rc = f(g.data_ptr(), cin, taps, rows, nb, ts_ if ts_ else CT, s * CT, Lp * CT, c[0], c[1], c[2], c[3], out.data_ptr(), bx[0], bx[1], bx[2], bx[3])
print("args", sys.argv[1:], "rc", rc, L.ag_last_error_string(), flush=True)
T.cuda.synchronize()
o = out.cpu().view(128, 8, 8)       # assumed [row][16-byte chunk][8 elements]
for r in (0, 1, 2, 9, 127):
    print("smem row", r)
    for ch in range(8):
        v = o[r, ch]
        print("   chunk", ch, [(int(x) // 16384, (int(x) % 16384) // 128, int(x) % 128) if x >= 0 else None for x in v[:2]], "...", (int(v[7]) % 16384 // 128, int(v[7]) % 128))
