import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch as T, torch.nn.functional as F
from audiogan_b200 import kernels as Kn
Bn, cin, CT, hid, L, k, s, PAD = [int(a) for a in sys.argv[1:9]]
T.manual_seed(22)
p, Lh = (k - 1) // 2, L // s
Lp = L + 2 * PAD
Xd = T.zeros(Bn, Lp, CT)
Xd[:, PAD:PAD + L] = T.randn(Bn, L, CT)
Xd = Xd.bfloat16()
w = (T.randn(hid, cin, k) / (cin * k) ** 0.5).bfloat16()
b = T.randn(hid)
G_, KT = cin // 8, (k + 7) // 8
Kq = G_ * KT * 64
wq = T.zeros(hid, G_, KT * 8, 8)
wq[:, :, :k] = w.float().view(hid, G_, 8, k).permute(0, 1, 3, 2)
wq = wq.reshape(hid, Kq).bfloat16().cuda()
x = Xd[:, PAD:PAD + L, :cin].float().permute(0, 2, 1)
ref = F.leaky_relu(F.conv1d(x, w.float(), b, stride=s, padding=p)[:, :, :Lh], 0.01)
Hh = T.zeros(Bn, Lh + 2, hid, device="cuda", dtype=T.bfloat16)
Xg = Xd.cuda()
print("A base mod 128:", (Xg.data_ptr() + (PAD - p) * CT * 2) % 128, flush=True)
Kn.gemm_nt(Bn * Lh, hid, k * cin, (Xg, (PAD - p) * CT), (Lh, Lp * CT, s * CT, cin, CT), wq, Kq, (Hh, hid),
           (Lh, (Lh + 2) * hid, hid), bias=b.cuda(), act=1, tc=True, a_layout=1)
T.cuda.synchronize()
d = (Hh[:, 1:Lh + 1].float().cpu() - ref.permute(0, 2, 1)).abs().max() / ref.abs().max()
print("args", sys.argv[1:], "rel err", float(d), flush=True)
