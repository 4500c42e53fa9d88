"""Per-phase cycle breakdown of the persistent LSTM kernels (desc.dbg counters), bf16 mode, bench shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as T
from audiogan_b200 import kernels as Kn
dev = "cuda"
def run(H, B, Tn, ndir, Fr, prec=1):
    sc = 1.0 / H ** 0.5
    FP = (Fr + 1 + 7) // 8 * 8 if Fr else 0
    pre = T.randn(B, Tn, ndir * 4 * H, device=dev)
    w1 = (T.randn(ndir, 4 * H, H + Fr, device=dev) * sc).contiguous()
    w2 = (T.randn(Fr + 1, H, device=dev) * sc) if Fr else None
    b2 = (T.randn(Fr + 1, device=dev) * sc) if Fr else None
    if Fr:
        w1t = T.cat([w1[0, :, :H].t(), w2[:Fr].t(), w2[Fr:].t(), T.zeros(H, FP - Fr - 1, device=dev)], 1).contiguous()
        wxt = w1[0, :, H:].t().contiguous()
    else:
        w1t, wxt = w1.permute(0, 2, 1).contiguous(), None
    dh_ext = None if Fr else T.randn(B, Tn, ndir * H, device=dev)
    dx_ext = T.randn(B, Tn, Fr, device=dev) if Fr else None
    hbuf, gates, cbuf = T.zeros(B, Tn + 2, ndir * H, device=dev), T.empty(B, Tn, ndir * 4 * H, device=dev), T.empty(B, Tn, ndir * H, device=dev)
    xbuf = T.zeros(B, Tn + 1, Fr, device=dev) if Fr else None
    sbuf = T.zeros(B, Tn, device=dev) if Fr else None
    misc = T.zeros(16, dtype=T.int32, device=dev)
    dbg = T.zeros(148, 8, dtype=T.int64, device=dev)
    kw = dict(prec=prec, hbuf16=T.zeros(B, Tn + 2, ndir * H, device=dev, dtype=T.bfloat16),
              xbuf16=T.zeros(B, Tn + 1, Fr, device=dev, dtype=T.bfloat16) if Fr else None)
    for name in ("fwd", "bwd"):
        for rep in range(2):
            dbg.zero_()
            e0, e1 = T.cuda.Event(enable_timing=True), T.cuda.Event(enable_timing=True)
            e0.record()
            if name == "fwd":
                Kn.lstm_fwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=ndir, F=Fr, pre=pre, w1=w1, w2=w2, b2=b2, hbuf=hbuf, gates=gates, cbuf=cbuf,
                            xbuf=xbuf, sbuf=sbuf, t_end=(misc, 8) if Fr else None, barrier=misc, dbg=dbg, **kw)
            else:
                dgates = T.empty(B, Tn, ndir * 4 * H, device=dev)
                dpx = T.empty(B, Tn, FP, device=dev) if Fr else None
                kb = dict(prec=prec, dgates16=T.empty(B, Tn, ndir * 4 * H, device=dev, dtype=T.bfloat16),
                          dpx16=T.empty(B, Tn, FP, device=dev, dtype=T.bfloat16) if Fr else None)
                Kn.lstm_bwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=ndir, F=Fr, gates=gates, cbuf=cbuf, xbuf=xbuf, dh_ext=dh_ext,
                            dx_ext=dx_ext, dgates=dgates, dpx=dpx, w1t=w1t, wxt=wxt, barrier=misc, dbg=dbg, **kb)
            e1.record(); T.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        dd = dbg.cpu().float()
        used = dd[:, 4] > 0
        m = dd[used].mean(0) / Tn
        mx = dd[used].max(0)[0] / Tn
        print("%s H%d ndir%d F%d T%d B%d prec%d: %.3f ms (%.1f us/step), %d CTAs; cycles/step mean gemm %.0f cell %.0f barrier %.0f phase2/A %.0f total %.0f | gemm split: wait+sync %.0f stage-issue %.0f mma %.0f" % (
            name, H, ndir, Fr, Tn, B, prec, ms, ms * 1e3 / Tn, int(used.sum()), m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7]))
for prec in (1,):
    run(512, 64, 250, 2, 0, prec)
    run(1024, 64, 80, 1, 200, prec)
