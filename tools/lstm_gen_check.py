"""TMEM-resident generator recurrence (csrc/lstm_gen.cu): parity against the fp32 and bf16 grid-barrier kernels on the
same inputs (with and without stop sampling / early exit), timing, per-phase cycle counters."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as T
from audiogan_b200 import kernels as Kn

dev = "cuda"


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run(H, B, Tn, Fr=200, stops=False, time_it=True, bwd=True):
    T.manual_seed(7)
    sc = 1.0 / H ** 0.5
    FP = (Fr + 1 + 7) // 8 * 8
    pre = T.randn(B, Tn, 4 * H, device=dev)
    w1 = (T.randn(1, 4 * H, H + Fr, device=dev) * sc).contiguous()
    w2 = (T.randn(Fr + 1, H, device=dev) * sc).contiguous()
    b2 = (T.randn(Fr + 1, device=dev) * sc).contiguous()
    if stops:
        b2[Fr] = -3.0                                    # stop probability ~5 % per frame
    u = T.rand(B, Tn, device=dev) if stops else None
    w1t = T.cat([w1[0, :, :H].t(), w2[:Fr].t(), w2[Fr:].t(), T.zeros(H, FP - Fr - 1, device=dev)], 1).contiguous()
    wxt = w1[0, :, H:].t().contiguous()
    dx_ext = T.randn(B, Tn, Fr, device=dev)
    ds_ext = T.randn(B, Tn, device=dev)
    out = {}
    for name, prec, flags in (("fp32", 0, 0), ("grid16", 1, 1), ("tmem", 1, 2)):
        tf = tb = 0.0
        for rep in range(3 if time_it else 1):
            hbuf, gates, cbuf = T.zeros(B, Tn + 2, H, device=dev), T.zeros(B, Tn, 4 * H, device=dev), T.zeros(B, Tn, H, device=dev)
            xbuf, sbuf = T.zeros(B, Tn + 1, Fr, device=dev), T.zeros(B, Tn, device=dev)
            stop, glen = T.zeros(B, Tn, dtype=T.int32, device=dev), T.zeros(B, dtype=T.int32, device=dev)
            misc = T.zeros(1024, dtype=T.int32, device=dev)
            dbg = T.zeros(148, 8, dtype=T.int64, device=dev)
            ll_ws = Kn.lstm_workspace(B, H, Fr, False, dev)            # sized by the library (ag_lstm_workspace_bytes)
            hbuf16 = T.zeros(B, Tn + 2, H, device=dev, dtype=T.bfloat16) if prec else None
            xbuf16 = T.zeros(B, Tn + 1, Fr, device=dev, dtype=T.bfloat16) if prec else None
            e = [T.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            Kn.lstm_fwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=1, F=Fr, pre=pre, w1=w1, w2=w2, b2=b2, hbuf=hbuf, gates=gates, cbuf=cbuf,
                        xbuf=xbuf, sbuf=sbuf, u=u, stop=stop, glen=glen, t_end=(misc, 8), barrier=misc, prec=prec,
                        flags=flags, hbuf16=hbuf16, xbuf16=xbuf16, dbg=dbg if name == "tmem" else None,
                        ll_ws=ll_ws, ll_ws_bytes=ll_ws.numel())
            e[1].record()
            T.cuda.synchronize()
            t_end = int(misc[8].item())
            dfw = dbg.cpu().float().clone()
            dgates = T.zeros(B, Tn, 4 * H, device=dev)
            dpx = T.zeros(B, Tn, FP, device=dev)
            dgates16 = T.zeros(B, Tn, 4 * H, device=dev, dtype=T.bfloat16) if prec else None
            dpx16 = T.zeros(B, Tn, FP, device=dev, dtype=T.bfloat16) if prec else None
            dbg.zero_()
            misc2 = T.zeros(1024, dtype=T.int32, device=dev)
            ngr = (B + 15) // 16
            ll_wb = Kn.lstm_workspace(B, H, Fr, True, dev)
            e[1].record()
            if bwd:
                Kn.lstm_bwd(B=B, T=t_end, Tcap=Tn, H=H, ndir=1, F=Fr, gates=gates, cbuf=cbuf, xbuf=xbuf, dx_ext=dx_ext, ds_ext=ds_ext,
                            dgates=dgates, dpx=dpx, w1t=w1t, wxt=wxt, barrier=misc2, prec=prec, flags=flags, dgates16=dgates16,
                            dpx16=dpx16, dbg=dbg if name == "tmem" else None, ll_ws=ll_wb, ll_ws_bytes=ll_wb.numel())
            e[2].record()
            T.cuda.synchronize()
            dbw = dbg.cpu().float().clone()
            tf, tb = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        out[name] = dict(h=hbuf, x=xbuf, s=sbuf, gates=gates, c=cbuf, dgates=dgates, dpx=dpx, stop=stop, glen=glen, t_end=t_end,
                         x16=xbuf16, h16=hbuf16)
        print("  %-7s H%d B%d T%d stops=%d: fwd %.3f ms (%.2f us/step) bwd %.3f ms (%.2f us/step) t_end %d" % (
            name, H, B, Tn, stops, tf, tf * 1e3 / max(t_end, 1), tb, tb * 1e3 / max(t_end, 1), t_end), flush=True)
        if name == "tmem":
            for nm, dd, labels in (("fwd", dfw, "x-exchange, mma-x+wait, tmem-ld+cell+LL-store, stores+prefetch, h-exchange, phase2+LL-store, -"),
                                   ("bwd", dbw, "mma-x+emit x tiles, x-rows reduce+dpx, wait mma+emit unit tiles, unit reduce, dpx read+wp^T dpx, cell, -")):
                used = dd[:, 7] > 0
                if used.sum() == 0:
                    print("    %s: TMEM-resident kernel did not run (fallback)" % nm)
                    continue
                m = dd[used].mean(0)
                print("    %s cycles/CTA total %.0f; per step [%s]: %s" % (nm, m[7], labels, " ".join("%.0f" % x for x in (m[:7] / max(t_end, 1)))))
    f, g, c = out["fp32"], out["grid16"], out["tmem"]
    ok = True
    te = f["t_end"]
    for nm in ("h", "x", "s", "gates", "c", "dgates", "dpx"):
        if not bwd and nm in ("dgates", "dpx"):
            continue
        rc, rg = rel(c[nm][:, :te + 1] if nm in ("h", "x") else c[nm][:, :te], f[nm][:, :te + 1] if nm in ("h", "x") else f[nm][:, :te]), rel(g[nm], f[nm])
        print("    %-6s rel err vs fp32: tmem %.2e   grid-bf16 %.2e" % (nm, rc, rg))
        ok &= rc < 3e-2
    if stops:
        same = bool((c["stop"][:, :te] == f["stop"][:, :te]).all()) and bool((c["glen"] == f["glen"]).all()) and c["t_end"] == f["t_end"]
        print("    stop flags / lengths / t_end identical to the fp32 kernel: %s (t_end %d vs %d)" % (same, c["t_end"], f["t_end"]))
    return ok


if __name__ == "__main__":
    ok = True
    for (H, B, Tn, st) in ((1024, 16, 3, False), (1024, 64, 6, False), (1024, 37, 12, True), (1024, 64, 80, False), (1024, 64, 80, True), (1024, 128, 80, False)):
        print("case H=%d B=%d T=%d stops=%s" % (H, B, Tn, st), flush=True)
        ok &= run(H, B, Tn, stops=st, time_it=Tn > 50)
    print("ALL OK" if ok else "MISMATCH")
