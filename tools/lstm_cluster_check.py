"""Cluster-resident BiLSTM kernels (csrc/lstm_cluster.cu): parity against the fp32 grid-barrier kernels on the same
inputs, timing against the bf16 grid-barrier kernels, and the per-phase cycle counters (desc.dbg)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as T
from audiogan_b200 import kernels as Kn, _abi as A

dev = "cuda"


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run(H, B, Tn, ndir=2, mixed=True, time_it=True):
    T.manual_seed(5)
    sc = 1.0 / H ** 0.5
    pre = T.randn(B, Tn, ndir * 4 * H, device=dev)
    w1 = (T.randn(ndir, 4 * H, H, device=dev) * sc).contiguous()
    w1t = w1.permute(0, 2, 1).contiguous()
    lens = T.randint(1, Tn + 1, (B,), device=dev, dtype=T.int32) if mixed else None
    if lens is not None:
        lens[0] = Tn
    dh_ext = T.randn(B, Tn, ndir * H, device=dev)
    out = {}
    for name, prec, flags in (("fp32", 0, 0), ("grid16", 1, 1), ("cluster", 1, 0)):
        hbuf, gates, cbuf = T.zeros(B, Tn + 2, ndir * H, device=dev), T.zeros(B, Tn, ndir * 4 * H, device=dev), T.zeros(B, Tn, ndir * H, device=dev)
        misc = T.zeros(16, dtype=T.int32, device=dev)
        dbg = T.zeros(148 * 4, 8, dtype=T.int64, device=dev)
        hbuf16 = T.zeros(B, Tn + 2, ndir * H, device=dev, dtype=T.bfloat16) if prec else None
        dgates = T.zeros(B, Tn, ndir * 4 * H, device=dev)
        dgates16 = T.zeros(B, Tn, ndir * 4 * H, device=dev, dtype=T.bfloat16) if prec else None
        tf = tb = 0.0
        for rep in range(3 if time_it else 1):
            dbg.zero_()
            e = [T.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            Kn.lstm_fwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=ndir, F=0, pre=pre, w1=w1, hbuf=hbuf, gates=gates, cbuf=cbuf, len=lens,
                        barrier=misc, prec=prec, flags=flags, hbuf16=hbuf16, dbg=dbg if name == "cluster" else None)
            e[1].record()
            T.cuda.synchronize()
            dfw = dbg.cpu().float().clone()
            dbg.zero_()
            e[1].record()
            Kn.lstm_bwd(B=B, T=Tn, Tcap=Tn, H=H, ndir=ndir, F=0, gates=gates, cbuf=cbuf, len=lens, dh_ext=dh_ext, dgates=dgates,
                        w1t=w1t, barrier=misc, prec=prec, flags=flags, dgates16=dgates16, dbg=dbg if name == "cluster" else None)
            e[2].record()
            T.cuda.synchronize()
            dbw = dbg.cpu().float().clone()
            tf, tb = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        out[name] = (hbuf.clone(), gates.clone(), cbuf.clone(), dgates.clone(), hbuf16, dgates16)
        print("  %-8s H%d B%d T%d: fwd %.3f ms (%.2f us/step)  bwd %.3f ms (%.2f us/step)" % (name, H, B, Tn, tf, tf * 1e3 / Tn, tb, tb * 1e3 / Tn), flush=True)
        if name == "cluster":
            for nm, dd, labels in (("fwd", dfw, "wait-h, mma, tmem-ld, act+sync, cell+fence+sync, push, prefetch+stores"), ("bwd", dbw, "mma, tmem-ld+stage+sync, push, wait-partials, reduce+cell+Bop+sync, stores+rearm, -")):
                used = dd[:, 7] > 0
                if used.sum() == 0:
                    print("    %s: cluster kernel did not run (fallback)" % nm)
                    continue
                m = dd[used].mean(0)
                print("    %s cycles/CTA total %.0f, in-loop %.0f; per phase [%s]: %s" % (
                    nm, m[7], m[:7].sum(), labels, " ".join("%.0f" % x for x in (m[:7] / Tn))))
    f, g, c = out["fp32"], out["grid16"], out["cluster"]
    names = ("h", "gates", "c", "dgates")
    ok = True
    for i, nm in enumerate(names):
        rc, rg = rel(c[i], f[i]), rel(g[i], f[i])
        print("    %-6s rel err vs fp32: cluster %.2e   grid-bf16 %.2e" % (nm, rc, rg))
        ok &= rc < 3e-2
    print("    h16 shadow %.2e  dgates16 shadow %.2e" % (rel(c[4].float(), c[0]), rel(c[5].float(), c[3])))
    return ok


if __name__ == "__main__":
    L = A.lib()
    print("max active clusters: H512 fwd %d bwd %d; H256 fwd %d; H128 fwd %d" % (
        L.ag_lstm_cluster_max_active(512, 0), L.ag_lstm_cluster_max_active(512, 1), L.ag_lstm_cluster_max_active(256, 0),
        L.ag_lstm_cluster_max_active(128, 0)), flush=True)
    ok = True
    for (H, B, Tn) in ((512, 16, 3), (512, 40, 9), (128, 5, 12), (256, 70, 17), (512, 64, 250), (512, 128, 250)):
        print("case H=%d B=%d T=%d" % (H, B, Tn), flush=True)
        ok &= run(H, B, Tn, time_it=Tn > 100)
    print("ALL OK" if ok else "MISMATCH")
