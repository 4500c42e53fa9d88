#!/usr/bin/env python
"""Summarise ncu outputs into profiles/: (1) a launch list CSV (gpu__time_duration.sum) -> per-kernel share table,
(2) a .ncu-rep (--set full) -> the handful of metrics the roofline argument uses."""
import collections
import csv
import re
import subprocess
import sys

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "lts__t_bytes.sum", "sm__inst_executed_pipe_lsu", "smsp__inst_executed.sum")


def launches(path, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# source: %s ; total %.3f ms over %d launches\n" % (path, tot / 1e6, sum(a[0] for a in agg.values())))
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-72s n=%5d %11.3f ms %6.2f%%\n" % (k[:72], n, v / 1e6, 100 * v / tot))
    print(open(out).read())


# C-ABI entry point (bench.py's kernel families) <- kernel names
FAMILY = (("gemm_nt_tma_kernel", "ag_gemm_nt_tc"), ("gemm_nt_tc_kernel", "ag_gemm_nt_tc"), ("gemm_tn_tma_kernel", "ag_gemm_tn_tc"),
          ("gemm_tn_tc_kernel", "ag_gemm_tn_tc"), ("tn_bias_kernel", "ag_gemm_tn_tc"), ("lstm_cl_fwd", "ag_lstm_fwd"), ("lstm_gen_fwd", "ag_lstm_fwd"),
          ("lstm_fwd_kernel", "ag_lstm_fwd"), ("lstm_cl_bwd", "ag_lstm_bwd"), ("lstm_gen_bwd", "ag_lstm_bwd"), ("lstm_bwd_kernel", "ag_lstm_bwd"))


def traffic(path, out):
    """launch list with dram__bytes_read.sum / dram__bytes_write.sum beside gpu__time_duration.sum -> per kernel and per C-ABI
    family: launches, time share, DRAM bytes per launch (what bench.py reports as roofline.traffic)."""
    import json
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()          # launch id -> [name, ns, bytes]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", "")) * scale.get(row["Metric Unit"], 1.0)
        e = per.setdefault(row["ID"], [re.sub(r"\(.*", "", row["Kernel Name"]), 0.0, 0.0])
        if row["Metric Name"] == "gpu__time_duration.sum":
            e[1] += v
        elif row["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            e[2] += v
    kern, fam = collections.OrderedDict(), collections.OrderedDict()
    for name, ns, by in per.values():
        k = kern.setdefault(name, [0, 0.0, 0.0])
        k[0] += 1; k[1] += ns; k[2] += by
        fn = next((f_ for pat, f_ in FAMILY if pat in name), None)
        if fn:
            q = fam.setdefault(fn, [0, 0.0, 0.0])
            q[0] += 1; q[1] += ns; q[2] += by
    tot = sum(k[1] for k in kern.values())
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none\n")
        f.write("# (cold-cache, serialised: compare SHARES) source: %s ; total %.3f ms over %d launches\n" % (path, tot / 1e6, len(per)))
        for k, (n, ns, by) in sorted(kern.items(), key=lambda kv: -kv[1][1]):
            f.write("%-64s n=%5d %10.3f ms %6.2f%%  DRAM %9.2f MB/launch  %7.0f GB/s\n" % (k[:64], n, ns / 1e6, 100 * ns / tot, by / n / 1e6, by / ns))
        f.write("# per C-ABI family\n")
        for k, (n, ns, by) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
            f.write("%-64s n=%5d %10.3f ms %6.2f%%  DRAM %9.2f MB/launch\n" % (k, n, ns / 1e6, 100 * ns / tot, by / n / 1e6))
    with open(out.rsplit(".", 1)[0] + ".json", "w") as f:
        json.dump({k: {"launches": n, "dram_bytes_per_launch": by / n, "ms": ns / 1e6} for k, (n, ns, by) in fam.items()}, f, indent=1)
    print(open(out).read())


def report(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none ; source %s\n" % path)
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("\n== %s  grid %s block %s\n" % (d.get("Kernel Name", "?")[:100], d.get("Grid Size"), d.get("Block Size")))
            for h, u, v in zip(hdr, units, r):
                if any(h.startswith(k) for k in KEYS):
                    f.write("  %-75s %16s %s\n" % (h, v, u))
    print(open(out).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "report": report, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
