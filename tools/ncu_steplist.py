#!/usr/bin/env python
"""One training step of an ncu launch-list CSV in launch order: kernel, grid, stream, duration, DRAM bytes.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --launch-skip 1000 --launch-count 500 --csv --log-file launches.csv python bench.py --quick --steps 2 --warmup 3 --no-graph
    python tools/ncu_steplist.py launches.csv [--torch]

The step is delimited by two consecutive launches of the generator's forward recurrence (the first kernel of a core step);
torch's own fill / copy kernels are hidden unless --torch is given.  Times are cold-cache and serialised (every kernel alone
on the GPU): they locate the work of a step phase by phase, they are not the step time."""
import csv
import re
import sys


def main(path, show_torch):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    by = {}
    for x in csv.DictReader(lines):
        i = int(x["ID"])
        e = by.setdefault(i, {"name": x["Kernel Name"], "grid": x["Grid Size"], "stream": x["Stream"]})
        e[x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
    ids = sorted(by)
    gen = [i for i in ids if "lstm_gen_fwd" in by[i]["name"] or "lstm_step_cell_fwd" in by[i]["name"]]
    if len(gen) < 2:
        raise SystemExit("need two generator forward launches in the capture window (found %d)" % len(gen))
    s, e = gen[-2], gen[-1]
    tot = lib = 0.0
    for i in range(s, e):
        k = by[i]
        us = k.get("gpu__time_duration.sum", 0.0) / 1000
        tot += us
        own = "at::" not in k["name"]
        lib += us if own else 0.0
        if own or show_torch:
            print("%4d %-60s %-14s s%-4s %8.1f us  R %7.1f W %7.1f MB" % (
                i - s, re.sub(r"\(.*", "", k["name"])[:60], k["grid"], k["stream"], us,
                k.get("dram__bytes_read.sum", 0) / 1e6, k.get("dram__bytes_write.sum", 0) / 1e6))
    print("# step of %d launches: %.1f us serialised, %.1f us in the library's own kernels" % (e - s, tot, lib))


if __name__ == "__main__":
    main(sys.argv[1], "--torch" in sys.argv[2:])
