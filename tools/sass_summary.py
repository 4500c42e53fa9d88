"""cuobjdump -sass of the built library -> per-kernel counts of the Blackwell-specific instructions
(tcgen05 MMA = UTC*MMA, tensor-memory load/store = LDTM/STTM, TMA = UTMALDG/UTMASTG/UBLKCP, cluster / DSMEM traffic,
legacy HMMA) -> profiles/r2_sass_summary.txt.   python tools/sass_summary.py [out]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audiogan_b200", "lib", "libaudiogan_b200.so")
PAT = collections.OrderedDict([
    ("tcgen05.mma (UTC*MMA)", re.compile(r"\bUTC[A-Z0-9]*MMA")), ("tcgen05.ld (LDTM)", re.compile(r"\bLDTM")),
    ("tcgen05.st (STTM)", re.compile(r"\bSTTM")), ("tcgen05.cp (UTCCP)", re.compile(r"\bUTCCP")),
    ("TMA tensor load (UTMALDG)", re.compile(r"\bUTMALDG")), ("TMA tensor store (UTMASTG)", re.compile(r"\bUTMASTG")),
    ("TMA L2 prefetch (UTMAPF)", re.compile(r"\bUTMAPF")),
    ("bulk copy (UBLKCP)", re.compile(r"\bUBLKCP")), ("mbarrier (SYNCS)", re.compile(r"\bSYNCS")),
    ("cluster barrier (UCGABAR)", re.compile(r"\bUCGABAR")), ("legacy mma.sync (HMMA)", re.compile(r"\bHMMA")),
    ("FFMA", re.compile(r"\bFFMA")), ("MUFU", re.compile(r"\bMUFU"))])


def main(out):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern, counts, size = None, collections.OrderedDict(), {}
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            counts[kern] = collections.Counter()
            size[kern] = 0
            continue
        if kern is None or "/*" not in line:
            continue
        ins = line.split("/*")[1] if line.strip().startswith("/*") else line
        if re.search(r"/\*[0-9a-f]{4,5}\*/", line):
            size[kern] += 1
            for name, pat in PAT.items():
                if pat.search(line):
                    counts[kern][name] += 1
    rows = ["# cuobjdump -sass %s : instruction counts per kernel (static SASS, sm_100a)" % os.path.relpath(LIB, ROOT),
            "# columns: " + " | ".join(PAT)]
    tot = collections.Counter()
    for k, c in sorted(counts.items(), key=lambda kv: -sum(kv[1][n] for n in list(PAT)[:7])):
        tot.update(c)
        rows.append("%-90s insts %6d | %s" % (k[:90], size[k], " ".join("%5d" % c[n] for n in PAT)))
    rows.insert(2, "%-90s              | %s" % ("TOTAL", " ".join("%5d" % tot[n] for n in PAT)))
    with open(out, "w") as f:
        f.write("\n".join(rows) + "\n")
    print("\n".join(rows[:14]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_summary.txt"))
