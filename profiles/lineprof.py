#!/usr/bin/env python
"""Per-CUDA-source-line totals (warp instructions executed, stall samples) of one kernel in an ncu report (needs -lineinfo
and --import-source on).  usage: python profiles/lineprof.py report.ncu-rep kernel-regex [topN]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ci, cs, cl, csrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), 0, 1
tot = {}
cur = None
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    if r[0] != "":
        cur = (r[0], r[1].strip())
        tot.setdefault(cur, [0, 0])
    elif cur:
        tot[cur][0] += int(r[ci]) if r[ci].isdigit() else 0
        tot[cur][1] += int(r[cs]) if r[cs].isdigit() else 0
TI = sum(v[0] for v in tot.values()) or 1
TS = sum(v[1] for v in tot.values()) or 1
print("total warp instructions %d, samples %d" % (TI, TS))
for (ln, src), (ins, smp) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5s  inst %5.1f%%  samples %5.1f%%  %s" % (ln, 100.0 * ins / TI, 100.0 * smp / TS, src[:100]))
