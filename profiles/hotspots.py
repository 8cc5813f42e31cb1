#!/usr/bin/env python
"""Top stall hot spots of a kernel from an ncu report's source page (SASS view with -lineinfo).
usage: python profiles/hotspots.py report.ncu-rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name"')
for blk in blocks[1:]:
    lines = blk.splitlines()
    print("== kernel", lines[0][:120])
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[1:] if len(r) == len(hdr)]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:top]
    for i in sorted(order):
        r = data[i]; n = int(r[ix["# Samples"]] or 0)
        st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
        print("%5d %5.1f%%  %-70s %s" % (i, 100.0 * n / tot, r[ix["Source"]].strip()[:70], " ".join("%s:%d" % (c, v) for v, c in st if v)))
