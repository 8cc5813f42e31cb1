#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.  usage: python profiles/launch_summary.py file.csv [skip_first_n]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[h]
kn, mv, mu, idc = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
agg, cnt = collections.OrderedDict(), collections.Counter()
for r in rows[h + 1:]:
    if len(r) <= mv or int(r[idc]) < skip:
        continue
    name = r[kn].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:60]
    v = float(r[mv].replace(",", ""))
    v = v / 1e3 if r[mu] == "ns" else (v * 1e3 if r[mu] == "ms" else v)
    agg[name] = agg.get(name, 0) + v
    cnt[name] += 1
tot = sum(agg.values())
for k, v in sorted(agg.items(), key=lambda x: -x[1]):
    print("%-50s n=%4d  total %10.1f us  avg %9.1f us  %5.1f%%" % (k, cnt[k], v, v / cnt[k], 100 * v / tot))
