#!/usr/bin/env python
"""Per-kernel totals of one steady-state step from an ncu launch list (gpu__time_duration.sum).
usage: python scripts/launch_summary.py launches.csv [step_index_from_end=2]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = next(k for k, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[hi + 1:]
names = [r[ix["Kernel Name"]] for r in data]
starts = [k for k, n in enumerate(names) if "k_bbox<" in n]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 2
s = starts[-which]; e = starts[-which + 1] if which > 1 else len(data)
tot = 0.0
agg = collections.OrderedDict()
for r in data[s:e]:
    v = float(r[ix["Metric Value"]].replace(",", ""))
    u = r[ix["Metric Unit"]]
    v = v*{"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
    nm = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    agg[nm] = agg.get(nm, 0.0) + v
    tot += v
for k, v in agg.items():
    print("%-60s %9.1f us %5.1f%%" % (k[:60], v, 100*v/tot))
print("%-60s %9.1f us (%d launches)" % ("TOTAL one step", tot, e - s))
