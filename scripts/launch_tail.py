#!/usr/bin/env python
"""Print the last N launches (name, us) of an ncu launch list. usage: launch_tail.py launches.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(k for k, r in enumerate(rows) if r and r[0] == "ID")
ix = {h: i for i, h in enumerate(rows[hi])}
data = rows[hi + 1:]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
for r in data[-n:]:
    v = float(r[ix["Metric Value"]].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ix["Metric Unit"]], 1e-3)
    print("%-70s %9.1f us" % (r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:70], v))
