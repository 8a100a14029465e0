#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: stall-reason totals and the hottest SASS instructions.
usage: ncu -i prof.ncu-rep --page source --csv --launch-skip K --launch-count 1 > src.csv ; python scripts/ncu_stalls.py src.csv [ntop]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = next(k for k, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
print(rows[0][:2])
S = ix["# Samples"]
tot = sum(int(r[S] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
    print("  %-24s %8d %5.1f%%" % (s, v, 100.0*v/max(tot, 1)))
top = sorted(range(len(data)), key=lambda k: -int(data[k][S] or 0))[:ntop]
for k in sorted(top):
    r = data[k]
    st = {s: int(r[ix[s]] or 0) for s in stalls}
    b = max(st, key=st.get)
    print("%5d %-72s %6s %s=%d" % (k, r[ix["Source"]].strip()[:72], r[S], b, st[b]))
