#!/usr/bin/env python
"""Extract per-launch DRAM traffic of named kernels from an `ncu --page raw --csv` dump into profiles/traffic.json.
usage: python scripts/ncu_traffic.py <raw.csv> <workload> <source label> [kernel substrings...]"""
import csv, json, os, sys
raw, workload, source = sys.argv[1:4]
want = sys.argv[4:] or ["k_sph_derivs", "k_crk_derivs", "k_nbr_build"]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v*{"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
out = json.load(open(path)) if os.path.exists(path) else {}
ent = out.setdefault(workload, {})
for r in rows[2:]:
    name = r[ix["Kernel Name"]]
    for w in want:
        if w in name:
            ent[w] = {"dram_bytes_read": to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]),
                      "dram_bytes_write": to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]]),
                      "duration_ms_under_ncu": float(r[ix["gpu__time_duration.sum"]].replace(",", ""))*{"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[ix["gpu__time_duration.sum"]]],
                      "kernel": name.split("(")[0], "source": source}
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(ent, indent=1))
