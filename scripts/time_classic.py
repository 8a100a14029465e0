"""Diagnostic: cost of the classic ASPH ideal H (k_asph_classic) on the 8 M bench workload.  python scripts/time_classic.py [workload]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from spheral_b200 import _lib as L
spec = bench.workload_spec(sys.argv[1] if len(sys.argv) > 1 else "noh8m")
hp = bench.HotPath(spec, spec["n"], 0, 1, 0, None, 0)
e = hp.e
def t(K=6):
    for _ in range(2): hp.step()
    e.sync(); t0 = time.perf_counter()
    for _ in range(K): hp.step()
    e.sync()
    return (time.perf_counter() - t0)/K*1e3
a = t()
e.set_options(hEvolution=L.H_ASPH_CLASSIC, hmin=1.0e-4, hmax=1.0, hminratio=0.1)
b = t()
print("[%s] step with the ASPH tensor derivative only %.3f ms, with the classic ideal H %.3f ms (k_asph_classic %.3f ms)" % (sys.argv[1] if len(sys.argv) > 1 else "noh8m", a, b, b - a))
