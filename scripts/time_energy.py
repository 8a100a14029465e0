"""Diagnostic: time of one compatible-energy update (k_energy_prep + k_energy) on a bench workload.  python scripts/time_energy.py [workload]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
spec = bench.workload_spec(sys.argv[1] if len(sys.argv) > 1 else "noh8m")
hp = bench.HotPath(spec, spec["n"], 0, 1, 0, None, 0)
e = hp.e
for _ in range(2): hp.step()
e.sync()
for _ in range(2): e.update_energy_compatible(1.0e-9)
e.sync(); t0 = time.perf_counter()
K = 10
for _ in range(K): e.update_energy_compatible(1.0e-9)
e.sync()
print("[%s] update_energy_compatible: %.3f ms" % (os.environ.get("SPHB200_LIB", "default").split("_")[-1], (time.perf_counter() - t0)/K*1e3))
