"""Diagnostic: time of the per-step pair loops around the hot path on a bench workload (sum density, grad-h correction, dt vote, compatible energy).
python scripts/time_loops.py [workload]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
spec = bench.workload_spec(sys.argv[1] if len(sys.argv) > 1 else "noh8m")
hp = bench.HotPath(spec, spec["n"], 0, 1, 0, None, 0)
e = hp.e
for _ in range(2): hp.step()
e.sync()
def t(name, fn, K=6):
    for _ in range(2): fn()
    e.sync(); t0 = time.perf_counter()
    for _ in range(K): fn()
    e.sync()
    return "%s %.3f" % (name, (time.perf_counter() - t0)/K*1e3)
tag = os.environ.get("SPHB200_LIB", "default").split("_")[-1]
out = [t("energy", lambda: e.update_energy_compatible(1.0e-9)), t("dt", lambda: e.compute_dt(0.25, False)),
       t("omega", lambda: e.compute_omega_gradh()), t("sumrho", lambda: e.sum_mass_density())]
print("[%s %s] ms: %s" % (tag, sys.argv[1] if len(sys.argv) > 1 else "noh8m", "  ".join(out)))
