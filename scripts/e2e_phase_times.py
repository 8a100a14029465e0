"""Diagnostic: phases of one end-to-end step (pinned host buffers in, derivatives out), host wall clock.  python scripts/e2e_phase_times.py [workload] [nside]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
spec = bench.workload_spec(sys.argv[1] if len(sys.argv) > 1 else "noh8m")
n = int(sys.argv[2]) if len(sys.argv) > 2 else spec["n"]
hp = bench.HotPath(spec, n, 0, 1, 0, None, 0)
e = hp.e
for _ in range(3): hp.step_e2e()
e.sync()
def run(sync_between):
    T = {}
    def tick(name, fn):
        t0 = time.perf_counter(); fn()
        if sync_between: e.sync()
        T[name] = T.get(name, 0.0) + time.perf_counter() - t0
    K = 5
    e.sync(); t00 = time.perf_counter()
    for _ in range(K):
        if e.nGhost: e.set_nodes(hp.N, 0)
        tick("upload (9 fields, %.2f GB)" % (hp.h2d/1e9), lambda: e.upload_state_pinned(hp.up_mask, hp.hs))
        tick("plane ghosts", lambda: e.reflect_set_ghost_nodes() if hp.planes else None)
        tick("build_pairs", lambda: e.build_pairs())
        tick("evaluate_derivatives", lambda: e.evaluate_derivatives(0.0, 1.0))
        tick("download (%.2f GB)" % (hp.d2h/1e9), lambda: hp.download())
    e.sync(); tot = time.perf_counter() - t00
    out = {k: round(v/K*1e3, 2) for k, v in T.items()}; out["TOTAL per step"] = round(tot/K*1e3, 2)
    return out
print("synchronised after every phase:", json.dumps(run(True), indent=1))
print("as the bench runs it (asynchronous):", json.dumps(run(False), indent=1))
