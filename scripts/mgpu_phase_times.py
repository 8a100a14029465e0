"""Diagnostic: where does a decomposed hot-path step spend its time?  Host wall clock around synchronised phases (so the numbers
add up to MORE than an un-synchronised step).  torchrun --nproc-per-node N scripts/mgpu_phase_times.py [nside]"""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
import torch
rank, world, local, dist = bench.dist_setup()
spec = bench.workload_spec("noh8m")
n = int(sys.argv[1]) if len(sys.argv) > 1 else spec["n"]
hp = bench.HotPath(spec, n, rank, world, local, dist, 0)
e, d = hp.e, hp.dsph
for _ in range(3): hp.step()
e.sync()
T = {}
def tick(name, fn):
    e.sync(); torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); e.sync(); torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + time.perf_counter() - t0; return r
K = 5
for _ in range(K):
    nPG = tick("plane ghosts (reflect_set_ghost_nodes)", lambda: e.reflect_set_ghost_nodes())
    if d is not None:
        tick("halo: select + counts + exchange + build_pairs (refresh_ghosts)", lambda: d.refresh_ghosts(build=True, boundary_ghosts=nPG))
    else:
        tick("build_pairs", lambda: e.build_pairs())
    s = e.stats(); T["  of which build_pairs (library events)"] = T.get("  of which build_pairs (library events)", 0.0) + s["ms_build_pairs"]*1e-3
    tick("evaluate_derivatives", lambda: e.evaluate_derivatives(0.0, 1.0))
if rank == 0:
    print(json.dumps({k: round(v/K*1e3, 3) for k, v in T.items()}, indent=1))
# host-side enqueue times of the un-synchronised loop (SPHB200_HALO_TIMING=1)
if d is not None and d._timing:
    d.cpu_ms.clear()
    tp = {}
    e.sync(); t00 = time.perf_counter()
    for _ in range(20):
        t0 = time.perf_counter(); nPG = e.reflect_set_ghost_nodes(); t1 = time.perf_counter()
        d.refresh_ghosts(build=True, boundary_ghosts=nPG); t2 = time.perf_counter()
        e.evaluate_derivatives(0.0, 1.0); t3 = time.perf_counter()
        for k, v in (("plane ghosts (3 syncs; waits for the previous pair kernel)", t1 - t0), ("refresh_ghosts", t2 - t1), ("evaluate_derivatives (enqueue)", t3 - t2)):
            tp[k] = tp.get(k, 0.0) + v*1e3
    e.sync(); tot = (time.perf_counter() - t00)*1e3/20
    if rank == 0:
        print("host timeline, un-synchronised loop, ms per step: total %.3f" % tot)
        print(json.dumps({k: round(v/20, 3) for k, v in tp.items()}, indent=1))
        print(json.dumps({k: round(v/20, 3) for k, v in d.cpu_ms.items()}, indent=1))
if dist is not None: dist.destroy_process_group()
