#!/usr/bin/env python
"""bench.py -- particle-updates/sec of the SPH hot path (neighbour-pair build + evaluateDerivatives) on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload sedov1m|noh8m|glass:<n>:<nbrs>]

A "step" is one pass of the hot path over one synthetic particle set: sphb200_build_pairs (Morton cell sort +
neighbour lists, replacing Neighbor::updateNodes + ConnectivityMap::computeConnectivity) followed by
sphb200_evaluate_derivatives (SPH::evaluateDerivatives + smoothing-scale derivatives).  `value` is measured with the
inputs resident in HBM (CUDA events on the engine's stream, max over ranks); `e2e` repeats the step through the public
API with pinned HOST buffers, i.e. upload of the state and download of the derivatives inside the timed region.

`--impl reference` times the CPU implementation of the same path on the box's host cores.  The real Spheral cannot be
built in this image, so that arm runs the oracle restatement (oracle/sph_oracle.c, OpenMP, the
reference's pair-list + per-thread-scratch strategy) -- "kind": "port".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from spheral_b200 import nodegen as ng  # noqa: E402


# ---------------------------------------------------------------------------------------------------------------------
def workload_spec(name):
    """BASELINE.json configs -> synthetic inputs."""
    if name == "sedov1m":       # configs[1]: Sedov-spherical-3d SPH, 100^3 = 1M particles, BSpline nPerh=1.51 (~115 nbrs)
        return dict(label="Sedov-spherical-3d SPH 100^3 lattice octant (jitter 0.05dx), BSpline table(1000), nPerh=1.51, "
                          "MonaghanGingold Q, compatible energy, XSPH, SPH smoothing scale",
                    n=100, nPerh=1.51, asph=False, kind="sedov")
    if name == "noh8m":         # configs[2]: Noh-spherical-3d ASPH compatible energy, 200^3 = 8M, ~100 nbrs
        return dict(label="Noh-spherical-3d ASPH 200^3 lattice octant (jitter 0.05dx), BSpline table(1000), nPerh=1.44, "
                          "MonaghanGingold Q, compatible energy, XSPH, ASPH smoothing scale",
                    n=200, nPerh=1.44, asph=True, kind="noh")
    if name == "crksph4m":      # configs[3]: CRKSPH Sedov 3-D, 160^3 = 4.1M particles (RK volumes + corrections + the pair loop)
        return dict(label="CRKSPH Sedov-spherical-3d 160^3 lattice octant (jitter 0.05dx), BSpline table(1000), nPerh=1.51, "
                          "LinearOrder RK corrections, RKSumVolume, LimitedMonaghanGingold Q (factory default), compatible energy, "
                          "XSPH, SPH smoothing scale",
                    n=160, nPerh=1.51, asph=False, kind="sedov", hydro="crksph")
    if name.startswith("glass"):  # configs[4]: glass:<n per side>:<neighbours>
        _, n, nb = name.split(":")
        nperh = {32: 1.97, 64: 2.48, 128: 3.13}[int(nb)]/2.0
        return dict(label="synthetic 3-D glass (lattice jitter 0.2dx) %s^3, %s neighbours, SPH" % (n, nb),
                    n=int(n), nPerh=nperh, asph=False, kind="glass")
    raise SystemExit("unknown workload " + name)


def make_inputs(spec, seed=14892042, n_override=None, slab=0):
    """One unit cube of the workload; slab k of a domain-decomposed run is the cube shifted to x in [k, k+1)."""
    n = n_override or spec["n"]
    nPerh = spec["nPerh"]
    pos, mass, H, d = ng.lattice(3, n, nPerh=nPerh)
    N = pos.shape[0]
    jit = 0.2 if spec["kind"] == "glass" else 0.05
    pos = ng.jitter(pos, jit, d, seed=seed)
    pos[:, 0] += float(slab)
    rho = np.ones(N)
    r = np.linalg.norm(pos, axis=1)
    if spec["kind"] == "noh":
        vel = -pos/np.maximum(r, 1e-12)[:, None]
        eps = np.full(N, 1.0e-6)
    else:
        rng = np.random.default_rng(seed + 1)
        vel = 0.1*np.stack([np.sin(3*pos[:, 1]), np.sin(3*pos[:, 2]), np.sin(3*pos[:, 0])], axis=1) + 0.01*rng.standard_normal((N, 3))
        # smoothed energy spike at the origin (Sedov-spherical-3d.py:245-262) on top of a small floor
        h = nPerh*d[0]
        eps = 1.0e-4 + 0.125*np.exp(-(r/(2*h))**2)
    if spec["asph"]:
        # mildly anisotropic H tensors (an evolved ASPH state): compress radially by up to 30 %
        rng = np.random.default_rng(seed + 2)
        F = ng.sym_to_full(3, H)
        rh = pos/np.maximum(r, 1e-12)[:, None]
        s = 1.0 + 0.3*rng.uniform(size=N)
        F = F + (s - 1.0)[:, None, None]*F[:, :1, :1]*np.einsum("na,nb->nab", rh, rh)
        H = ng.full_to_sym(3, F)
    P, cs = ng.gamma_law(rho, eps)
    st = dict(position=pos, velocity=vel, H=H, mass=mass, massDensity=rho, specificThermalEnergy=eps, pressure=P,
              soundSpeed=cs, omegaGradh=np.ones(N))
    if spec.get("hydro") == "crksph":
        # LimitedMonaghanGingold reads the velocity gradient of the previous evaluation: the analytic gradient of the field above
        g = np.zeros((N, 3, 3))
        g[:, 0, 1] = 0.3*np.cos(3*pos[:, 1]); g[:, 1, 2] = 0.3*np.cos(3*pos[:, 2]); g[:, 2, 0] = 0.3*np.cos(3*pos[:, 0])
        st["DvDxQ"] = g.reshape(N, 9)
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in st.items()}, N


def options_kwargs(spec):
    if spec.get("hydro") == "crksph":      # CRKSPHHydros.py:64-68 defaults: Cl = 2(kext/4), Cq = (kext/4)^2 with kext = 2
        return dict(nPerh=spec["nPerh"], compatibleEnergy=1, XSPH=1, Qkind=1, Cl=1.0, Cq=0.25, hEvolution=0)
    return dict(nPerh=spec["nPerh"], compatibleEnergy=1, XSPH=1, correctVelocityGradient=1, Qkind=0, Cl=2.0, Cq=2.0,
                hEvolution=1 if spec["asph"] else 0)


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(ngpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local)
        dist_.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        dist = dist_
    return rank, world, local, dist


def max_over_ranks(dist, x, local):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(dist, local):
    if dist is not None:
        import torch
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()


# ---------------------------------------------------------------------------------------------------------------------
def cpu_port_run(spec, sample_n, steps, warmup, threads):
    """The reference arm: oracle restatement on the host cores.  Returns particle-updates/s and a description."""
    from oracle import oracle as orc
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common
    from spheral_b200 import kernel as K
    orc.build()
    st, N = make_inputs(spec, n_override=sample_n)
    WT = K.TableKernel(K.BSplineKernel(3), 1000)
    OT = common.oracle_table(orc, WT)
    oo = orc.default_options(3, **options_kwargs(spec))
    s = common.to_oracle_state(st)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    times = []
    crk = spec.get("hydro") == "crksph"
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        pi, pj, cnt = orc.pairs(3, N, 0, s["pos"], s["H"], WT.kernelExtent)
        if crk:
            vol = orc.crk_sum_volume(3, OT, N, 0, s["pos"], s["H"], pi, pj)
            corr = orc.crk_corrections(3, OT, N, 0, s["pos"], s["H"], vol, pi, pj)
            orc.crk_evaluate_derivatives(oo, OT, s, vol, corr, N, 0, pi, pj)
        else:
            orc.evaluate_derivatives(oo, OT, s, N, 0, pi, pj, cnt, nthreads=threads)
        t1 = time.perf_counter()
        if it >= warmup:
            times.append(t1 - t0)
    tmed = float(np.median(times))
    return N/tmed, tmed, N, len(pi), ("%d^3 = %d particles of the same workload (%.1f neighbours/particle), pair build + "
                                      "evaluateDerivatives, median of %d" % (sample_n, N, 2.0*len(pi)/N, len(times)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sedov1m")
    ap.add_argument("--n", type=int, default=0, help="override lattice points per side (debug)")
    ap.add_argument("--cpu-sample", type=int, default=48, help="lattice points per side of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="A/B diagnostic: skip the device-resident RK2 leg (scripts/gpu_ab.sh)")
    ap.add_argument("--rk2-trace", action="store_true", help="diagnostic: per-step breakdown of the device-resident RK2 leg on stderr")
    ap.add_argument("--hjitter", type=float, default=0.0,
                    help="scale every node's H by a random factor in [1-x, 1+x] (diagnostic: the lattice workloads have a constant h)")
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    spec = workload_spec(args.workload)
    if args.n:
        spec = dict(spec, n=args.n, label=spec["label"] + " [n overridden to %d]" % args.n)
    threads = os.cpu_count() or 1
    metric = "particle-updates/sec (3D SPH derivs+neighbour)"
    config = {"workload": spec["label"], "particles_per_gpu": spec["n"]**3, "dim": 3,
              "decomposition": "1-D slabs along x, one unit cube of the named size per GPU, ghost halo exchanged over NCCL every step (weak)",
              "l2": "per-step working set (node rows + neighbour lists + pair accelerations) exceeds the 126 MB L2"}

    # ------------------------------------------------------------------ reference arm (CPU) ----------------------------
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return 0
        val, tmed, N, npairs, sample = cpu_port_run(spec, args.cpu_sample, max(args.steps, 1), min(warmup, 1), threads)
        line = {"impl": "reference", "metric": metric, "value": val, "unit": "particle-updates/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": tmed*1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "particle-updates/s", "cores": threads, "kind": "port", "sample": sample,
                                 "note": "oracle restatement on host cores -- not the Spheral MPI build (unbuildable here)"},
                "e2e": {"value": val, "unit": "particle-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (GPU) ---------------------------------------
    rank, world, local, dist = dist_setup(args.gpus)
    import torch
    from spheral_b200 import _lib as L, engine, kernel as K
    import ctypes as C

    st, N = make_inputs(spec, seed=14892042 + rank, slab=rank)
    if args.hjitter:
        st["H"] = st["H"]*(1.0 + args.hjitter*np.random.default_rng(5 + rank).uniform(-1.0, 1.0, size=(N, 1)))
        config["workload"] += " [h jittered by +-%g]" % args.hjitter
    WT = K.TableKernel(K.BSplineKernel(3), 1000)
    crk = spec.get("hydro") == "crksph"
    if crk and world > 1:
        raise SystemExit("the CRKSPH workload is single-GPU in this round (the halo plumbing exchanges the SPH fields only)")
    e = engine.Engine(3, device=local, hydro=(L.HYDRO_CRKSPH if crk else L.HYDRO_SPH), **options_kwargs(spec))
    e.set_kernel_table(WT)
    e.set_nodes(N, 0)
    ext = torch.cuda.ExternalStream(e.stream, device=torch.device("cuda", local))

    # pinned host buffers for the e2e leg (internal nodes only: ghosts arrive over NVLink, not from the host)
    up_names = ("position", "velocity", "H", "mass", "massDensity", "specificThermalEnergy", "pressure", "soundSpeed", "omegaGradh")
    if crk:
        up_names = up_names[:-1] + ("DvDxQ",)
    down_names = ("DxDt", "DrhoDt", "DvDt", "DepsDt", "DvDx", "DHDt", "Hideal")
    hs = L.HostState()
    pinned, up_mask, h2d = [], 0, 0
    dp = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
    for k in up_names:
        t = torch.from_numpy(st[k]).clone().pin_memory()
        pinned.append(t); setattr(hs, k, dp(t)); up_mask |= L.STATE_BITS[k]; h2d += t.numel()*8
    down_mask = 0
    for k in down_names:
        down_mask |= L.DERIV_BITS[k]
    d2h = sum(N*L.deriv_width(3, k)*8 for k in down_names)
    down_bufs = {}

    def download():
        # destination sized for internal + current ghosts (the C ABI writes every node; ghost entries are zeros)
        n = e.nInternal + e.nGhost
        if down_bufs.get("n", 0) < n:
            hd = L.HostDerivs()
            keep = []
            for k in down_names:
                t = torch.empty((n + n//16)*L.deriv_width(3, k), dtype=torch.float64).pin_memory()
                keep.append(t); setattr(hd, k, dp(t))
            down_bufs.update(n=n + n//16, hd=hd, keep=keep)
        e._check(e._lib.sphb200_download_derivs(e._h, down_mask, C.byref(down_bufs["hd"])))

    e.upload_state_pinned(up_mask, hs)
    e.sync()

    dsph = None
    if world > 1:
        from spheral_b200 import distributed as D
        dsph = D.DistributedSPH(e, 0, float(rank), float(rank + 1))

    def step():
        if dsph is not None:
            dsph.step_connectivity_and_derivatives(0.0, 1.0)     # ghost selection + NVLink halo exchange + K1..K5
        else:
            e.build_pairs()
            if crk:                                              # RKCorrections::preStepInitialize / initialize
                e.crk_compute_volume()
                e.crk_compute_corrections()
            e.evaluate_derivatives(0.0, 1.0)

    def step_e2e():
        if dsph is not None:
            e.set_nodes(N, 0)
        e.upload_state_pinned(up_mask, hs)
        step()
        download()

    def timed(fn, k):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(dist, local)
        t0 = time.perf_counter()
        ev0.record(ext)
        for _ in range(k):
            fn()
        ev1.record(ext)
        e.sync()
        wall = time.perf_counter() - t0
        barrier(dist, local)
        return ev0.elapsed_time(ev1)*1e-3, wall

    for _ in range(warmup):
        step()
    e.sync()
    fp64_peak = e.measure_fp64_peak()
    launches0 = e.stats()["launches"]

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_s, wall = timed(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = e.stats()["launches"] - launches0
    dev_s = max_over_ranks(dist, dev_s, local)
    wall_s = max_over_ranks(dist, wall, local)

    # per-kernel breakdown (CUDA events inside the library, on the same stream), outside the timed region
    pair_ms, nbr_ms, build_ms, eval_ms = [], [], [], []
    for _ in range(3):
        step()
        s_ = e.stats()
        pair_ms.append(s_["ms_pair_kernel"]); nbr_ms.append(s_["ms_neighbor_kernels"])
        build_ms.append(s_["ms_build_pairs"]); eval_ms.append(s_["ms_evaluate"])
    edges = e.stats()["directed_edges"]
    stencil_radius = e.stats()["stencil_radius"]
    halo_info = dsph.info() if dsph is not None else None

    # device-resident CheapSynchronousRK2 steps: neighbour update + sum density + dt vote + trial advance +
    # grad-h correction + derivatives + compatible-energy update + full advance, no field leaving the GPU
    rk2 = None
    if dsph is None and not args.quick:
        from spheral_b200 import integrator as I
        rk = I.CheapSynchronousRK2(e, engine.make_step_options())
        rk.initializeDerivatives()
        for _ in range(2):
            rk.step()
        e.sync()
        nrk = max(3, args.steps//2)
        if args.rk2_trace and rank == 0:
            import time as _t
            for q in range(nrk):
                e.sync(); t0 = _t.perf_counter(); l0 = e.stats()["launches"]
                rk.step()
                e.sync(); t1 = _t.perf_counter(); s_ = e.stats()
                sys.stderr.write("rk2 step %2d: wall %.2f ms  build %.2f  nbr %.2f  eval %.2f  pair %.2f  energy %.2f  launches %d  radius %d  edges/node %.1f\n"
                                 % (q, (t1 - t0)*1e3, s_["ms_build_pairs"], s_["ms_neighbor_kernels"], s_["ms_evaluate"], s_["ms_pair_kernel"],
                                    s_["ms_energy"], s_["launches"] - l0, s_["stencil_radius"], s_["directed_edges"]/float(N)))
        rk_s, _ = timed(rk.step, nrk)
        rk2 = {"ms_per_step": rk_s/nrk*1e3, "value": N*nrk/rk_s, "unit": "particle-updates/s", "steps": nrk,
               "what": ("CheapSynchronousRK2 step with the state resident in HBM: build_pairs + computeRKSumVolume + computeCRKSPHSumMassDensity + "
                        "RK corrections (x2) + GenericHydro::dt + State::update (x2) + evaluateDerivatives + compatible energy; "
                        "one 16-byte read-back (dt) per step" if crk else
                        "CheapSynchronousRK2 step with the state resident in HBM: build_pairs + computeSPHSumMassDensity + GenericHydro::dt + "
                        "State::update (x2) + computeSPHOmegaGradhCorrection (x2) + evaluateDerivatives + compatible energy; "
                        "one 16-byte read-back (dt) per step"),
               "last_dt": rk.lastDt, "dt_reason": rk.lastDtReason}
        # put the bench state back (the e2e leg uploads it anyway)
        e.upload_state_pinned(up_mask, hs)
        e.sync()

    # e2e leg: host buffers in, host buffers out, every step
    for _ in range(2):
        step_e2e()
    e.sync()
    _, e2e_wall = timed(step_e2e, args.steps)
    e2e_s = max_over_ranks(dist, e2e_wall, local)

    if rank == 0:
        nbrs = edges/float(N)
        total_updates = float(N)*world*args.steps
        value = total_updates/dev_s
        # roofline of the dominant kernel (k_sph_derivs): algorithmic work per launch
        t_pair = float(np.mean(pair_ms))*1e-3
        bytes_alg = N*(672.0 + 12.0*nbrs)            # API-faithful variant: state in, derivatives out, 24 B per pair
        flops_alg = N*250.0*nbrs                     # ~500 flop per pair, each pair counted once
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        if crk:                                      # + volume and 16 RK coefficients per node in; ~560 flop per pair
            bytes_alg = N*(672.0 + 136.0 + 12.0*nbrs)
            flops_alg = N*280.0*nbrs
        # DRAM bytes the dominant kernel really moved per launch: from the committed `ncu --set full` capture of this workload
        # (profiles/traffic.json, written by scripts/ncu_traffic.py); null for workloads / sizes without a capture
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tj.get(args.workload if not args.n and not args.hjitter else "", {})
            key = "k_crk_derivs" if crk else "k_sph_derivs"
            if key in ent:
                traffic = float(ent[key]["dram_bytes_read"]) + float(ent[key]["dram_bytes_write"])
                traffic_src = ent[key].get("source")
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": ("k_crk_derivs (CRKSPH pair loop + finalize + smoothing scale)" if crk else
                                               "k_sph_derivs (SPH pair loop + finalize + smoothing scale)"),
                    "achieved": bytes_alg/t_pair/1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": bytes_alg/t_pair/1e9/hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": hbm_src,
                    "algorithmic_bytes_per_particle": bytes_alg/N,
                    "note": "HBM is not the binding roof at ~100 neighbours (arithmetic intensity ~13 flop/B vs machine balance ~5): the FP64 pipe is (roofline_fp64), and the unit the kernel actually saturates is the LSU data pipe of the L1 (neighbour-row gather + table look-ups: 81 % of peak wavefronts, FP64 pipe 38 %, profiles/r01_ncu_full_end_of_round_excerpt.csv)"}
        roofline_fp64 = {"bound": "fp64", "achieved": flops_alg/t_pair/1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": flops_alg/t_pair/1e12/fp64_peak, "algorithmic_flops_per_particle": flops_alg/N,
                         "peak_source": "DFMA microbenchmark run in this process (sphb200_measure_fp64_peak)"}
        line = {"metric": metric, "value": value, "unit": "particle-updates/s", "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": dev_s/args.steps*1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(config, neighbours_per_particle=nbrs, halo=halo_info, grid_stencil_radius=int(stencil_radius), grid_fine_walk=int(e.stats().get("fine_walk", 0)),
                               timing="CUDA events on the engine stream around the K steps, max over ranks"),
                "clocks": clocks,
                "e2e": {"value": total_updates/e2e_s, "unit": "particle-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "timing": "host wall clock around synchronised C-ABI calls (pinned host buffers)"},
                "gpu_launches": int(launches),
                "breakdown_ms": {"build_pairs": float(np.mean(build_ms)), "neighbor_kernels": float(np.mean(nbr_ms)),
                                 "evaluate": float(np.mean(eval_ms)), "pair_kernel": float(np.mean(pair_ms)),
                                 "wall_per_step": wall_s/args.steps*1e3},
                "roofline": roofline, "roofline_fp64": roofline_fp64}
        if rk2 is not None:
            line["rk2_step_resident"] = rk2
        if not args.no_cpu_baseline and world == 1:
            val, tmed, Ns, npairs, sample = cpu_port_run(spec, args.cpu_sample, 3, 1, threads)
            line["cpu_baseline"] = {"value": val, "unit": "particle-updates/s", "cores": threads, "kind": "port", "sample": sample,
                                    "note": "oracle restatement on host cores -- not the Spheral MPI build (unbuildable here)"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
