#!/usr/bin/env python
"""bench.py -- particle-updates/sec of the SPH hot path (neighbour-pair build + evaluateDerivatives) on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload noh8m|sedov1m|crksph4m|glass:<n>:<nbrs>]

A "step" is one pass of the hot path over one synthetic particle set: ghost generation for the reflecting planes of the
workload (Integrator::setGhostNodes), sphb200_build_pairs (Morton cell sort + neighbour lists, replacing
Neighbor::updateNodes + ConnectivityMap::computeConnectivity) and sphb200_evaluate_derivatives (SPH::evaluateDerivatives +
the smoothing-scale derivatives).  `value` is measured with the inputs resident in HBM (CUDA events on the engine's
stream, max over ranks); `e2e` repeats the step through the public API with pinned HOST buffers, i.e. upload of the state
and download of the derivatives inside the timed region.

Default workload = BASELINE.json configs[2]: Noh-spherical-3d ASPH with compatible energy, 200^3 = 8 M particles, the
octant's three reflecting planes, options as the stock script runs it (tests/functional/Hydro/Noh/Noh-spherical-3d.py:
XSPH False, compatibleEnergy True, correctVelocityGradient True; tests/performance.py:152-251 sizes it as cbrt(Ntotal)^3).
`--gpus N` splits THAT problem into N slabs along x (strong scaling): every rank owns the lattice planes ix in
[r n/N, (r+1) n/N), ghosts of the neighbouring slabs arrive over NCCL, plane ghosts are generated on the device.  A
weak-scaling line (one 100^3 cube per GPU, the round-1 measurement) is added under "weak" when N > 1.

`--impl reference` times the CPU implementation of the same path on the box's host cores.  The real Spheral cannot be
built in this image, so that arm runs the oracle restatement (oracle/sph_oracle.c, OpenMP, the reference's
pair-list + per-thread-scratch strategy) -- "kind": "port" -- on a bounded sample of the same workload; it never loads
libsphb200.so.
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

METRIC = "particle-updates/sec (3D SPH derivs+neighbour)"
UNIT = "particle-updates/s"


# ---------------------------------------------------------------------------------------------------------------------
# Host-side input generation (numpy only: shared by both arms, so the reference arm never imports the product)
def _sym_from_diag3(diag):
    H = np.zeros((diag.shape[0], 6))
    H[:, 0], H[:, 3], H[:, 5] = diag[:, 0], diag[:, 1], diag[:, 2]
    return H


def workload_spec(name):
    """BASELINE.json configs -> synthetic inputs."""
    if name == "noh8m":         # configs[2]: Noh-spherical-3d ASPH compatible energy, 200^3 = 8M, ~100 nbrs, 3 reflecting planes
        return dict(name=name, label="Noh-spherical-3d ASPH 200^3 lattice octant (jitter 0.05dx), reflecting planes x=y=z=0, BSpline "
                          "table(1000), nPerh=1.44, MonaghanGingold Q, compatible energy, XSPH off (stock script), ASPH smoothing scale",
                    n=200, nPerh=1.44, asph=True, kind="noh", planes=True, cpu_sample=80, parity_sample=40)
    if name == "sedov1m":       # configs[1]: Sedov-spherical-3d SPH, 100^3 = 1M particles, BSpline nPerh=1.51 (~115 nbrs)
        return dict(name=name, label="Sedov-spherical-3d SPH 100^3 lattice (jitter 0.05dx), BSpline table(1000), nPerh=1.51, "
                          "MonaghanGingold Q, compatible energy, XSPH off (stock script), SPH smoothing scale",
                    n=100, nPerh=1.51, asph=False, kind="sedov", planes=False, cpu_sample=64, parity_sample=40)
    if name == "crksph4m":      # configs[3]: CRKSPH Sedov 3-D, 160^3 = 4.1M particles (RK volumes + corrections + the pair loop)
        return dict(name=name, label="CRKSPH Sedov-spherical-3d 160^3 lattice (jitter 0.05dx), BSpline table(1000), nPerh=1.51, "
                          "LinearOrder RK corrections, RKSumVolume, LimitedMonaghanGingold Q (factory default), compatible energy, "
                          "XSPH off (stock script), SPH smoothing scale",
                    n=160, nPerh=1.51, asph=False, kind="sedov", hydro="crksph", planes=False, cpu_sample=48, parity_sample=32)
    if name.startswith("glass"):  # configs[4]: glass:<n per side>:<neighbours>
        _, n, nb = name.split(":")
        nperh = {32: 1.97, 64: 2.48, 128: 3.13}[int(nb)]/2.0
        return dict(name=name, label="synthetic 3-D glass (lattice jitter 0.2dx) %s^3, %s neighbours, SPH" % (n, nb),
                    n=int(n), nPerh=nperh, asph=False, kind="glass", planes=False, cpu_sample=48, parity_sample=32)
    raise SystemExit("unknown workload " + name)


def make_inputs(spec, seed=14892042, n=None, slab=(0, 1), shift=0.0):
    """The workload's particle set on the unit cube, n^3 lattice (GenerateNodeDistribution3d.py:585-622 restated: x = (i + 1/2)/n,
    m = rho0/n^3, H = I/(nPerh dx)), node order x fastest.  slab = (r, R): only the lattice planes ix in [r n/R, (r+1) n/R) are
    returned -- every random number is drawn for the whole lattice first, so the union of the slabs IS the single-GPU problem.
    shift: the cube is moved by `shift` along x (weak-scaling runs put one cube per rank side by side)."""
    n = n or spec["n"]
    nPerh = spec["nPerh"]
    d = 1.0/n
    ax = (np.arange(n) + 0.5)*d
    Z, Y, X = np.meshgrid(ax, ax, ax, indexing="ij")
    N = n**3
    jit = 0.2 if spec["kind"] == "glass" else 0.05
    rng = np.random.default_rng(seed)
    J = rng.uniform(-1.0, 1.0, size=(N, 3))
    r, R = slab
    ix = np.tile(np.arange(n), n*n)
    sel = np.nonzero((ix >= (r*n)//R) & (ix < ((r + 1)*n)//R))[0] if R > 1 else slice(None)
    pos = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)[sel] + jit*d*J[sel]
    del J, X, Y, Z
    M = pos.shape[0]
    mass = np.full(M, d**3)
    H = _sym_from_diag3(np.full((M, 3), 1.0/(nPerh*d)))
    rho = np.ones(M)
    rr = np.linalg.norm(pos, axis=1)
    if spec["kind"] == "noh":
        vel = -pos/np.maximum(rr, 1e-12)[:, None]
        eps = np.full(M, 1.0e-6)
    else:
        rng1 = np.random.default_rng(seed + 1)
        noise = rng1.standard_normal((N, 3))[sel]
        vel = 0.1*np.stack([np.sin(3*pos[:, 1]), np.sin(3*pos[:, 2]), np.sin(3*pos[:, 0])], axis=1) + 0.01*noise
        # smoothed energy spike at the origin (Sedov-spherical-3d.py:245-262) on top of a small floor
        h = nPerh*d
        eps = 1.0e-4 + 0.125*np.exp(-(rr/(2*h))**2)
    if spec["asph"]:
        # mildly anisotropic H tensors (an evolved ASPH state): compress radially by up to 30 %
        rng2 = np.random.default_rng(seed + 2)
        s = (1.0 + 0.3*rng2.uniform(size=N))[sel]
        rh = pos/np.maximum(rr, 1e-12)[:, None]
        F = np.zeros((M, 3, 3))
        F[:, 0, 0] = F[:, 1, 1] = F[:, 2, 2] = 1.0/(nPerh*d)
        F = F + ((s - 1.0)/(nPerh*d))[:, None, None]*np.einsum("na,nb->nab", rh, rh)
        H = np.stack([F[:, 0, 0], F[:, 0, 1], F[:, 0, 2], F[:, 1, 1], F[:, 1, 2], F[:, 2, 2]], axis=1)
    gamma = 5.0/3.0                                      # GammaLawGas.cc:185-243
    P = (gamma - 1.0)*rho*eps
    cs = np.sqrt(np.maximum(0.0, gamma*(gamma - 1.0)*eps))
    pos[:, 0] += float(shift)
    st = dict(position=pos, velocity=vel, H=H, mass=mass, massDensity=rho, specificThermalEnergy=eps, pressure=P,
              soundSpeed=cs, omegaGradh=np.ones(M))
    if spec.get("hydro") == "crksph":
        # LimitedMonaghanGingold reads the velocity gradient of the previous evaluation: the analytic gradient of the field above
        g = np.zeros((M, 3, 3))
        g[:, 0, 1] = 0.3*np.cos(3*pos[:, 1]); g[:, 1, 2] = 0.3*np.cos(3*pos[:, 2]); g[:, 2, 0] = 0.3*np.cos(3*pos[:, 0])
        st["DvDxQ"] = g.reshape(M, 9)
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in st.items()}, M


def plane_list(spec, shift=0.0):
    """(point, inward normal) of the reflecting planes of the workload (Noh-spherical-3d.py:385-396: x = 0, y = 0, z = 0)."""
    if not spec.get("planes"):
        return []
    return [(np.array([float(shift), 0.0, 0.0]), np.array([1.0, 0.0, 0.0])),
            (np.array([0.0, 0.0, 0.0]), np.array([0.0, 1.0, 0.0])),
            (np.array([0.0, 0.0, 0.0]), np.array([0.0, 0.0, 1.0]))]


def options_kwargs(spec, xsph=0):
    if spec.get("hydro") == "crksph":      # CRKSPHHydros.py:64-68 defaults: Cl = 2(kext/4), Cq = (kext/4)^2 with kext = 2
        return dict(nPerh=spec["nPerh"], compatibleEnergy=1, XSPH=xsph, Qkind=1, Cl=1.0, Cq=0.25, hEvolution=0)
    return dict(nPerh=spec["nPerh"], compatibleEnergy=1, XSPH=xsph, correctVelocityGradient=1, Qkind=0, Cl=2.0, Cq=2.0,
                hEvolution=1 if spec["asph"] else 0)


def bench_config(spec, world, n):
    """The SAME dictionary on both arms (the driver compares them)."""
    return {"workload": spec["label"] + ("" if n == spec["n"] else " [n overridden to %d]" % n), "particles": n**3, "dim": 3, "n_gpus": world,
            "decomposition": ("single GPU" if world == 1 else
                              "strong: the %d^3 problem cut into %d slabs along x, slab halo over NCCL send/recv every step, plane ghosts on the device" % (n, world)),
            "l2": "per-step working set (node rows + neighbour lists + pair accelerations, > 10 GB at 8 M) exceeds the 126 MB L2",
            "timing": "CUDA events on the engine stream around the K steps, max over ranks (reference arm: host steady clock)"}


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


def claim_stdout():
    """stdout carries the ONE JSON line and nothing else: file descriptor 1 is pointed at stderr for the duration of the run (libraries
    that write to it directly -- NCCL's version banner, a compiler invoked by a build step -- land in the log) and the returned file
    object, a duplicate of the original descriptor, receives the line at the end."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def dist_setup():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        # keep NCCL's own account of the communicator (ranks, transport: P2P / NVLS over NVLink) in the run's stderr
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ["NCCL_DEBUG_SUBSYS"] = "INIT"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")          # stdout carries the JSON line only
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local)
        dist_.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        dist = dist_
    return rank, world, local, dist


def reduce_over_ranks(dist, x, local, op="max"):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM}[op])
    return float(t.item())


def barrier(dist, local):
    if dist is not None:
        import torch
        dist.barrier(device_ids=[local])
        torch.cuda.synchronize()


def pin_to_gpu_numa(local):
    """Bind this rank's host threads to the CPUs next to its GPU (pinned buffers are then first-touched on that NUMA node)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------------------
# The checker / CPU arm: oracle restatement on the host cores (test infrastructure; never on the product path)
def oracle_problem(spec, sample_n, xsph):
    """Inputs of the bounded CPU sample: the same workload on a sample_n^3 lattice, plane ghosts by the host restatement."""
    from oracle import oracle as orc
    from spheral_b200 import nodegen as og          # numpy-only restatement of the reflecting-plane ghosts; does not load libsphb200.so
    orc.build()
    st, N = make_inputs(spec, n=sample_n)
    OT = orc.TableKernel(orc.KERNEL_BSPLINE, 3, 1000)
    oo = orc.default_options(3, **options_kwargs(spec, xsph))
    s = dict(pos=st["position"], vel=st["velocity"], H=st["H"], mass=st["mass"], rho=st["massDensity"], P=st["pressure"],
             cs=st["soundSpeed"], omega=st["omegaGradh"], eps=st["specificThermalEnergy"])
    if "DvDxQ" in st:
        s["DvDxQ"] = st["DvDxQ"]
    return orc, og, OT, oo, s, N


def host_ghosts(og, s0, planes, kext):
    """Reflecting-plane ghosts of the sample by the host restatement (nodegen.reflect_ghosts): (fields incl. ghosts, nGhost)."""
    if not planes:
        return s0, 0
    out, _, n0 = og.reflect_ghosts(3, s0, planes, kext)
    return {k: np.ascontiguousarray(v) for k, v in out.items()}, out["pos"].shape[0] - n0


def cpu_port_run(spec, sample_n, steps, warmup, threads, xsph):
    """The reference arm: oracle restatement on the host cores.  Returns particle-updates/s and a description."""
    orc, og, OT, oo, s0, N = oracle_problem(spec, sample_n, xsph)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    planes = plane_list(spec)
    times, npairs = [], 0
    crk = spec.get("hydro") == "crksph"
    # Integrator::setGhostNodes: done once, outside the timed loop (the numpy restatement of the plane ghosts is far slower than the
    # reference's C++ would be; leaving it out can only favour the CPU arm -- the GPU arm regenerates its ghosts every step)
    s, nG = host_ghosts(og, s0, planes, OT.kext)
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        pi, pj, cnt = orc.pairs(3, N, nG, s["pos"], s["H"], OT.kext)
        if crk:
            vol = orc.crk_sum_volume(3, OT, N, nG, s["pos"], s["H"], pi, pj)
            corr = orc.crk_corrections(3, OT, N, nG, s["pos"], s["H"], vol, pi, pj)
            orc.crk_evaluate_derivatives(oo, OT, s, vol, corr, N, nG, pi, pj)
        else:
            orc.evaluate_derivatives(oo, OT, s, N, nG, pi, pj, cnt, nthreads=threads)
        t1 = time.perf_counter()
        npairs = len(pi)
        if it >= warmup:
            times.append(t1 - t0)
    tmed = float(np.median(times))
    sample = ("%d^3 = %d particles (+%d plane ghosts) of the same workload (%.1f neighbours/particle), pair build + "
              "evaluateDerivatives per step, median of %d" % (sample_n, N, nG, float(np.mean(cnt[:N])), len(times)))
    return N/tmed, tmed, N, npairs, sample


def parity_check(spec, sample_n, xsph, device):
    """GPU vs oracle on a bounded sample of the bench workload: pair sets memcmp-equal, worst field error (SURVEY 8c metric)."""
    import ctypes as C
    from spheral_b200 import _lib as L, engine
    orc, og, OT, oo, s0, N = oracle_problem(spec, sample_n, xsph)
    planes = plane_list(spec)
    s, nG = host_ghosts(og, s0, planes, OT.kext)
    pi, pj, cnt = orc.pairs(3, N, nG, s["pos"], s["H"], OT.kext)
    crk = spec.get("hydro") == "crksph"
    e = engine.Engine(3, device=device, hydro=(L.HYDRO_CRKSPH if crk else L.HYDRO_SPH), **options_kwargs(spec, xsph))
    e.set_kernel_table(OracleTableView(OT))                # ONE table on both sides (the oracle's, uploaded through the C ABI)
    e.set_nodes(N, nG)
    names = dict(pos="position", vel="velocity", H="H", mass="mass", rho="massDensity", P="pressure", cs="soundSpeed", omega="omegaGradh",
                 eps="specificThermalEnergy", DvDxQ="DvDxQ")
    e.upload_state(**{names[k]: v for k, v in s.items()})
    npairs = e.build_pairs()
    gi, gj = e.download_pairs()
    pairs_equal = bool(npairs == len(pi) and np.array_equal(gi, pi) and np.array_equal(gj, pj))
    if crk:
        vol = orc.crk_sum_volume(3, OT, N, nG, s["pos"], s["H"], pi, pj)
        corr = orc.crk_corrections(3, OT, N, nG, s["pos"], s["H"], vol, pi, pj)
        ref = orc.crk_evaluate_derivatives(oo, OT, s, vol, corr, N, nG, pi, pj)
        e.crk_compute_volume(); e.crk_compute_corrections()
    else:
        ref = orc.evaluate_derivatives(oo, OT, s, N, nG, pi, pj, cnt)
    e.evaluate_derivatives(0.0, 1.0)
    got = e.download_derivs()
    h = 1.0/s["H"][:N, 0].mean()
    csm = max(float(s["cs"][:N].max()), 1e-30)
    v = max(float(np.abs(s["vel"][:N]).max()), csm)
    floors = dict(DxDt=v, DrhoDt=v/h, DvDt=csm*csm/h, DepsDt=csm*csm*v/h, DvDx=v/h, DHDt=v/(h*h), rhoSum=1.0, normalization=1.0,
                  gradRho=1.0/h, M=1.0, maxViscousPressure=max(float(np.abs(s["P"][:N]).max()), 1e-30))
    worst, where = 0.0, ""
    for k, f in floors.items():
        if k not in ref or k not in got:
            continue
        a, b = np.asarray(got[k])[:N], np.asarray(ref[k])[:N]
        err = float(np.abs(a - b).max())/max(float(np.abs(b).max()), f)
        if err > worst:
            worst, where = err, k
    e.close()
    return {"sample": "%d^3 = %d particles (+%d plane ghosts) of the bench workload, GPU (C ABI) vs oracle" % (sample_n, N, nG),
            "pairs": int(len(pi)), "pairs_equal": pairs_equal, "worst_field_err": worst, "worst_field": where, "tolerance": 1.0e-10,
            "ok": bool(pairs_equal and worst <= 1.0e-10)}


class OracleTableView:
    """The oracle's TableKernel presented with the attribute names Engine.set_kernel_table reads."""

    def __init__(self, OT):
        self.ndim, self.kernelExtent, self.xmin, self.xstep, self.n1 = OT.ndim, OT.kext, OT.xmin, OT.xstep, OT.n1
        self.Wcoef, self.gradWcoef, self.grad2Wcoef = OT.Wcoef, OT.gradWcoef, OT.grad2Wcoef
        self.nperhVals, self.nperhRange, self.wsumVals, self.wsumRange = OT.nperhVals, OT.nperhRange, OT.wsumVals, OT.wsumRange


# ---------------------------------------------------------------------------------------------------------------------
class HotPath:
    """One rank's share of a workload on its GPU: engine, pinned host buffers, the step and its end-to-end variant."""

    def __init__(self, spec, n, rank, world, local, dist, xsph, weak=False):
        import ctypes as C
        import torch
        from spheral_b200 import _lib as L, engine, kernel as K
        self.torch, self.L, self.C = torch, L, C
        self.spec, self.rank, self.world, self.local = spec, rank, world, local
        crk = self.crk = spec.get("hydro") == "crksph"
        if weak:      # one full cube per rank, side by side along x (round-1 measurement)
            st, N = make_inputs(spec, seed=14892042 + rank, n=n, shift=float(rank))
            self.lo, self.hi = float(rank), float(rank + 1)
            # a plane at x = rank would mirror the neighbouring cube: weak cubes keep only the y and z planes
            self.planes = plane_list(spec)[1:] if spec.get("planes") else []
        else:
            st, N = make_inputs(spec, n=n, slab=(rank, world))
            self.lo, self.hi = ((rank*n)//world)/float(n), (((rank + 1)*n)//world)/float(n)
            self.planes = plane_list(spec)
        self.N = N
        e = self.e = engine.Engine(3, device=local, hydro=(L.HYDRO_CRKSPH if crk else L.HYDRO_SPH), **options_kwargs(spec, xsph))
        e.set_kernel_table(K.TableKernel(K.BSplineKernel(3), 1000))
        e.set_nodes(N, 0)
        if self.planes:
            e.reflect_configure(self.planes)
        self.ext = torch.cuda.ExternalStream(e.stream, device=torch.device("cuda", local))
        # pinned host buffers for the e2e leg (internal nodes only: ghosts come from the planes / over NVLink, not from the host)
        up = ["position", "velocity", "H", "mass", "massDensity", "specificThermalEnergy", "pressure", "soundSpeed", "omegaGradh"]
        if crk:
            up[-1] = "DvDxQ"
        # what a host-side integrator reads back: the derivatives State::update consumes; "new H" only where the path computes it (the
        # ASPH ideal H is not part of evaluateDerivatives -- ASPHSmoothingScale.cc:110-147 -- so the field would be 48 B/node of zeros)
        self.down_names = ("DxDt", "DrhoDt", "DvDt", "DepsDt", "DvDx", "DHDt") + (() if spec["asph"] else ("Hideal",))
        self.hs, self.pinned, self.up_mask, self.h2d = L.HostState(), [], 0, 0
        dp = self.dp = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
        for k in up:
            t = torch.from_numpy(st[k]).clone().pin_memory()
            self.pinned.append(t); setattr(self.hs, k, dp(t)); self.up_mask |= L.STATE_BITS[k]; self.h2d += t.numel()*8
        self.down_mask = 0
        for k in self.down_names:
            self.down_mask |= L.DERIV_BITS[k]
        self.d2h = sum(N*L.deriv_width(3, k)*8 for k in self.down_names)
        self.down_bufs = {}
        e.upload_state_pinned(self.up_mask, self.hs)
        e.sync()
        # SPHB200_E2E_FUSED=0: evaluate_derivatives + download_derivs as two calls (A/B)
        self.fused_e2e = os.environ.get("SPHB200_E2E_FUSED", "1") != "0"
        self.dsph = None
        if world > 1:
            from spheral_b200 import distributed as D
            # CRKSPH: the volumes and the RK coefficients are state fields too and travel once the package has computed them
            self.dsph = D.DistributedSPH(e, 0, self.lo, self.hi, extra_fields=("volume", "rkCorrections") if crk else ())

    def step(self, evaluate=True):
        e = self.e
        nPG = e.reflect_set_ghost_nodes() if self.planes else 0          # Integrator::setGhostNodes: plane ghosts first ...
        if self.dsph is not None:
            self.dsph.refresh_ghosts(build=True, boundary_ghosts=nPG)    # ... then the slab halo over NVLink, then K1 + K2
        else:
            e.build_pairs()
        if self.crk:                                                     # RKCorrections::preStepInitialize / initialize, each followed
            e.crk_compute_volume()                                       # by applyGhostBoundaries (RKCorrections.cc:298-372)
            if self.dsph is not None:
                self.dsph.mark_ready("volume"); self.dsph.apply_ghosts(("volume",))
            e.crk_compute_corrections()
            if self.dsph is not None:
                self.dsph.mark_ready("rkCorrections"); self.dsph.apply_ghosts(("rkCorrections",))
        if evaluate:
            e.evaluate_derivatives(0.0, 1.0)

    def download(self, fused=False):
        """Derivative fields to the pinned host buffers.  fused: evaluateDerivatives and the download in one C-ABI call
        (sphb200_evaluate_derivatives_to_host: the pair loop runs in chunks of the host index range and the download of one chunk
        overlaps the computation of the next)."""
        e, L, torch = self.e, self.L, self.torch
        n = e.nInternal + e.nGhost      # the C ABI writes every node; ghost entries are zeros
        if self.down_bufs.get("n", 0) < n:
            hd, keep = L.HostDerivs(), []
            for k in self.down_names:
                t = torch.empty((n + n//16)*L.deriv_width(3, k), dtype=torch.float64).pin_memory()
                keep.append(t); setattr(hd, k, self.dp(t))
            self.down_bufs.update(n=n + n//16, hd=hd, keep=keep)
        if fused:
            e._check(e._lib.sphb200_evaluate_derivatives_to_host(e._h, 0.0, 1.0, self.down_mask, self.C.byref(self.down_bufs["hd"])))
        else:
            e._check(e._lib.sphb200_download_derivs(e._h, self.down_mask, self.C.byref(self.down_bufs["hd"])))

    def step_e2e(self):
        e = self.e
        if e.nGhost:
            e.set_nodes(self.N, 0)
        e.upload_state_pinned(self.up_mask, self.hs)
        if self.fused_e2e:
            self.step(evaluate=False)
            self.download(fused=True)
        else:
            self.step()
            self.download()

    def timed(self, fn, k, dist):
        torch = self.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier(dist, self.local)
        t0 = time.perf_counter()
        ev0.record(self.ext)
        for _ in range(k):
            fn()
        ev1.record(self.ext)
        self.e.sync()
        wall = time.perf_counter() - t0
        barrier(dist, self.local)
        return ev0.elapsed_time(ev1)*1e-3, wall

    def checksum(self, dist):
        """Decomposition-independent sums over the internal nodes of the last evaluation (compare an N-GPU run with N = 1)."""
        d = self.e.download_derivs("DvDt", "DepsDt", "DrhoDt")
        N = self.N
        cnt = self.e.download_neighbor_counts()
        vals = [float(np.abs(d["DvDt"][:N]).sum()), float(d["DepsDt"][:N].sum()), float(np.abs(d["DrhoDt"][:N]).sum()), float(cnt.sum()), float(N)]
        vals = [reduce_over_ranks(dist, v, self.local, "sum") for v in vals]
        return {"sum_abs_DvDt": vals[0], "sum_DepsDt": vals[1], "sum_abs_DrhoDt": vals[2], "directed_edges": int(vals[3]), "particles": int(vals[4])}


    def e2e_sums(self, dist):
        """The same sums as checksum(), taken from the HOST buffers the end-to-end leg filled (the derivative fields delivered by
        sphb200_evaluate_derivatives_to_host / sphb200_download_derivs): equal to the device-resident run's to the last bit."""
        try:
            got = {k: t.numpy() for k, t in zip(self.down_names, self.down_bufs["keep"])}
            N = self.N
            vals = [float(np.abs(got["DvDt"].reshape(-1, 3)[:N]).sum()), float(got["DepsDt"][:N].sum()), float(np.abs(got["DrhoDt"][:N]).sum())]
        except Exception:
            vals = [float("nan")]*3
        vals = [reduce_over_ranks(dist, v, self.local, "sum") for v in vals]
        return {"sum_abs_DvDt": vals[0], "sum_DepsDt": vals[1], "sum_abs_DrhoDt": vals[2]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="noh8m")
    ap.add_argument("--nside", type=int, default=0, help="override lattice points per side (debug)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="lattice points per side of the bounded CPU sample (default: per workload)")
    ap.add_argument("--xsph", type=int, default=0, help="1: XSPH on (the round-1 bench setting; the stock scripts run with XSPH=False)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the extra weak-scaling measurement")
    ap.add_argument("--quick", action="store_true", help="A/B diagnostic: hot path only (no RK2 leg, no weak leg, no parity, no CPU baseline)")
    ap.add_argument("--rk2", action="store_true", help="also time device-resident CheapSynchronousRK2 steps (single GPU)")
    args = ap.parse_args()
    warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    spec = workload_spec(args.workload)
    n = args.nside or spec["n"]
    threads = os.cpu_count() or 1
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    config = bench_config(spec, max(world_env, 1) if args.impl == "ours" else max(args.gpus, 1), n)
    cpu_sample = args.cpu_sample or spec["cpu_sample"]

    # ------------------------------------------------------------------ reference arm (CPU) ----------------------------
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        out = claim_stdout()
        val, tmed, N, npairs, sample = cpu_port_run(spec, cpu_sample, max(args.steps, 1), min(warmup, 2), threads, args.xsph)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": tmed*1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                                 "note": "oracle restatement on host cores -- not the Spheral MPI build (unbuildable here)"},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=out, flush=True)
        return 0

    # ------------------------------------------------------------------ our arm (GPU) ---------------------------------------
    out = claim_stdout()
    rank, world, local, dist = dist_setup()
    affinity = pin_to_gpu_numa(local)
    hp = HotPath(spec, n, rank, world, local, dist, args.xsph)
    e = hp.e
    for _ in range(warmup):
        hp.step()
    e.sync()
    fp64_peak = e.measure_fp64_peak()
    launches0 = e.stats()["launches"]

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_s, wall = hp.timed(hp.step, args.steps, dist)
    clocks = sampler.stop() if rank == 0 else None
    launches = e.stats()["launches"] - launches0
    dev_s = reduce_over_ranks(dist, dev_s, local)
    wall_s = reduce_over_ranks(dist, wall, local)

    # per-kernel breakdown (CUDA events inside the library, on the same stream), outside the timed region
    pair_ms, nbr_ms, build_ms, eval_ms = [], [], [], []
    for _ in range(3):
        hp.step()
        s_ = e.stats()
        pair_ms.append(s_["ms_pair_kernel"]); nbr_ms.append(s_["ms_neighbor_kernels"])
        build_ms.append(s_["ms_build_pairs"]); eval_ms.append(s_["ms_evaluate"])
    st_ = e.stats()
    edges_local, nGhost_local = st_["directed_edges"], e.nGhost
    edges = reduce_over_ranks(dist, float(edges_local), local, "sum")
    ghosts = reduce_over_ranks(dist, float(nGhost_local), local, "sum")
    Ntot = int(reduce_over_ranks(dist, float(hp.N), local, "sum"))
    halo_info = hp.dsph.info() if hp.dsph is not None else None
    checksum = hp.checksum(dist)

    # device-resident CheapSynchronousRK2 steps, single GPU, on request
    rk2 = None
    if args.rk2 and hp.dsph is None and not args.quick:
        from spheral_b200 import integrator as I, engine
        rk = I.CheapSynchronousRK2(e, engine.make_step_options(), reflectingPlanes=hp.planes or None)
        rk.initializeDerivatives()
        for _ in range(2):
            rk.step()
        e.sync()
        nrk = max(3, args.steps//2)
        rk_s, _ = hp.timed(rk.step, nrk, dist)
        rk2 = {"ms_per_step": rk_s/nrk*1e3, "value": hp.N*nrk/rk_s, "unit": UNIT, "steps": nrk, "last_dt": rk.lastDt, "dt_reason": rk.lastDtReason,
               "what": "CheapSynchronousRK2 step with the state resident in HBM (ghosts, build_pairs, sum density, dt, State::update x2, grad-h x2, "
                       "evaluateDerivatives, compatible energy); one 16-byte read-back (dt) per step"}
        # the same with the end-of-step grad-h correction computed on demand only (integrator.lazyOmega: no evaluation reads it)
        rk.lazyOmega = True
        rk_l, _ = hp.timed(rk.step, nrk, dist)
        rk.ensureOmega()
        rk2["ms_per_step_lazy_omega"] = rk_l/nrk*1e3
        e.set_nodes(hp.N, 0)
        e.upload_state_pinned(hp.up_mask, hp.hs)
        e.sync()

    # e2e leg: host buffers in, host buffers out, every step
    for _ in range(3):
        hp.step_e2e()
    e.sync()
    _, e2e_wall = hp.timed(hp.step_e2e, args.steps, dist)
    e2e_s = reduce_over_ranks(dist, e2e_wall, local)
    h2d = reduce_over_ranks(dist, float(hp.h2d), local, "sum")
    d2h = reduce_over_ranks(dist, float(hp.d2h), local, "sum")
    e2e_sums = hp.e2e_sums(dist)
    e2e_same = all(e2e_sums[k] == checksum[k] for k in e2e_sums)

    # weak-scaling extra (N > 1): one 100^3 cube of the same workload per GPU, the round-1 measurement
    weak = None
    if world > 1 and not args.no_weak and not args.quick:
        del hp.pinned, hp.down_bufs
        e.close()
        wn = 100
        hw = HotPath(spec, wn, rank, world, local, dist, args.xsph, weak=True)
        for _ in range(3):
            hw.step()
        hw.e.sync()
        wk = max(5, args.steps//2)
        w_s, _ = hw.timed(hw.step, wk, dist)
        w_s = reduce_over_ranks(dist, w_s, local)
        weak = {"scaling": "weak", "particles_per_gpu": wn**3, "value": float(wn**3)*world*wk/w_s, "unit": UNIT, "ms_per_step": w_s/wk*1e3, "steps": wk,
                "what": "one %d^3 cube of the same workload per GPU, side by side along x, slab halo over NCCL" % wn}
        hw.e.close()

    if rank == 0:
        nbrs = edges/float(Ntot)
        total_updates = float(Ntot)*args.steps
        value = total_updates/dev_s
        # roofline of the dominant kernel (pair loop): algorithmic work per launch of rank 0
        N0 = hp.N
        nb0 = edges_local/float(N0)
        t_pair = float(np.mean(pair_ms))*1e-3
        crk = hp.crk
        bytes_alg = N0*(672.0 + 12.0*nb0)            # API-faithful variant: state in, derivatives out, 24 B per pair
        flops_alg = N0*250.0*nb0                     # ~500 flop per pair, each pair counted once
        if crk:                                      # + volume and 16 RK coefficients per node in; ~560 flop per pair
            bytes_alg = N0*(672.0 + 136.0 + 12.0*nb0)
            flops_alg = N0*280.0*nb0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tj.get(args.workload if (not args.nside and world == 1 and not args.xsph) else "", {})
            key = "k_crk_derivs" if crk else "k_sph_derivs"
            if key in ent:
                traffic = float(ent[key]["dram_bytes_read"]) + float(ent[key]["dram_bytes_write"])
                traffic_src = ent[key].get("source")
        except Exception:
            pass
        kname = "k_crk_derivs (CRKSPH pair loop + finalize + smoothing scale)" if crk else "k_sph_derivs (SPH pair loop + finalize + smoothing scale)"
        roofline = {"bound": "hbm", "kernel": kname, "achieved": bytes_alg/t_pair/1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": bytes_alg/t_pair/1e9/hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": hbm_src,
                    "algorithmic_bytes_per_particle": bytes_alg/N0, "launch_ms": t_pair*1e3,
                    "note": "HBM is not the binding roof at ~100 neighbours (arithmetic intensity ~13 flop/B vs machine balance ~5): the FP64 pipe is; see roofline_fp64"}
        t_call = dev_s/args.steps
        roofline_fp64 = {"bound": "fp64", "kernel": kname, "achieved": flops_alg/t_pair/1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": flops_alg/t_pair/1e12/fp64_peak, "algorithmic_flops_per_particle": flops_alg/N0,
                         "whole_call_frac": (250.0*nbrs*Ntot/world)/t_call/1e12/fp64_peak,
                         "peak_source": "DFMA microbenchmark run in this process (sphb200_measure_fp64_peak)",
                         "note": "whole_call_frac = algorithmic pair flops of one rank / (ghosts + build_pairs + evaluateDerivatives) time"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": dev_s/args.steps*1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "details": {"neighbours_per_particle": nbrs, "ghost_nodes_all_ranks": int(ghosts), "halo": halo_info,
                            "grid_stencil_radius": int(st_["stencil_radius"]), "xsph": int(args.xsph), "cpu_affinity": affinity},
                "clocks": clocks,
                "e2e": {"value": float(Ntot)*args.steps/e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "timing": "host wall clock around synchronised C-ABI calls (pinned host buffers), max over ranks",
                        "host_results_equal_device_resident": bool(e2e_same),
                        "calls": ("upload_state, reflect_set_ghost_nodes, [halo], build_pairs, evaluate_derivatives_to_host (pair loop in chunks of the host index "
                                  "range, each chunk's download overlapped with the next chunk's computation)") if hp.fused_e2e else
                                 "upload_state, reflect_set_ghost_nodes, [halo], build_pairs, evaluate_derivatives, download_derivs"},
                "gpu_launches": int(launches),
                "breakdown_ms": {"build_pairs": float(np.mean(build_ms)), "neighbor_kernels": float(np.mean(nbr_ms)),
                                 "evaluate": float(np.mean(eval_ms)), "pair_kernel": float(np.mean(pair_ms)),
                                 "wall_per_step": wall_s/args.steps*1e3},
                "roofline": roofline, "roofline_fp64": roofline_fp64, "checksum": checksum}
        if rk2 is not None:
            line["rk2_step_resident"] = rk2
        if weak is not None:
            line["weak"] = weak
        if world == 1 and not args.quick:
            if not args.no_parity:
                try:
                    line["parity"] = parity_check(spec, spec["parity_sample"], args.xsph, local)
                except Exception as ex:          # a failed check must be visible, never silently absent
                    line["parity"] = {"ok": False, "error": repr(ex)}
            if not args.no_cpu_baseline:
                val, tmed, Ns, npairs, sample = cpu_port_run(spec, cpu_sample, 3, 1, threads, args.xsph)
                line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                                        "note": "oracle restatement on host cores -- not the Spheral MPI build (unbuildable here)"}
        print(json.dumps(line), file=out, flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
