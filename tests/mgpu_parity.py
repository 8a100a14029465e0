"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_parity.py

Each rank owns one slab of a jittered lattice, exchanges ghosts with its neighbours through DistributedSPH (device-side
selection / pack / unpack + NCCL send/recv), builds its pairs and evaluates the derivatives on its GPU.  Rank by rank the
internal-node results are compared with the CPU oracle run on the WHOLE problem: neighbour counts bit-exact, derivative
fields within 1e-10 (field-wise max-norm metric of SURVEY.md 8c)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("MGPU_HANG_DUMP", "150")), exit=True)    # a deadlock prints its stack and ends
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import common
    from oracle import oracle as orc
    from spheral_b200 import distributed as D, engine, kernel as K, nodegen as ng
    n, nPerh, ndim = int(os.environ.get("MGPU_N", "20")), 1.51, 3
    asph = os.environ.get("MGPU_ASPH", "0") == "1"
    # global problem: world cubes side by side along x, same on every rank
    parts = []
    for k in range(world):
        st, nInt, _ = common.make_problem(ndim, n, nPerh=nPerh, seed=101 + k)
        st["position"][:, 0] += k
        parts.append(st)
    G = {k: np.ascontiguousarray(np.concatenate([p[k] for p in parts])) for k in parts[0]}
    if asph:
        rng = np.random.default_rng(9)
        F = ng.sym_to_full(ndim, G["H"])
        for q in range(F.shape[0]):
            Rm = ng.random_rotation(ndim, rng)
            F[q] = Rm @ (F[q] @ np.diag(rng.uniform(0.75, 1.25, size=ndim))) @ Rm.T
        G["H"] = np.ascontiguousarray(ng.full_to_sym(ndim, 0.5*(F + np.swapaxes(F, 1, 2))))
    # MGPU_QKIND=1: LimitedMonaghanGingold with Balsara switch and Cl / Cq multipliers -- the pair loop then reads the velocity
    # gradient and the multipliers of GHOST neighbours, so all three must travel with the halo
    qkind = int(os.environ.get("MGPU_QKIND", "0"))
    if qkind:
        G = common.add_q_fields(G, ndim, seed=31)
    N = n**3
    mine = slice(rank*N, (rank + 1)*N)
    WT = K.TableKernel(K.BSplineKernel(ndim), 1000)
    crk = os.environ.get("MGPU_CRK", "0") == "1"          # CRKSPH: volumes and RK corrections travel with the halo
    kw = dict(nPerh=nPerh, Cl=2.0, Cq=2.0, hEvolution=1 if asph else 0)
    if qkind:
        kw.update(Qkind=1, balsara=1)
    if crk:
        kw = dict(nPerh=nPerh, Cl=1.0, Cq=0.25, Qkind=0, correctVelocityGradient=0)
        G["velocity"] = 0.3*G["velocity"]
    oo, po = common.opts_pair(orc, engine, ndim, **kw)
    if crk:
        from spheral_b200 import _lib as L
        po.hydro = L.HYDRO_CRKSPH
    e = engine.Engine(ndim, device=local, options=po)
    e.set_kernel_table(WT)
    e.set_nodes(N, 0)
    e.upload_state(**{k: v[mine] for k, v in G.items()})
    d = D.DistributedSPH(e, 0, float(rank), float(rank + 1), extra_fields=("volume", "rkCorrections") if crk else ())
    if os.environ.get("MGPU_RK2", "0") == "1" or crk:
        return rk2_mode(rank, world, e, d, G, mine, N, oo, WT, orc, common, engine, crk)
    for _ in range(2):                       # twice: the second pass reuses every buffer
        npairs = d.step_connectivity_and_derivatives(0.0, 1.0)
    got = e.download_derivs()
    cnt = e.download_neighbor_counts()
    e.sync()

    sg = common.to_oracle_state(G)
    gpi, gpj, gcnt = orc.pairs(ndim, world*N, 0, sg["pos"], sg["H"], WT.kernelExtent)
    ref = orc.evaluate_derivatives(oo, common.oracle_table(orc, WT), sg, world*N, 0, gpi, gpj, gcnt, nthreads=8)
    ok = bool(np.array_equal(cnt, gcnt[mine]))
    floors = common.physical_floors(G, world*N, ndim)
    worst = {}
    for k, f in floors.items():
        a, b = np.asarray(got[k])[:N], np.asarray(ref[k])[mine]
        worst[k] = float(np.abs(a - b).max()/max(np.abs(b).max(), f))
    w = max(worst.values())
    res = dict(rank=rank, world=world, nodes=N, ghosts=e.nGhost, pairs=int(npairs), counts_equal=ok, worst_field_error=w,
               halo=d.info(), asph=asph, qkind=qkind)
    print(json.dumps(res), flush=True)
    flag = torch.tensor([1.0 if (ok and w <= 1.0e-10) else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


def rk2_mode(rank, world, e, d, G, mine, N, oo, WT, orc, common, engine, crk=False):
    """Device-resident CheapSynchronousRK2 steps on the decomposed problem (ghost refresh between the stages, DvDt / DepsDt of the
    ghosts after the derivatives, dt all-reduced) against the oracle-driven integrator on the WHOLE problem."""
    from spheral_b200 import integrator
    nsteps = int(os.environ.get("MGPU_STEPS", "3"))
    so = orc.default_step_options()
    # MGPU_PLANES=1: reflecting planes x = 0, y = 0, z = 0 on top of the slabs (the ghost tail of a rank is then
    # [its plane ghosts | halo], and the halo carries the neighbours' plane ghosts near the shared face)
    planes = [(np.zeros(3), np.eye(3)[a]) for a in range(3)] if os.environ.get("MGPU_PLANES", "0") == "1" else None
    rk = integrator.CheapSynchronousRK2(e, engine.make_step_options(), densityUpdate=1, distributed=d, reflectingPlanes=planes)
    rk.initializeDerivatives()
    for _ in range(nsteps):
        assert rk.step()
    fields = ("position", "velocity", "H", "massDensity", "specificThermalEnergy") + (("volume",) if crk else ("omegaGradh",))
    got = e.download_state(*fields)
    e.sync()
    G2 = dict(G); G2["velocity"] = G["velocity"]
    ref = common.OracleRK2(orc, oo, so, common.oracle_table(orc, WT), G2, densityUpdate=1, planes=planes, crk=crk)
    ref.initializeDerivatives()
    for _ in range(nsteps):
        ref.step()
    names = dict(position="pos", velocity="vel", H="H", massDensity="rho", specificThermalEnergy="eps")
    names.update(dict(volume="vol") if crk else dict(omegaGradh="omega"))
    refInt = {o: ref.s[o][:world*N] for o in names.values()}           # the oracle state carries its plane ghosts behind the internal nodes
    worst = {k: float(np.abs(got[k][:N] - refInt[o][mine]).max()/max(np.abs(refInt[o][mine]).max(), 1e-300)) for k, o in names.items()}
    w = max(worst.values())
    dt_err = abs(rk.lastDt - ref.lastDt)/ref.lastDt
    # global energy on the device results
    m = G["mass"][mine]
    Eloc = float(np.sum(m*(0.5*np.sum(got["velocity"][:N]**2, axis=1) + got["specificThermalEnergy"][:N])))
    E0loc = float(np.sum(m*(0.5*np.sum(G["velocity"][mine]**2, axis=1) + G["specificThermalEnergy"][mine])))
    t = torch.tensor([Eloc, E0loc], dtype=torch.float64, device="cuda")
    dist.all_reduce(t)
    dE = float((t[0] - t[1])/t[1])
    print(json.dumps(dict(rank=rank, world=world, mode="rk2", crk=crk, planes=bool(planes), steps=nsteps, ghosts=e.nGhost, halo=d.last,
                          oracle_plane_ghosts=ref.nGhost, worst_state_error=w, worst=worst,
                          dt_rel_err=dt_err, dE_over_E=dE, time=rk.currentTime)), flush=True)
    ok = (w <= 1.0e-9) and (dt_err <= 1.0e-10) and (abs(dE) <= 1.0e-12)
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
