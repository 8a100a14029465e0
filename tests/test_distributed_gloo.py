"""World-size-2 and -3 CPU (gloo) tests of the multi-GPU host logic: slab decomposition, ghost selection, count exchange and the
two-phase point-to-point halo exchange of spheral_b200/distributed.py.  Each rank assembles its slab (internal + received
ghosts), runs the ORACLE on it, and the internal-node derivatives must equal the oracle's on the undecomposed problem
(the pair loop needs no communication once the ghosts are in place)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _global_problem(ndim, n, nPerh, aniso):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import common
    st, nInt, nGhost = common.make_problem(ndim, n, nPerh=nPerh, seed=23)
    if aniso:       # mildly anisotropic, rotated H so that gather and scatter neighbours differ
        from spheral_b200 import nodegen as ng
        rng = np.random.default_rng(5)
        F = ng.sym_to_full(ndim, st["H"])
        for k in range(F.shape[0]):
            Rm = ng.random_rotation(ndim, rng)
            s = np.diag(rng.uniform(0.7, 1.3, size=ndim))
            F[k] = Rm @ (F[k] @ s) @ Rm.T
        st["H"] = np.ascontiguousarray(ng.full_to_sym(ndim, 0.5*(F + np.swapaxes(F, 1, 2))))
    return st, nInt


def _worker(rank, world, port, ndim, n, nPerh, aniso, planes=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import common
        from oracle import oracle as orc
        from spheral_b200 import distributed as D, kernel as K, _lib as L
        st, N = _global_problem(ndim, n, nPerh, aniso)
        WT = K.TableKernel(K.BSplineKernel(ndim), 200)
        OT = common.oracle_table(orc, WT)
        kext = WT.kernelExtent
        oo = orc.default_options(ndim, nPerh=nPerh, Cl=2.0, Cq=2.0)
        axis = 0
        edges = D.slab_edges(0.0, 1.0, world)
        lo, hi = edges[rank], edges[rank + 1]
        x = st["position"][:, axis]
        if rank == world - 1:
            mine = np.nonzero((x >= lo))[0]
        elif rank == 0:
            mine = np.nonzero((x < hi))[0]
        else:
            mine = np.nonzero((x >= lo) & (x < hi))[0]
        local = {k: np.ascontiguousarray(v[mine]) for k, v in st.items()}
        nInt = len(mine)
        nBG = 0
        if planes:
            # reflecting planes through the origin on every axis, generated per slab BEFORE the halo (boundary order of the
            # reference: the problem's boundaries first, the DistributedBoundary last); the plane ghosts then take part in the
            # send-node selection like internal nodes (DistributedSPH.refresh_ghosts(boundary_ghosts=...))
            from spheral_b200 import nodegen as ng
            PN = dict(position="pos", velocity="vel", H="H", mass="mass", massDensity="rho", specificThermalEnergy="eps", pressure="P",
                      soundSpeed="cs", omegaGradh="omega")
            plist = [(np.zeros(ndim), np.eye(ndim)[a]) for a in range(ndim)]
            out, _, _ = ng.reflect_ghosts(ndim, {o: local[k] for k, o in PN.items()}, plist, kext, per_plane=True)
            local = {k: np.ascontiguousarray(out[o]) for k, o in PN.items()}
            nBG = local["position"].shape[0] - nInt
            assert nBG > 0
            gout, _, _ = ng.reflect_ghosts(ndim, {o: st[k] for k, o in PN.items()}, plist, kext, per_plane=True)
            stg = {k: np.ascontiguousarray(gout[o]) for k, o in PN.items()}
            NG = stg["position"].shape[0] - N
        else:
            stg, NG = st, 0

        halo = D.SlabHalo()
        assert (halo.lower, halo.upper) == (rank - 1 if rank > 0 else None, rank + 1 if rank < world - 1 else None)
        ext = D.kernel_extent_axis(local["H"], ndim, kext, axis).max()
        nOwn = nInt + nBG
        width = float(halo.allreduce_max(torch.tensor([ext], dtype=torch.float64)).item())*(1.0 + 1e-9)
        idxLow, idxHigh = D.select_halo_numpy(local["position"], nOwn, axis, lo, hi, width)
        if halo.lower is None:
            idxLow = idxLow[:0]
        if halo.upper is None:
            idxHigh = idxHigh[:0]
        nFL, nFU = halo.exchange_counts(len(idxLow), len(idxHigh), torch.device("cpu"))
        ghosts = {}
        for names in (D.PHASE_A, D.PHASE_B):
            wid = sum(L.state_width(ndim, k) for k in names)
            sL = torch.from_numpy(D.pack_fields_numpy(local, names, idxLow, ndim))
            sH = torch.from_numpy(D.pack_fields_numpy(local, names, idxHigh, ndim))
            rL, rH = torch.empty(nFL*wid, dtype=torch.float64), torch.empty(nFU*wid, dtype=torch.float64)
            halo.finish(halo.start(sL, sH, rL, rH))
            gl, gh = D.unpack_fields_numpy(rL.numpy(), names, nFL, ndim), D.unpack_fields_numpy(rH.numpy(), names, nFU, ndim)
            for k in names:
                ghosts[k] = np.concatenate([gl[k], gh[k]])
        full = {k: np.ascontiguousarray(np.concatenate([local[k], ghosts[k]])) for k in local}
        nGhost = nBG + nFL + nFU
        assert nFL + nFU > 0

        # oracle on the slab (internal + ghosts) versus oracle on the whole problem
        s = common.to_oracle_state(full)
        pi, pj, cnt = orc.pairs(ndim, nInt, nGhost, s["pos"], s["H"], kext)
        got = orc.evaluate_derivatives(oo, OT, s, nInt, nGhost, pi, pj, cnt)
        sg = common.to_oracle_state(stg)
        gpi, gpj, gcnt = orc.pairs(ndim, N, NG, sg["pos"], sg["H"], kext)
        ref = orc.evaluate_derivatives(oo, OT, sg, N, NG, gpi, gpj, gcnt)
        assert np.array_equal(cnt, gcnt[mine]), "a slab node lost or gained neighbours: the halo is not a superset"
        floors = common.physical_floors(st, N, ndim)
        for k, f in floors.items():
            a, b = np.asarray(got[k])[:nInt], np.asarray(ref[k])[mine]
            err = np.abs(a - b).max()/max(np.abs(b).max(), f)
            assert err <= 1.0e-12, (k, err)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ndim,n,nPerh,aniso,planes", [(3, 10, 1.51, False, False), (2, 24, 2.01, True, False),
                                                       (3, 10, 1.51, False, True), (2, 24, 2.01, True, True)])
def test_two_slab_halo_exchange_reproduces_global_derivatives(ndim, n, nPerh, aniso, planes):
    """planes=True: reflecting planes on top of the slabs -- the ghost tail of a slab is [its plane ghosts | halo] and the halo
    carries the neighbour's plane ghosts near the shared face; the derivatives of the internal nodes must still equal those of the
    undecomposed problem with the same planes."""
    from oracle import oracle as orc
    orc.build()
    from spheral_b200 import build as b
    b.build()
    mp.spawn(_worker, args=(2, _free_port(), ndim, n, nPerh, aniso, planes), nprocs=2, join=True)


@pytest.mark.parametrize("ndim,n,nPerh,aniso,planes", [(3, 12, 1.51, False, False), (2, 30, 2.01, True, True)])
def test_three_slab_halo_exchange_with_a_two_peer_rank(ndim, n, nPerh, aniso, planes):
    """World size 3: the middle slab exchanges with two peers (both send lists used, ghosts from below and from above), the end slabs
    with one; same criterion as the two-slab test."""
    from oracle import oracle as orc
    orc.build()
    from spheral_b200 import build as b
    b.build()
    mp.spawn(_worker, args=(3, _free_port(), ndim, n, nPerh, aniso, planes), nprocs=3, join=True)


def test_send_list_redo_decision_is_the_same_on_every_rank():
    """The capacity check of the send-node selection (DistributedSPH.refresh_ghosts) is decided from the gathered table, so every
    rank redoes the selection or none does.  Lists that are never sent (the first slab's low list -- e.g. full of the ghosts of an
    x = 0 reflecting plane -- and the last slab's high list) must not trigger a redo."""
    from spheral_b200 import distributed as D
    # rank 0 overflows only its unused low list; everything that is sent fits
    allc = np.array([[9000, 2500, 5000], [2600, 2700, 5000], [2500, 8000, 5000]])
    need, fits = D.send_list_needs(allc, 3)
    assert need == [2500, 2700, 2500] and fits
    # one rank's used list does not fit: every rank sees the same verdict and the same size to grow to
    allc[1, 1] = 5001
    need, fits = D.send_list_needs(allc, 3)
    assert not fits and max(need) == 5001
    # single slab: nothing is sent at all
    assert D.send_list_needs(np.array([[10**6, 10**6, 1024]]), 1) == ([0], True)
