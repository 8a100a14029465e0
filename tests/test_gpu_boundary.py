"""GPU parity of the device-side reflecting boundaries and the Noh problem run with the state
resident on the GPU (BASELINE configs[0]: Noh-cylindrical-2d, two reflecting planes, v = -rhat, eps = 0)."""
import numpy as np
import pytest

import common
from spheral_b200 import kernel as K, nodegen as ng

pytestmark = pytest.mark.gpu
PN = dict(position="pos", velocity="vel", H="H", mass="mass", massDensity="rho", specificThermalEnergy="eps", pressure="P",
          soundSpeed="cs", omegaGradh="omega")


@pytest.fixture(scope="module")
def mods(sphlib):
    from spheral_b200 import engine, integrator
    return engine, integrator


def planes_of(ndim):
    return [(np.zeros(ndim), np.eye(ndim)[a]) for a in range(ndim)]


@pytest.mark.parametrize("ndim,n,nPerh,kind", [(2, 24, 2.01, "lattice"), (3, 9, 1.51, "lattice"), (2, 16, 2.01, "aniso"), (3, 7, 1.51, "aniso")])
def test_ghost_generation_matches_host_restatement(oracle, mods, ndim, n, nPerh, kind):
    engine, _ = mods
    st, nInt, _ = common.make_problem(ndim, n, nPerh=nPerh, kind=kind)
    WT = K.TableKernel(K.BSplineKernel(ndim), 1000)
    f = {o: st[k] for k, o in PN.items()}
    ref, lists, n0 = ng.reflect_ghosts(ndim, f, planes_of(ndim), WT.kernelExtent, per_plane=True)
    e = engine.Engine(ndim, nPerh=nPerh)
    e.set_kernel_table(WT)
    e.set_nodes(nInt, 0)
    e.upload_state(**st)
    e.reflect_configure(planes_of(ndim))
    ng_dev = e.reflect_set_ghost_nodes()
    assert ng_dev == ref["pos"].shape[0] - nInt and ng_dev > 0
    got = e.download_state(*PN.keys())
    for k, o in PN.items():
        a, b = got[k], ref[o]
        assert a.shape == b.shape
        assert np.array_equal(a[:nInt], b[:nInt]), k                                  # internal state untouched
        assert np.abs(a - b).max() <= 1e-14*max(np.abs(b).max(), 1e-300), k          # ghosts: same controls, same order, same values
    # derivatives on the device-generated ghost set == oracle on the host-generated one
    OT = common.oracle_table(oracle, WT)
    oo, po = common.opts_pair(oracle, engine, ndim, nPerh=nPerh)
    s = dict(ref)
    pi, pj, cnt = oracle.pairs(ndim, nInt, ng_dev, s["pos"], s["H"], OT.kext)
    d = oracle.evaluate_derivatives(oo, OT, s, nInt, ng_dev, pi, pj, cnt)
    assert e.build_pairs() == len(pi)
    e.evaluate_derivatives()
    g = e.download_derivs("DvDt", "DrhoDt", "DepsDt", "DvDx")
    st_g = {k: got[k] for k in PN}
    floors = common.physical_floors(st_g, nInt, ndim)
    for k in g:
        assert common.field_err(g[k], d[k], nInt, floors[k]) <= 1e-10, k
    # refresh after a state change on the device: bump the velocity of the internal nodes, re-apply the boundary
    v = got["velocity"].copy(); v[:nInt] *= 1.5
    e.upload_state(velocity=v)
    e.reflect_apply_ghosts()
    ref2 = {k: np.array(x, copy=True) for k, x in ref.items()}
    ref2["vel"][:nInt] *= 1.5
    ng.reflect_apply(ndim, ref2, planes_of(ndim), lists, nInt)
    got2 = e.download_state("velocity", "H")
    assert np.abs(got2["velocity"] - ref2["vel"]).max() <= 1e-14*np.abs(ref2["vel"]).max()
    assert np.abs(got2["H"] - ref2["H"]).max() <= 1e-14*np.abs(ref2["H"]).max()


@pytest.mark.parametrize("ndim,n,nPerh", [(2, 20, 2.01), (3, 8, 1.51)])
def test_rk_coefficients_are_transformed_not_copied_into_reflecting_ghosts(oracle, mods, ndim, n, nPerh):
    """ReflectingBoundary::applyGhostBoundary(Field<RKCoefficients>) (Boundary/ReflectingBoundary.cc:403-432) applies
    RKUtilities::getTransformationMatrix(R) (RK/RKUtilities.cc:637-715) to the copied coefficients: A' = A, B' = R.B,
    (grad A)' = R.grad A, (grad B)' = R.(grad B).R.  Device ghosts of a CRKSPH context against the numpy restatement, and the
    restatement against the truth: the transformed coefficients of a mirror image equal the coefficients computed for a real node at
    the mirrored position (found with the reference's Noh-planar-1d CRKSPH golden: plain copies were off by O(1))."""
    engine, _ = mods
    from spheral_b200 import _lib as L
    st, nInt, _ = common.make_problem(ndim, n, nPerh=nPerh, kind="lattice", seed=31)
    WT = K.TableKernel(K.BSplineKernel(ndim), 1000)
    OT = common.oracle_table(oracle, WT)
    s0 = common.to_oracle_state(st)
    pi, pj, cnt = oracle.pairs(ndim, nInt, 0, s0["pos"], s0["H"], OT.kext)
    vol = oracle.crk_sum_volume(ndim, OT, nInt, 0, s0["pos"], s0["H"], pi, pj)
    corr = oracle.crk_corrections(ndim, OT, nInt, 0, s0["pos"], s0["H"], vol, pi, pj)       # one-sided at the faces: B and grad A non-zero
    f = {o: st[k] for k, o in PN.items()}
    f["vol"], f["corr"] = vol, corr
    ref, lists, n0 = ng.reflect_ghosts(ndim, f, planes_of(ndim), WT.kernelExtent, per_plane=True)
    po = engine.make_options(ndim, nPerh=nPerh, hydro=L.HYDRO_CRKSPH)
    e = engine.Engine(ndim, options=po)
    e.set_kernel_table(WT)
    e.set_nodes(nInt, 0)
    e.upload_state(volume=vol, rkCorrections=corr, **st)
    e.reflect_configure(planes_of(ndim))
    nG = e.reflect_set_ghost_nodes()
    assert nG == ref["pos"].shape[0] - nInt and nG > 0
    got = e.download_state("rkCorrections", "volume")
    assert np.abs(got["volume"] - ref["vol"]).max() <= 1e-14*np.abs(ref["vol"]).max()
    scale = np.abs(ref["corr"]).max()
    assert np.abs(got["rkCorrections"] - ref["corr"]).max() <= 1e-13*scale
    # a copy would be wrong: the first plane flips the sign of B_x and d A / dx of its ghosts
    ctl0 = lists[0]
    g0 = ref["corr"][nInt:nInt + len(ctl0)]
    assert np.abs(g0[:, 1] + corr[ctl0, 1]).max() <= 1e-13*scale and np.abs(corr[ctl0, 1]).max() > 1e-3*scale


def test_enforce_maps_violators_back(mods):
    engine, _ = mods
    ndim = 2
    st, nInt, _ = common.make_problem(ndim, 10, nPerh=2.01)
    e = engine.Engine(ndim, nPerh=2.01)
    e.set_kernel_table(K.TableKernel(K.BSplineKernel(ndim), 1000))
    e.set_nodes(nInt, 0)
    pos = st["position"].copy(); vel = st["velocity"].copy()
    pos[3, 0] = -0.02; pos[7, 1] = -0.01; pos[9] = (-0.03, -0.04)
    e.upload_state(**dict(st, position=pos, velocity=vel))
    e.reflect_configure(planes_of(ndim))
    assert e.reflect_enforce(count=True) == 4          # node 9 violates both planes
    got = e.download_state("position", "velocity")
    rp, rv = pos.copy(), vel.copy()
    for a in range(ndim):
        bad = rp[:, a] < 0
        rp[bad, a] *= -1.0; rv[bad, a] *= -1.0
    assert np.abs(got["position"] - rp).max() <= 1e-16 and np.abs(got["velocity"] - rv).max() <= 1e-15
    assert e.reflect_enforce(count=True) == 0


@pytest.mark.parametrize("ndim", [2, 3])
def test_enforce_reflects_the_H_tensor_of_violation_nodes(mods, ndim):
    """PlanarBoundary::updateViolationNodes (PlanarBoundary.cc:179-195) enforces the boundary on position, velocity AND H;
    ReflectingBoundary::enforceBoundary(Field<SymTensor>) (ReflectingBoundary.cc:493-501) maps H -> (R H R).Symmetric().  An oblique
    plane and anisotropic (ASPH) tensors make the difference visible."""
    engine, _ = mods
    st, nInt, _ = common.make_problem(ndim, 7 if ndim == 3 else 12, nPerh=1.51, kind="aniso", seed=5)
    nrm = np.array([1.0, 0.7, -0.4][:ndim]); nrm /= np.linalg.norm(nrm)
    point = np.full(ndim, 0.5)
    e = engine.Engine(ndim, nPerh=1.51, hEvolution=1)
    e.set_kernel_table(K.TableKernel(K.BSplineKernel(ndim), 1000))
    e.set_nodes(nInt, 0)
    e.upload_state(**st)
    e.reflect_configure([(point, nrm)])
    sd = (st["position"] - point) @ nrm
    bad = np.nonzero(sd < 0.0)[0]
    assert 0 < len(bad) < nInt
    assert e.reflect_enforce(count=True) == len(bad)
    got = e.download_state("position", "velocity", "H")
    ref = {k: st[k].copy() for k in ("position", "velocity", "H")}
    ref["position"][bad] -= 2.0*np.outer(sd[bad], nrm)
    ref["velocity"][bad] -= 2.0*np.outer(st["velocity"][bad] @ nrm, nrm)
    ref["H"][bad] = ng.reflect_map(ndim, "H", st["H"][bad], sd[bad], nrm)
    for k in ref:
        assert np.abs(got[k] - ref[k]).max() <= 1e-13*np.abs(ref[k]).max(), k
    # the reflected tensors really differ from the originals (the old behaviour), and untouched nodes keep theirs bit for bit
    assert np.abs(got["H"][bad] - st["H"][bad]).max() > 1e-3*np.abs(st["H"]).max()
    good = np.setdiff1d(np.arange(nInt), bad)
    assert np.array_equal(got["H"][good], st["H"][good])


def noh_2d(nRadial, nPerh):
    pos, mass, H = ng.constant_dtheta_2d(nRadial, nPerh=nPerh)
    N = len(pos)
    r = np.linalg.norm(pos, axis=1)
    st = dict(position=pos, velocity=-pos/r[:, None], H=H, mass=mass, massDensity=np.ones(N), specificThermalEnergy=np.zeros(N),
              pressure=np.zeros(N), soundSpeed=np.zeros(N), omegaGradh=np.ones(N))
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in st.items()}, N


def test_noh_cylindrical_2d_device_resident(oracle, mods):
    """Noh-cylindrical-2d.py set-up (WendlandC4, nPerh 4.01, Cl = Cq = 1, hmin 1e-4, hmax 0.1, hminratio 0.1, RigorousSumDensity,
    IdealH, compatible energy, dtMin 1e-8, dtMax 0.1, dtGrowth 2, goalTime 0.6), scaled to nRadial = 50, state resident on the GPU:
    three steps against the oracle-driven integrator, then on to t = 0.6 against (i) the golden radial profile of the oracle run
    (tests/golden/noh2d_nr50_t06.json), (ii) the analytic solution (NohAnalyticSolution.py: rho = 1 + t/r ahead of the shock, a
    16-fold compression behind it -- 13.5 at this resolution) and (iii) the energy check of Noh-cylindrical-2d.py:803-808."""
    import json
    import os
    engine, integrator = mods
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "noh2d_nr50_t06.json")))
    ndim, nRadial, nPerh, tend = 2, gold["nRadial"], gold["nPerh"], gold["tend"]
    st, N = noh_2d(nRadial, nPerh)
    assert N == gold["nodes"]
    WT = K.TableKernel(K.WendlandC4Kernel(ndim), 1000)
    OT = common.oracle_table(oracle, WT)
    okw = dict(nPerh=nPerh, Cl=1.0, Cq=1.0, hmin=1.0e-4, hmax=0.1)
    oo, po = common.opts_pair(oracle, engine, ndim, **okw)
    so = oracle.default_step_options(hminratio=0.1)
    planes = planes_of(ndim)
    ref = common.OracleRK2(oracle, oo, so, OT, st, densityUpdate=1, planes=planes, dtMin=1.0e-8, dtMax=0.1)
    e = engine.Engine(ndim, options=po)
    e.set_kernel_table(WT)
    e.set_nodes(N, 0)
    e.upload_state(**st)
    rk = integrator.CheapSynchronousRK2(e, engine.make_step_options(hminratio=0.1), densityUpdate=1, reflectingPlanes=planes,
                                        dtMin=1.0e-8, dtMax=0.1)
    m = st["mass"]
    E0 = float(np.sum(m*0.5*np.sum(st["velocity"]**2, axis=1)))
    ref.initializeDerivatives()
    rk.initializeDerivatives()
    assert e.nGhost == ref.nGhost > 0
    for _ in range(3):
        dt_ref = ref.step(tend)
        assert rk.step(tend)
        assert abs(rk.lastDt - dt_ref) <= 1e-10*dt_ref, (rk.lastDt, dt_ref)
        assert e.nGhost == ref.nGhost
    got = e.download_state("position", "velocity", "H", "massDensity", "specificThermalEnergy")
    worst = {k: float(np.abs(got[k][:N] - ref.s[PN[k]][:N]).max()/max(np.abs(ref.s[PN[k]][:N]).max(), 1e-300)) for k in got}
    assert all(v <= 1.0e-9 for v in worst.values()), worst
    # on to t = 0.6
    rk.advance(tend, maxSteps=4000)
    assert abs(rk.currentTime - tend) <= 1e-12, rk.currentTime
    got = e.download_state("position", "velocity", "massDensity", "specificThermalEnergy")
    pos, vel, rho, eps = (got[k][:N] for k in ("position", "velocity", "massDensity", "specificThermalEnergy"))
    E1 = float(np.sum(m*(0.5*np.sum(vel*vel, axis=1) + eps)))
    print("Noh 2-D on the device: %d steps (oracle %d), dE/E = %.3e" % (rk.currentCycle, gold["cycles"], (E1 - E0)/E0))
    assert abs(E1 - E0) <= 1e-12*abs(E0)
    assert abs(rk.currentCycle - gold["cycles"]) <= 2
    r = np.linalg.norm(pos, axis=1)
    bins = np.array(gold["bins"])
    which = np.digitize(r, bins) - 1
    prof = np.array([rho[which == b].mean() if np.any(which == b) else 0.0 for b in range(len(bins) - 1)])
    gp = np.array(gold["rho_profile"])
    print("rho profile (device):", np.round(prof, 4).tolist())
    assert np.abs(prof - gp).max() <= 1e-5*gp.max(), (prof.tolist(), gp.tolist())
    assert 12.0 < prof[3] < 16.5                                   # the compressed plateau behind the shock at r = 0.2
    assert abs(prof[6]/(1.0 + tend/0.325) - 1.0) < 0.06            # pre-shock: rho = 1 + t/r
    assert np.all(pos >= 0.0)


@pytest.mark.parametrize("ndim,n,nPerh", [(2, 20, 2.01), (3, 8, 1.51)])
def test_periodic_and_reflecting_boundaries(oracle, mods, ndim, n, nPerh):
    """PeriodicBoundary in x (two planar boundaries, plane1 -> plane2 and back; PeriodicBoundary.cc:60-62) combined with
    reflecting planes on the other axes: device ghost generation / refresh / enforcement against the numpy restatement
    (nodegen.boundary_ghosts), and the derivatives on that ghost set against the oracle."""
    engine, _ = mods
    st, nInt, _ = common.make_problem(ndim, n, nPerh=nPerh, seed=13)
    WT = K.TableKernel(K.BSplineKernel(ndim), 1000)
    ex = np.eye(ndim)[0]
    bnds = [("periodic", (np.zeros(ndim), ex), (ex.copy(), -ex))] + [("reflecting", (np.zeros(ndim), np.eye(ndim)[a])) for a in range(1, ndim)]
    f = {o: st[k] for k, o in PN.items()}
    ref, lists, n0 = ng.boundary_ghosts(ndim, f, bnds, WT.kernelExtent)
    e = engine.Engine(ndim, nPerh=nPerh)
    e.set_kernel_table(WT)
    e.set_nodes(nInt, 0)
    e.upload_state(**st)
    e.boundary_configure(bnds)
    ngh = e.reflect_set_ghost_nodes()
    assert ngh == ref["pos"].shape[0] - nInt and len(lists) == ndim + 1 and all(len(l) > 0 for l in lists)
    got = e.download_state(*PN.keys())
    for k, o in PN.items():
        assert np.abs(got[k] - ref[o]).max() <= 1e-14*max(np.abs(ref[o]).max(), 1e-300), k
    # periodic images are displaced by exactly one period and carry unreflected values
    g0 = slice(nInt, nInt + len(lists[0]))
    assert np.allclose(got["position"][g0] - got["position"][lists[0]], -ex, atol=1e-15)      # controls near x = 1 -> ghosts below x = 0
    assert np.array_equal(got["velocity"][g0], got["velocity"][lists[0]])
    # derivatives on that ghost set
    OT = common.oracle_table(oracle, WT)
    oo, po = common.opts_pair(oracle, engine, ndim, nPerh=nPerh)
    pi, pj, cnt = oracle.pairs(ndim, nInt, ngh, ref["pos"], ref["H"], OT.kext)
    d = oracle.evaluate_derivatives(oo, OT, dict(ref), nInt, ngh, pi, pj, cnt)
    assert e.build_pairs() == len(pi)
    e.evaluate_derivatives()
    g = e.download_derivs("DvDt", "DrhoDt", "DepsDt", "DvDx")
    floors = common.physical_floors({k: got[k] for k in PN}, nInt, ndim)
    for k in g:
        assert common.field_err(g[k], d[k], nInt, floors[k]) <= 1e-10, k
    # enforcement: a node pushed through the periodic plane re-enters on the other side with its velocity unchanged, one pushed
    # through a reflecting plane is mirrored
    pos, vel = got["position"].copy(), got["velocity"].copy()
    pos[5, 0] = -0.02
    pos[6, 0] = 1.03
    pos[7, 1] = -0.01
    e.upload_state(position=pos)
    assert e.reflect_enforce(count=True) == 3
    back = e.download_state("position", "velocity")
    assert abs(back["position"][5, 0] - 0.98) < 1e-15 and abs(back["position"][6, 0] - 0.03) < 1e-15
    assert abs(back["position"][7, 1] - 0.01) < 1e-16
    assert np.array_equal(back["velocity"][5], vel[5]) and np.array_equal(back["velocity"][6], vel[6])
    assert back["velocity"][7, 1] == -vel[7, 1] and back["velocity"][7, 0] == vel[7, 0]
