"""The host-side mirror of the reference's Physics package interface (spheral_b200/physics.py): field keys, factory
defaults and error behaviour (CPU), and one evaluateDerivatives(time, dt, dataBase, state, derivs) call against the
oracle (GPU) written the way the reference's own unit test drives the package
(tests/unit/SPH/testLinearVelocityGradient.py:268-292)."""
import numpy as np
import pytest

import common
from spheral_b200 import kernel as K


def _setup(ndim=3, n=9, nPerh=1.51, **hydro_kw):
    from spheral_b200 import physics as P
    st, nInt, nGhost = common.make_problem(ndim, n, nPerh=nPerh, seed=77)
    nodes = P.FluidNodeList("nodes", ndim, nInt, nGhost, nPerh=nPerh)
    for abi in ("position", "velocity", "H", "mass", "massDensity", "specificThermalEnergy"):
        nodes.setField(P.STATE_KEYS[abi], st[abi])
    db = P.DataBase()
    db.appendNodeList(nodes)
    WT = K.TableKernel(K.BSplineKernel(ndim), 1000)
    hydro = P.SPH(dataBase=db, W=WT, **hydro_kw)
    return P, st, nodes, db, WT, hydro


def test_factory_defaults_and_registration_contract(sphlib):
    P, st, nodes, db, WT, hydro = _setup()
    # SPHHydros.py:88-94 -- default Q is LimitedMonaghanGingold with Cl = 2(kext/2), Cq = 2(kext/2)^2
    assert isinstance(hydro.Q, P.LimitedMonaghanGingoldViscosity) and hydro.Q.Cl == 2.0 and hydro.Q.Cq == 2.0
    assert hydro.preSubPackages() == [hydro.Q]
    assert isinstance(hydro.postSubPackages()[0], P.SPHSmoothingScale)
    assert hydro.requireConnectivity() and not hydro.requireGhostConnectivity()
    assert hydro.compatibleEnergyEvolution and hydro.XSPH and hydro.correctVelocityGradient and not hydro.evolveTotalEnergy
    assert isinstance(P.ASPH(WT, dataBase=db).postSubPackages()[0], P.ASPHSmoothingScale)
    state, derivs = P.State(db, [hydro]), P.StateDerivatives(db, [hydro])
    for key in ("mass", "position", "velocity", "mass density", "specific thermal energy", "H", "pressure", "sound speed",
                "grad h corrections", "time step mask", "velocity gradient for artificial viscosity"):
        assert state.registered(key, "nodes"), key
    for key in ("delta position", "delta mass density", "delta velocity hydro", "delta specific thermal energy",
                "velocity gradient", "internal velocity gradient", "mass density gradient", "M SPH gradient correction",
                "new mass density", "normalization", "XSPH weight sum", "XSPH delta vi", "max viscous pressure",
                "effective viscous pressure", "delta H", "new H", "mass zeroth moment", "mass first moment"):
        assert derivs.registered(key, "nodes"), key
    assert "pair-wise accelerations" in derivs
    assert state.field("position", "nodes") is nodes.positions()            # enrolled by reference, not copied
    assert state.policies["specific thermal energy"] == "SpecificThermalEnergyPolicy"


def test_compatible_and_total_energy_are_exclusive(sphlib):
    from spheral_b200 import physics as P
    with pytest.raises(P.SPHB200Error, match="cannot simultaneously"):
        _setup(compatibleEnergyEvolution=True, evolveTotalEnergy=True)


@pytest.mark.gpu
def test_evaluate_derivatives_through_physics_interface(sphlib, oracle):
    P, st, nodes, db, WT, hydro = _setup(Q=None)
    hydro.Q = P.MonaghanGingoldViscosity(2.0, 2.0)
    hydro.initializeProblemStartup(db)
    state, derivs = P.State(db, [hydro]), P.StateDerivatives(db, [hydro])
    state.field("pressure", "nodes")[...] = st["pressure"]
    state.field("sound speed", "nodes")[...] = st["soundSpeed"]
    state.field("grad h corrections", "nodes")[...] = st["omegaGradh"]
    with pytest.raises(P.SPHB200Error, match="no valid connectivity"):
        hydro.evaluateDerivatives(0.0, 1.0, db, state, derivs)
    npairs = hydro.updateConnectivity(db, state)
    derivs.Zero()
    hydro.evaluateDerivatives(0.0, 1.0, db, state, derivs)

    nInt = nodes.numInternalNodes
    oo = oracle.default_options(3, nPerh=1.51, Cl=2.0, Cq=2.0)
    s = common.to_oracle_state(st)
    pi, pj, cnt = oracle.pairs(3, nInt, 0, s["pos"], s["H"], WT.kernelExtent)
    ref = oracle.evaluate_derivatives(oo, common.oracle_table(oracle, WT), s, nInt, 0, pi, pj, cnt)
    assert npairs == len(pi)
    floors = common.physical_floors(st, nInt, 3)
    for abi, key in P.DERIV_KEYS.items():
        got = derivs.field(key, "nodes")
        assert common.field_err(got, np.asarray(ref[abi]).reshape(got.shape), nInt, floors[abi]) <= 1e-10, key
    pa = np.asarray(hydro.pairAccelerations)
    assert pa.shape == (npairs, 3) and np.abs(pa - ref["pairAccelerations"]).max() <= 1e-10*np.abs(ref["pairAccelerations"]).max()
    # compatible energy through the policy hook: total energy is conserved by construction
    eps0 = state.field("specific thermal energy", "nodes").copy()
    dtc = 1.0e-3
    hydro.updateSpecificThermalEnergy(dtc, db, state, derivs)
    eps1 = state.field("specific thermal energy", "nodes")
    m, v, a = st["mass"], st["velocity"], derivs.field("delta velocity hydro", "nodes")
    dKE = (m*((v + dtc*a)**2).sum(axis=1)).sum()*0.5 - (m*(v**2).sum(axis=1)).sum()*0.5
    dTE = (m*(eps1 - eps0)).sum()
    assert abs(dKE + dTE) <= 1e-12*max(abs(dKE), abs(dTE), 1e-300)


# ---- CRKSPH through the same interface (CRKSPH/CRKSPHHydros.py, RK/RKCorrections.cc, CRKSPH/CRKSPHBase.cc) -------------------
def _setup_crk(ndim=3, n=8, nPerh=1.51, **hydro_kw):
    from spheral_b200 import physics as P
    st, nInt, nGhost = common.make_problem(ndim, n, nPerh=nPerh, seed=79)
    nodes = P.FluidNodeList("nodes", ndim, nInt, nGhost, nPerh=nPerh)
    for abi in ("position", "velocity", "H", "mass", "massDensity", "specificThermalEnergy"):
        nodes.setField(P.STATE_KEYS[abi], st[abi])
    db = P.DataBase()
    db.appendNodeList(nodes)
    WT = K.TableKernel(K.BSplineKernel(ndim), 1000)
    hydro = P.CRKSPH(dataBase=db, W=WT, **hydro_kw)
    return P, st, nodes, db, WT, hydro


def test_crksph_factory_defaults_and_registration_contract(sphlib):
    P, st, nodes, db, WT, hydro = _setup_crk()
    # CRKSPHHydros.py:64-68 -- default Q is LimitedMonaghanGingold with Cl = 2(kext/4), Cq = (kext/4)^2
    assert isinstance(hydro.Q, P.LimitedMonaghanGingoldViscosity) and hydro.Q.Cl == 1.0 and hydro.Q.Cq == 0.25
    assert hydro.label() == "CRKSPH" and hydro.correctionOrder == P.RKOrder.LinearOrder
    assert hydro.requireReproducingKernels() == {P.RKOrder.ZerothOrder, P.RKOrder.LinearOrder}
    assert hydro.preSubPackages() == [hydro.Q] and isinstance(hydro.postSubPackages()[0], P.SPHSmoothingScale)
    assert isinstance(P.ACRKSPH(db, WT).postSubPackages()[0], P.ASPHSmoothingScale)
    state, derivs = P.State(db, [hydro]), P.StateDerivatives(db, [hydro])
    for key in ("mass", "position", "velocity", "mass density", "specific thermal energy", "H", "pressure", "sound speed",
                "volume", "rkCorrections_1", "velocity gradient for artificial viscosity"):
        assert state.registered(key, "nodes"), key
    for key in ("delta position", "delta mass density", "delta velocity hydro", "delta specific thermal energy",
                "velocity gradient", "internal velocity gradient", "XSPH delta vi", "max viscous pressure",
                "effective viscous pressure", "delta H", "new H"):
        assert derivs.registered(key, "nodes"), key
    assert "pair-wise accelerations" in derivs
    with pytest.raises(P.SPHB200Error, match="RKSumVolume"):
        P.CRKSPH(db, WT, volumeType=P.RKVoronoiVolume)


@pytest.mark.gpu
def test_crksph_hooks_against_oracle(sphlib, oracle):
    P, st, nodes, db, WT, hydro = _setup_crk(Q=None, densityUpdate="IntegrateDensity")
    hydro.Q = P.MonaghanGingoldViscosity(1.0, 0.5)
    hydro.initializeProblemStartup(db)
    state, derivs = P.State(db, [hydro]), P.StateDerivatives(db, [hydro])
    state.field("pressure", "nodes")[...] = st["pressure"]
    state.field("sound speed", "nodes")[...] = st["soundSpeed"]
    npairs = hydro.updateConnectivity(db, state)
    hydro.preStepInitialize(db, state, derivs)                     # volumes
    assert hydro.initialize(0.0, 1.0, db, state, derivs) is True   # corrections; True = re-apply ghost boundaries
    derivs.Zero()
    hydro.evaluateDerivatives(0.0, 1.0, db, state, derivs)

    nInt = nodes.numInternalNodes
    OT = common.oracle_table(oracle, WT)
    oo = oracle.default_options(3, nPerh=1.51, Cl=1.0, Cq=0.5)
    s = common.to_oracle_state(st)
    pi, pj, cnt = oracle.pairs(3, nInt, 0, s["pos"], s["H"], WT.kernelExtent)
    assert npairs == len(pi)
    vol = oracle.crk_sum_volume(3, OT, nInt, 0, s["pos"], s["H"], pi, pj)
    corr = oracle.crk_corrections(3, OT, nInt, 0, s["pos"], s["H"], vol, pi, pj)
    ref = oracle.crk_evaluate_derivatives(oo, OT, s, vol, corr, nInt, 0, pi, pj)
    assert np.abs(state.field("volume", "nodes") - vol).max() <= 1e-10*vol.max()
    floors = common.physical_floors(st, nInt, 3)
    for abi in ("DxDt", "DrhoDt", "DvDt", "DepsDt", "DvDx", "localDvDx", "maxViscousPressure", "effViscousPressure",
                "XSPHDeltaV", "DHDt", "Hideal"):
        got = derivs.field(P.DERIV_KEYS[abi], "nodes")
        assert common.field_err(got, np.asarray(ref[abi]).reshape(got.shape), nInt, floors[abi]) <= 1e-10, abi
    pa = np.asarray(hydro.pairAccelerations)
    assert np.abs(pa - ref["pairAccelerations"]).max() <= 1e-10*np.abs(ref["pairAccelerations"]).max()
    # RigorousSumDensity branch of preStepInitialize
    hydro.densityUpdate = P.RigorousSumDensity
    hydro.preStepInitialize(db, state, derivs)
    rho = oracle.crk_sum_density(3, OT, nInt, 0, s["pos"], s["mass"], vol, s["H"], pi, pj, rhoMin=nodes.rhoMin, rhoMax=nodes.rhoMax)
    assert np.abs(state.field("mass density", "nodes") - rho).max() <= 1e-10*rho.max()


@pytest.mark.gpu
def test_step_hooks_through_physics_interface(sphlib, oracle):
    """preStepInitialize (sum density), postStateUpdate (grad-h correction) and dt() driven the way Integrator does
    (Integrator.cc:177-183, 252-271, 114-166), against the oracle."""
    P, st, nodes, db, WT, hydro = _setup(Q=None)
    hydro.Q = P.MonaghanGingoldViscosity(2.0, 2.0)
    hydro.initializeProblemStartup(db)
    state, derivs = P.State(db, [hydro]), P.StateDerivatives(db, [hydro])
    state.field("pressure", "nodes")[...] = st["pressure"]
    state.field("sound speed", "nodes")[...] = st["soundSpeed"]
    state.field("grad h corrections", "nodes")[...] = st["omegaGradh"]
    hydro.updateConnectivity(db, state)
    nInt = nodes.numInternalNodes
    OT = common.oracle_table(oracle, WT)
    s = common.to_oracle_state(st)
    pi, pj, cnt = oracle.pairs(3, nInt, 0, s["pos"], s["H"], WT.kernelExtent)
    # preStepInitialize: RigorousSumDensity is the factory default
    hydro.preStepInitialize(db, state, derivs)
    rho_ref = oracle.sum_mass_density(3, OT, nInt, 0, s["pos"], s["mass"], s["H"], pi, pj)
    rho = state.field("mass density", "nodes")
    assert np.abs(rho - rho_ref).max() <= 1e-10*np.abs(rho_ref).max()
    # postStateUpdate: grad-h corrections, returns True (boundaries must be re-applied)
    assert hydro.postStateUpdate(0.0, 1.0, db, state, derivs) is True
    om_ref = oracle.omega_gradh(3, OT, nInt, 0, s["pos"], s["H"], pi, pj, cnt)
    om = state.field("grad h corrections", "nodes")
    assert np.abs(om - om_ref).max() <= 1e-10*np.abs(om_ref).max()
    # dt after one evaluation
    derivs.Zero()
    hydro.evaluateDerivatives(0.0, 1.0, db, state, derivs)
    s2 = dict(s, rho=rho_ref, omega=om_ref)
    oo = oracle.default_options(3, nPerh=1.51, Cl=2.0, Cq=2.0)
    d = oracle.evaluate_derivatives(oo, OT, s2, nInt, 0, pi, pj, cnt)
    ref_dt, why, node = oracle.hydro_dt(oo, oracle.default_step_options(cfl=hydro.cfl), nInt, s2["vel"], s2["H"], s2["rho"], s2["cs"], d, pi, pj)
    vote, reason = hydro.dt(db, state, derivs, 0.0)
    assert abs(vote - ref_dt) <= 1e-11*ref_dt and reason.lower().startswith(why)


@pytest.mark.gpu
def test_mid_step_refresh_keeps_the_step_start_connectivity(sphlib, oracle):
    """CheapSynchronousRK2.cc:76-99: state.update moves the nodes, postStateUpdate and evaluateDerivatives then run on the
    ConnectivityMap of the step start.  The package re-reads the State's (moved) positions without rebuilding the pair lists, sees
    host arrays that were modified IN PLACE (no markDirty, no new array identity), and evaluates on the old pairs -- the oracle does
    the same with the step-start pair list and the moved state."""
    P, st, nodes, db, WT, hydro = _setup(Q=None)
    hydro.Q = P.MonaghanGingoldViscosity(2.0, 2.0)
    hydro.initializeProblemStartup(db)
    state, derivs = P.State(db, [hydro]), P.StateDerivatives(db, [hydro])
    state.field("pressure", "nodes")[...] = st["pressure"]
    state.field("sound speed", "nodes")[...] = st["soundSpeed"]
    state.field("grad h corrections", "nodes")[...] = st["omegaGradh"]
    npairs = hydro.updateConnectivity(db, state)
    nInt = nodes.numInternalNodes
    s = common.to_oracle_state(st)
    pi, pj, cnt = oracle.pairs(3, nInt, 0, s["pos"], s["H"], WT.kernelExtent)
    assert npairs == len(pi)
    # the trial advance: positions and velocities change in place
    pos = state.field("position", "nodes")
    pos += 0.02*st["velocity"]/np.abs(st["velocity"]).max()*(1.0/9.0)
    state.field("velocity", "nodes")[...] *= 1.01
    assert hydro.postStateUpdate(0.0, 1.0, db, state, derivs) is True
    assert hydro._engine.connectivity_valid() and hydro._engine.npairs == npairs          # no rebuild happened
    OT = common.oracle_table(oracle, WT)
    s2 = dict(s, pos=np.ascontiguousarray(pos), vel=np.ascontiguousarray(state.field("velocity", "nodes")))
    om_ref = oracle.omega_gradh(3, OT, nInt, 0, s2["pos"], s2["H"], pi, pj, cnt)
    om = state.field("grad h corrections", "nodes")
    assert np.abs(om - om_ref).max() <= 1e-10*np.abs(om_ref).max()
    derivs.Zero()
    hydro.evaluateDerivatives(0.0, 1.0, db, state, derivs)                                # must not raise: connectivity is still the step-start one
    oo = oracle.default_options(3, nPerh=1.51, Cl=2.0, Cq=2.0)
    ref = oracle.evaluate_derivatives(oo, OT, dict(s2, omega=om_ref), nInt, 0, pi, pj, cnt)
    floors = common.physical_floors(st, nInt, 3)
    for abi in ("DvDt", "DepsDt", "DrhoDt", "DvDx"):
        got = derivs.field(P.DERIV_KEYS[abi], "nodes")
        assert common.field_err(got, np.asarray(ref[abi]).reshape(got.shape), nInt, floors[abi]) <= 1e-10, abi
