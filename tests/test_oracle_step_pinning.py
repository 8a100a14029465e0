"""Pin the oracle's per-step callers (oracle/step_oracle_dim.inc) before the GPU is compared to them.

The reference holds no golden vectors for these loops either; they are pinned through the properties its own tests rely on:
  * sum density of a uniform lattice reproduces rho0 (tests/functional/Hydro/Noh: sumDensity start-up), and equals a direct
    O(N^2) evaluation of the definition,
  * the grad-h correction of a uniform distribution is 1 (continuum identity, computeSPHOmegaGradhCorrection.cc:103-106),
  * eigenvalue clamps agree with LAPACK (numpy.linalg.eigh), and leave in-bounds tensors bit-identical,
  * GenericHydro::dt against an independent vectorised restatement,
  * total energy conserved to round-off over CheapSynchronousRK2 steps with compatibleEnergyEvolution
    (tests/functional/Hydro/Noh/Noh-cylindrical-2d.py:803-808).
"""
import numpy as np
import pytest

import common
from spheral_b200 import kernel as K, nodegen as ng


def _table(oracle, ndim):
    return common.oracle_table(oracle, K.TableKernel(K.BSplineKernel(ndim), 1000))


@pytest.mark.parametrize("ndim,n,nPerh", [(3, 10, 1.51), (2, 24, 2.01)])
def test_sum_density_matches_definition_and_lattice_density(oracle, ndim, n, nPerh):
    OT = _table(oracle, ndim)
    pos, mass, H, d = ng.lattice(ndim, n, nPerh=nPerh)
    N = pos.shape[0]
    pi, pj, cnt = oracle.pairs(ndim, N, 0, pos, H, OT.kext)
    rho = oracle.sum_mass_density(ndim, OT, N, 0, pos, mass, H, pi, pj)
    # definition, brute force: rho_i = sum_j m_j W(|H_j (r_i - r_j)|) det H_j   (self term included)
    F = ng.sym_to_full(ndim, H)
    ref = np.zeros(N)
    for j in range(N):
        eta = np.linalg.norm((pos - pos[j]) @ F[j].T, axis=1)
        ref += mass[j]*np.array([OT.kernelAndGradValue(e, np.linalg.det(F[j]))[0] for e in eta])
    assert np.abs(rho - ref).max() <= 1e-12*np.abs(ref).max()
    inner = np.all((pos > 0.3) & (pos < 0.7), axis=1)
    assert np.abs(rho[inner] - 1.0).max() < 2e-3        # unit density lattice


@pytest.mark.parametrize("ndim,n,nPerh", [(3, 12, 1.51), (2, 30, 2.01)])
def test_omega_gradh_is_one_on_uniform_lattice(oracle, ndim, n, nPerh):
    OT = _table(oracle, ndim)
    pos, mass, H, d = ng.lattice(ndim, n, nPerh=nPerh)
    N = pos.shape[0]
    pi, pj, cnt = oracle.pairs(ndim, N, 0, pos, H, OT.kext)
    om = oracle.omega_gradh(ndim, OT, N, 0, pos, H, pi, pj, cnt)
    inner = np.all((pos > 0.3) & (pos < 0.7), axis=1)
    assert np.abs(om[inner] - 1.0).max() < 2e-2
    # an isolated node gets exactly 1 (computeSPHOmegaGradhCorrection.cc:99-100)
    pos2 = np.vstack([pos, np.full((1, ndim), 50.0)]); H2 = np.vstack([H, H[:1]])
    pi, pj, cnt = oracle.pairs(ndim, N + 1, 0, pos2, H2, OT.kext)
    assert oracle.omega_gradh(ndim, OT, N + 1, 0, pos2, H2, pi, pj, cnt)[-1] == 1.0


@pytest.mark.parametrize("ndim", [2, 3])
def test_sym_bound_against_lapack(oracle, ndim):
    rng = np.random.default_rng(5)
    for _ in range(50):
        A = rng.standard_normal((ndim, ndim))
        F = A @ A.T + 0.1*np.eye(ndim)
        Hs = ng.full_to_sym(ndim, F[None])[0]
        lam, V = np.linalg.eigh(F)
        # in bounds: bit-identical
        assert np.array_equal(oracle.sym_bound(ndim, Hs, 0.5*lam.min(), 2.0*lam.max()), Hs)
        lo, hi = 0.5*(lam[0] + lam[1]), 0.5*(lam[-2] + lam[-1])
        if lo > hi:
            lo, hi = hi, lo
        ref = (V*np.clip(lam, lo, hi)) @ V.T
        got = ng.sym_to_full(ndim, oracle.sym_bound(ndim, Hs, lo, hi)[None])[0]
        assert np.abs(got - ref).max() <= 1e-12*np.abs(ref).max()


def _dt_numpy(oo, so, N, s, d, pi, pj, ndim):
    tiny = np.finfo(float).eps
    F = ng.sym_to_full(ndim, s["H"])
    scale = 1.0/np.linalg.eigvalsh(F).max(axis=1)/oo.nPerh
    vmag = np.linalg.norm(s["vel"], axis=1)
    amag = np.linalg.norm(d["DvDt"], axis=1)
    div = np.trace(d["DvDx"].reshape(N, ndim, ndim), axis1=1, axis2=2)
    cands = [scale/(s["cs"] + tiny), scale/(np.sqrt(d["maxViscousPressure"]/s["rho"]) + tiny), 1.0/(np.abs(div) + tiny),
             0.1*np.maximum(scale/(vmag + tiny), vmag/(amag + tiny))]
    vij = np.linalg.norm(s["vel"][pi] - s["vel"][pj], axis=1)
    cands.append(np.minimum(scale[pi], scale[pj])/np.maximum(tiny, vij))
    return so.cfl*min(float(c.min()) for c in cands)


@pytest.mark.parametrize("ndim,n,kind", [(3, 9, "lattice"), (2, 20, "lattice"), (3, 8, "aniso")])
def test_hydro_dt_against_independent_restatement(oracle, ndim, n, kind):
    nPerh = 1.51 if ndim == 3 else 2.01
    st, nInt, nGhost = common.make_problem(ndim, n, nPerh=nPerh, kind=kind)
    OT = _table(oracle, ndim)
    oo = oracle.default_options(ndim, nPerh=nPerh, hEvolution=oracle.H_ASPH if kind == "aniso" else oracle.H_SPH)
    so = oracle.default_step_options()
    s = common.to_oracle_state(st)
    pi, pj, cnt = oracle.pairs(ndim, nInt, 0, s["pos"], s["H"], OT.kext)
    d = oracle.evaluate_derivatives(oo, OT, s, nInt, 0, pi, pj, cnt)
    dt, why, node = oracle.hydro_dt(oo, so, nInt, s["vel"], s["H"], s["rho"], s["cs"], d, pi, pj)
    ref = _dt_numpy(oo, so, nInt, s, d, pi, pj, ndim)
    assert abs(dt - ref) <= 1e-12*ref
    assert why in oracle.DT_REASONS and 0 <= node < nInt


def test_state_update_policies(oracle):
    ndim, n, nPerh = 3, 7, 1.51
    st, nInt, _ = common.make_problem(ndim, n, nPerh=nPerh, kind="aniso")
    OT = _table(oracle, ndim)
    s = common.to_oracle_state(st); s["eps"] = st["specificThermalEnergy"]
    pi, pj, cnt = oracle.pairs(ndim, nInt, 0, s["pos"], s["H"], OT.kext)
    oo = oracle.default_options(ndim, nPerh=nPerh, hEvolution=oracle.H_ASPH, compatibleEnergy=0, hmin=1e-3, hmax=1e3)
    so = oracle.default_step_options(rhoMin=0.9, rhoMax=1.2)
    d = oracle.evaluate_derivatives(oo, OT, s, nInt, 0, pi, pj, cnt)
    mult = 1e-3
    out = oracle.state_update(oo, so, nInt, 0, mult, False, d, s)
    assert np.allclose(out["pos"], s["pos"] + mult*d["DxDt"], rtol=0, atol=1e-15)
    assert np.allclose(out["vel"], s["vel"] + mult*d["DvDt"], rtol=0, atol=1e-14)
    assert np.array_equal(out["rho"], np.clip(s["rho"] + mult*d["DrhoDt"], 0.9, 1.2))
    assert np.allclose(out["eps"], s["eps"] + mult*d["DepsDt"], rtol=1e-15)
    # ASPH increment: H + mult*DHDt, eigenvalues clipped to [1/hmax, min(1/hmin, lam_min/hminratio)]
    F = ng.sym_to_full(ndim, s["H"] + mult*d["DHDt"])
    lam, V = np.linalg.eigh(F)
    top = np.minimum(1e3, np.maximum(1e-3, lam[:, :1])/so.hminratio)
    ref = np.einsum("nik,nk,njk->nij", V, np.clip(lam, 1e-3, top), V)
    assert np.abs(ng.sym_to_full(ndim, out["H"]) - ref).max() <= 1e-12*np.abs(ref).max()
    P, cs = ng.gamma_law(out["rho"], out["eps"], so.gamma)
    assert np.allclose(out["P"], P, rtol=1e-15) and np.allclose(out["cs"], cs, rtol=1e-15)
    # SPH IdealH: replace by Hideal unless timeAdvanceOnly, where it degrades to an increment
    oo2 = oracle.default_options(ndim, nPerh=nPerh, hEvolution=oracle.H_SPH, compatibleEnergy=0)
    st2, nInt2, _ = common.make_problem(ndim, n, nPerh=nPerh)
    s2 = common.to_oracle_state(st2); s2["eps"] = st2["specificThermalEnergy"]
    pi, pj, cnt = oracle.pairs(ndim, nInt2, 0, s2["pos"], s2["H"], OT.kext)
    d2 = oracle.evaluate_derivatives(oo2, OT, s2, nInt2, 0, pi, pj, cnt)
    so2 = oracle.default_step_options()
    assert np.array_equal(oracle.state_update(oo2, so2, nInt2, 0, mult, False, d2, s2)["H"], d2["Hideal"])
    assert np.array_equal(oracle.state_update(oo2, so2, nInt2, 0, mult, True, d2, s2)["H"], s2["H"] + mult*d2["DHDt"])


@pytest.mark.parametrize("ndim,n,nPerh", [(2, 20, 2.01), (3, 8, 1.51)])
def test_rk2_conserves_total_energy(oracle, ndim, n, nPerh):
    """Compatible energy: |dE/E| at round-off over CheapSynchronousRK2 steps (Noh-cylindrical-2d.py:803-808, 1e-13 per run)."""
    st, nInt, _ = common.make_problem(ndim, n, nPerh=nPerh)
    st["velocity"] = 0.3*st["velocity"]
    OT = _table(oracle, ndim)
    oo = oracle.default_options(ndim, nPerh=nPerh, Cl=1.0, Cq=1.0)
    so = oracle.default_step_options()
    # IntegrateDensity: the sum-density replacement leaves the mass, velocity and eps -- hence E -- untouched either way
    rk = common.OracleRK2(oracle, oo, so, OT, st, densityUpdate=0)
    rk.initializeDerivatives()
    E0 = rk.total_energy()
    P0 = (rk.s["mass"][:, None]*rk.s["vel"]).sum(axis=0)
    for _ in range(4):
        rk.step()
    assert rk.t > 0 and rk.cycle == 4
    assert abs(rk.total_energy() - E0) <= 1e-13*abs(E0)
    P1 = (rk.s["mass"][:, None]*rk.s["vel"]).sum(axis=0)
    assert np.abs(P1 - P0).max() <= 1e-13*np.abs(rk.s["mass"][:, None]*rk.s["vel"]).sum()
