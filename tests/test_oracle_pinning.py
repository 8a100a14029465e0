"""CPU tests that PIN the oracle (oracle/sph_oracle.c) against the reference's own known-answer properties.

The reference holds no per-node golden derivative vectors for this path, so the restatement is pinned
by the properties the reference's tests assert, each cited below.
"""
import math
import numpy as np
import pytest

import common
from spheral_b200 import nodegen as ng


# ---- tests/unit/Kernel/testTableKernel.py:74-90 (testWlookup: W0tol 1e-3, W1tol 1e-2, fuzzyEqual) --------------------
def _fuzzy(a, b, tol):
    return abs(a - b) <= tol*max(1.0, abs(a) + abs(b))


@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_table_matches_analytic_kernel(oracle, ndim, kind):
    WT = oracle.TableKernel(kind, ndim, 100, with_nperh=False)
    assert WT.kext == (2.0 if kind == 0 else 1.0)
    nsamples = 1000
    deta = WT.kext/(nsamples - 1)
    for i in range(nsamples):
        eta = i*deta
        W, g = WT.kernelAndGradValue(eta, 1.0)
        Wa, ga, _ = oracle.kernel_analytic(kind, ndim, eta)
        assert _fuzzy(W, Wa, 1.0e-3)
        assert _fuzzy(g, ga, 1.0e-2)


def test_table_1000_points_is_tight(oracle):
    # tests/cpp/Kernel/tablekernel_tests.cc:57-75 compares at float precision for a fine table
    WT = oracle.TableKernel(oracle.KERNEL_BSPLINE, 3, 1000, with_nperh=False)
    for eta in np.linspace(0.0, 1.999, 777):
        W, g = WT.kernelAndGradValue(float(eta), 1.0)
        Wa, ga, _ = oracle.kernel_analytic(0, 3, float(eta))
        assert abs(W - Wa) < 2e-7 and abs(g - ga) < 2e-5


# ---- testTableKernel.py:95-106 (testMonotonicity) + round trip nperh(Wsum(nperh)) -------------------------------------
@pytest.mark.parametrize("ndim", [2, 3])
def test_nperh_lookup_monotonic_and_invertible(oracle, ndim):
    WT = oracle.TableKernel(oracle.KERNEL_BSPLINE, ndim, 100)
    xs = np.linspace(WT.wsumRange[0], WT.wsumRange[1], 200)
    ws = np.array([WT.equivalentWsum(float(x)) for x in xs])
    assert np.all(np.diff(ws) > 0.0)
    ys = np.linspace(WT.nperhRange[0], WT.nperhRange[1], 200)
    ns = np.array([WT.equivalentNodesPerSmoothingScale(float(y)) for y in ys])
    assert np.all(np.diff(ns) > 0.0)
    for nperh in (1.0, 1.51, 2.01, 4.01):
        back = WT.equivalentNodesPerSmoothingScale(WT.equivalentWsum(nperh))
        assert abs(back - nperh) < 0.05*nperh


# ---- tests/unit/Neighbor/NeighborTestBase.py:192-258: connectivity == brute force, incl. rotated anisotropic H ---------
@pytest.mark.parametrize("ndim,n", [(2, 700), (3, 1200)])
def test_cell_pairs_equal_bruteforce_random_anisotropic(oracle, ndim, n):
    pos, H = ng.random_anisotropic(ndim, n, [[0.0, 1.0]]*ndim, nPerh=2.01, seed=4599281940 + ndim)
    bi, bj, bc = oracle.pairs(ndim, n, 0, pos, H, 2.0, "brute")
    ci, cj, cc = oracle.pairs(ndim, n, 0, pos, H, 2.0, "cells")
    assert len(bi) > n
    assert np.array_equal(bi, ci) and np.array_equal(bj, cj) and np.array_equal(bc, cc)
    # pair list is sorted by (i, j) with i < j   (NodePairIdxType::operator<, ConnectivityMap.cc:1023-1029)
    assert np.all(bi < bj)
    key = bi.astype(np.int64)*n + bj
    assert np.all(np.diff(key) > 0)
    # SpheralTestUtilities.findNeighborNodes: min(|Hi.rij|, |Hj.rij|) <= kext, checked for sampled nodes
    F = ng.sym_to_full(ndim, H)
    for i in np.random.default_rng(1).integers(0, n, 10):
        rij = pos[i] - pos
        ei = np.linalg.norm(rij @ F[i].T, axis=1)
        ej = np.linalg.norm(np.einsum("nab,nb->na", F, rij), axis=1)
        nb = set(np.nonzero(np.minimum(ei, ej) <= 2.0)[0].tolist()) - {int(i)}
        got = set(bj[bi == i].tolist()) | set(bi[bj == i].tolist())
        assert nb == got
        assert bc[i] == len(nb)


def test_pairs_with_ghosts_orientation(oracle):
    st, nInt, nGhost = common.make_problem(2, 14, nPerh=2.01, ghosts=True)
    assert nGhost > 0
    pi, pj, cnt = oracle.pairs(2, nInt, nGhost, st["position"], st["H"], 2.0, "cells")
    bi, bj, bc = oracle.pairs(2, nInt, nGhost, st["position"], st["H"], 2.0, "brute")
    assert np.array_equal(pi, bi) and np.array_equal(pj, bj) and np.array_equal(cnt, bc)
    assert pi.max() < nInt            # the i node of every pair is internal (ConnectivityMapInline.hh:292-309)
    assert (pj >= nInt).any()         # ghost neighbours appear on the j side only


# ---- tests/unit/SPH/testLinearVelocityGradient.py (--nx1 10 --nx2 10, linear case, tolerance 5e-5) --------------------
@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("asph", [False, True])
def test_linear_velocity_gradient_is_exact_with_M_correction(oracle, ndim, asph):
    nx1 = nx2 = 10
    nPerh = 2.01
    lo, hi = np.zeros(ndim), np.ones(ndim)
    hi1 = hi.copy(); hi1[0] = 0.5
    lo2 = lo.copy(); lo2[0] = 0.5
    n1 = (nx1,) + (nx1 + nx2,)*(ndim - 1)
    p1, m1, H1, d1 = ng.lattice(ndim, n1, lo, hi1, 1.0, nPerh)
    p2, m2, H2, d2 = ng.lattice(ndim, n1, lo2, hi, 1.0, nPerh)
    pos = np.concatenate([p1, p2]); mass = np.concatenate([m1, m2]); H = np.concatenate([H1, H2])
    pos = ng.jitter_python_random(pos, 0.2, d1, seed=14892042)
    N = len(pos)
    if asph:   # the ASPH variant of the test ends up with tensor H; emulate with a mild random anisotropy
        rng = np.random.default_rng(5)
        F = ng.sym_to_full(ndim, H)
        for i in range(N):
            R = ng.random_rotation(ndim, rng)
            F[i] = R @ np.diag(np.diag(F[i])*rng.uniform(0.8, 1.25, ndim)) @ R.T
        H = ng.full_to_sym(ndim, F)
    y0, m0 = 1.0, 1.0
    vel = y0 + m0*pos
    rho = np.ones(N); eps = np.zeros(N)
    P, cs = ng.gamma_law(rho, eps)
    st = dict(pos=pos, vel=vel, H=H, mass=mass, rho=rho, P=P, cs=cs, omega=np.ones(N))
    WT = oracle.TableKernel(oracle.KERNEL_BSPLINE, ndim, 1000)
    pi, pj, cnt = oracle.pairs(ndim, N, 0, pos, H, WT.kext)
    err = {}
    for corr in (0, 1):
        o = oracle.default_options(ndim, nPerh=nPerh, correctVelocityGradient=corr, Cl=2.0, Cq=2.0,
                                   hEvolution=oracle.H_ASPH if asph else oracle.H_SPH)
        d = oracle.evaluate_derivatives(o, WT, st, N, 0, pi, pj, cnt)
        diff = d["DvDx"] - m0*np.eye(ndim).ravel()
        err[corr] = float((diff**2).sum(axis=1).max())
    assert err[1] <= 5.0e-5, err        # the reference's pass criterion
    assert err[1] < 1.0e-20             # and in fact exact to round-off
    assert err[0] > 1.0e-4              # while the uncorrected estimate is not


# ---- conservation properties (Noh-cylindrical-2d.py:803-808; SpecificThermalEnergyPolicy.cc:84-107 contract) ----------
@pytest.mark.parametrize("ndim,kind", [(2, "lattice"), (3, "lattice"), (3, "aniso")])
def test_momentum_and_pair_acceleration_contract(oracle, ndim, kind):
    st, nInt, nGhost = common.make_problem(ndim, 9 if ndim == 3 else 24, nPerh=1.51 if ndim == 3 else 2.01, kind=kind)
    WT = oracle.TableKernel(oracle.KERNEL_BSPLINE, ndim, 1000)
    s = common.to_oracle_state(st)
    pi, pj, cnt = oracle.pairs(ndim, nInt, 0, s["pos"], s["H"], WT.kext)
    o = oracle.default_options(ndim, nPerh=1.51)
    d = oracle.evaluate_derivatives(o, WT, s, nInt, 0, pi, pj, cnt)
    m = s["mass"]
    mom = (m[:, None]*d["DvDt"]).sum(axis=0)
    scale = np.abs(m[:, None]*d["DvDt"]).sum()
    assert np.abs(mom).max() < 1e-13*scale
    # DvDt_check(i) += paccij ; DvDt_check(j) -= paccij*mi/mj   must reproduce DvDt
    chk = np.zeros_like(d["DvDt"])
    pa = d["pairAccelerations"]
    np.add.at(chk, pi, pa)
    np.add.at(chk, pj, -pa*(m[pi]/m[pj])[:, None])
    assert np.abs(chk - d["DvDt"]).max() < 1e-12*np.abs(d["DvDt"]).max()


@pytest.mark.parametrize("ndim", [2, 3])
def test_compatible_energy_conserves_total_energy(oracle, ndim):
    st, nInt, _ = common.make_problem(ndim, 9 if ndim == 3 else 24, nPerh=1.51 if ndim == 3 else 2.01)
    WT = oracle.TableKernel(oracle.KERNEL_BSPLINE, ndim, 1000)
    s = common.to_oracle_state(st)
    pi, pj, cnt = oracle.pairs(ndim, nInt, 0, s["pos"], s["H"], WT.kext)
    o = oracle.default_options(ndim, nPerh=1.51, compatibleEnergy=1)
    d = oracle.evaluate_derivatives(o, WT, s, nInt, 0, pi, pj, cnt)
    m, v0, eps0 = s["mass"], s["vel"], st["specificThermalEnergy"]
    dt = 1.0e-3
    eps1 = oracle.update_energy_compatible(ndim, nInt, 0, m, v0, d["DvDt"], d["DepsDt"], pi, pj,
                                           d["pairAccelerations"], dt, eps0)
    v1 = v0 + dt*d["DvDt"]
    E0 = (m*(0.5*(v0**2).sum(axis=1) + eps0)).sum()
    E1 = (m*(0.5*(v1**2).sum(axis=1) + eps1)).sum()
    assert abs(E1 - E0)/abs(E0) < 1.0e-13
    # whereas the non-compatible update does not conserve to round-off
    E1n = (m*(0.5*(v1**2).sum(axis=1) + eps0 + dt*d["DepsDt"])).sum()
    assert abs(E1n - E0)/abs(E0) > 1.0e-12


def test_threaded_oracle_matches_serial(oracle):
    st, nInt, _ = common.make_problem(3, 9, nPerh=1.51)
    WT = oracle.TableKernel(oracle.KERNEL_BSPLINE, 3, 1000)
    s = common.to_oracle_state(st)
    pi, pj, cnt = oracle.pairs(3, nInt, 0, s["pos"], s["H"], WT.kext)
    o = oracle.default_options(3, nPerh=1.51)
    a = oracle.evaluate_derivatives(o, WT, s, nInt, 0, pi, pj, cnt, nthreads=1)
    b = oracle.evaluate_derivatives(o, WT, s, nInt, 0, pi, pj, cnt, nthreads=4)
    for k in a:
        assert np.allclose(a[k], b[k], rtol=1e-11, atol=1e-13*max(1.0, np.abs(a[k]).max())), k
    assert np.array_equal(a["pairAccelerations"], b["pairAccelerations"])


def test_compatible_and_total_energy_are_exclusive(oracle):
    st, nInt, _ = common.make_problem(2, 8, nPerh=2.01)
    WT = oracle.TableKernel(oracle.KERNEL_BSPLINE, 2, 100)
    s = common.to_oracle_state(st)
    pi, pj, cnt = oracle.pairs(2, nInt, 0, s["pos"], s["H"], WT.kext)
    o = oracle.default_options(2, compatibleEnergy=1, evolveTotalEnergy=1)
    with pytest.raises(RuntimeError, match="cannot simultaneously"):      # SPH.cc:97-98
        oracle.evaluate_derivatives(o, WT, s, nInt, 0, pi, pj, cnt)


def test_asph_dhdt_reduces_to_sph_for_isotropic_expansion(oracle):
    # for H = h^-1 I and DvDx = a I both forms give DHDt = -a H  (SPHSmoothingScale.cc:240 vs SmoothingScaleUtilities.hh:58-85)
    for ndim in (2, 3):
        n = 8
        pos, mass, H, d = ng.lattice(ndim, n, nPerh=2.01)
        N = len(pos)
        vel = 0.7*pos
        rho = np.ones(N); P, cs = ng.gamma_law(rho, np.ones(N))
        s = dict(pos=pos, vel=vel, H=H, mass=mass, rho=rho, P=P, cs=cs, omega=np.ones(N))
        WT = oracle.TableKernel(oracle.KERNEL_BSPLINE, ndim, 1000)
        pi, pj, cnt = oracle.pairs(ndim, N, 0, pos, H, WT.kext)
        res = {}
        for hev in (oracle.H_SPH, oracle.H_ASPH):
            o = oracle.default_options(ndim, hEvolution=hev, nPerh=2.01)
            res[hev] = oracle.evaluate_derivatives(o, WT, s, N, 0, pi, pj, cnt)["DHDt"]
        assert np.allclose(res[oracle.H_SPH], res[oracle.H_ASPH], rtol=0, atol=1e-10*np.abs(res[oracle.H_SPH]).max())
        assert np.allclose(res[oracle.H_SPH], -0.7*H, rtol=1e-9)
