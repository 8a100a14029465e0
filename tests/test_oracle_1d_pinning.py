"""CPU tests that pin the 1-D instantiation of the oracle (oracle/*_dim.inc with D = 1) against the reference's own 1-D
known-answer properties.  The GPU engine is 2-D / 3-D; the 1-D oracle exists so that the restated step can be checked against the
one stored numerical golden the reference holds for this path (Noh-planar-1d.py:226-240, a 1-D run).  The
dimension-specific code of the oracle is a handful of `#if D == ...` blocks, everything else is shared with 2-D / 3-D.
"""
import numpy as np
import pytest

from spheral_b200 import nodegen as ng

NBSPLINE5 = 105          # oracle.KERNEL_NBSPLINE + order


def _fuzzy(a, b, tol):
    return abs(a - b) <= tol*max(1.0, abs(a) + abs(b))


# ---- tests/unit/Kernel/testTableKernel.py:74-90 in 1-D, plus NBSplineKernel(5) (the kernel of the Noh-planar-1d golden run) -------
@pytest.mark.parametrize("kind,kext", [(0, 2.0), (1, 1.0), (2, 1.0), (NBSPLINE5, 3.0)])
def test_table_matches_analytic_kernel_1d(oracle, kind, kext):
    WT = oracle.TableKernel(kind, 1, 100, with_nperh=False)
    assert WT.kext == kext
    for eta in np.linspace(0.0, kext, 1000):
        W, g = WT.kernelAndGradValue(float(eta), 1.0)
        Wa, ga, _ = oracle.kernel_analytic(kind, 1, float(eta))
        assert _fuzzy(W, Wa, 1.0e-3) and _fuzzy(g, ga, 1.0e-2)


@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_nbspline5_is_the_normalised_quintic_spline(oracle, ndim):
    """NBSplineKernel.cc:17-122: order 5 is Schoenberg's quintic B-spline with support 3; the numerically integrated volume
    normalisation (simpsonsVolumeIntegral, 10000 bins) must reproduce the closed forms 1/120, 7/(478 pi), 1/(120 pi)."""
    closed = {1: 1.0/120.0, 2: 7.0/(478.0*np.pi), 3: 1.0/(120.0*np.pi)}[ndim]
    W0, g0, _ = oracle.kernel_analytic(NBSPLINE5, ndim, 0.0)
    assert abs(W0 - 66.0*closed) <= 1e-9*W0 and g0 == 0.0
    for eta in (0.3, 1.0, 1.7, 2.5):
        W, g, g2 = oracle.kernel_analytic(NBSPLINE5, ndim, eta)
        q = [(max(0.0, 3.0 - eta))**5, (max(0.0, 2.0 - eta))**5, (max(0.0, 1.0 - eta))**5]
        assert abs(W - closed*(q[0] - 6.0*q[1] + 15.0*q[2])) <= 1e-9*W0
        h = 1e-5                                   # derivatives consistent with the value
        Wp, Wm = oracle.kernel_analytic(NBSPLINE5, ndim, eta + h)[0], oracle.kernel_analytic(NBSPLINE5, ndim, eta - h)[0]
        assert abs(g - (Wp - Wm)/(2*h)) <= 1e-6*W0
    assert oracle.kernel_analytic(NBSPLINE5, ndim, 3.0)[0] == 0.0


# ---- tests/unit/Neighbor/NeighborTestBase.py:192-258, 1-D instantiation (testNestedGridNeighbor / testTreeNeighbor 1-D cases) --------
def test_cell_pairs_equal_bruteforce_random_1d(oracle):
    n = 600
    rng = np.random.default_rng(4599281941)
    pos = rng.uniform(0.0, 1.0, size=(n, 1))
    H = (1.0/(2.01*rng.uniform(0.5, 2.0, size=(n, 1))*(1.0/n)))
    bi, bj, bc = oracle.pairs(1, n, 0, pos, H, 2.0, "brute")
    ci, cj, cc = oracle.pairs(1, n, 0, pos, H, 2.0, "cells")
    assert len(bi) > n
    assert np.array_equal(bi, ci) and np.array_equal(bj, cj) and np.array_equal(bc, cc)
    assert np.all(bi < bj)
    # SpheralTestUtilities.findNeighborNodes: min(|Hi rij|, |Hj rij|) <= kext
    for i in rng.integers(0, n, 12):
        rij = np.abs(pos[i, 0] - pos[:, 0])
        nb = set(np.nonzero(np.minimum(H[i, 0]*rij, H[:, 0]*rij) <= 2.0)[0].tolist()) - {int(i)}
        got = set(bj[bi == i].tolist()) | set(bi[bj == i].tolist())
        assert nb == got and bc[i] == len(nb)


# ---- tests/unit/SPH/testLinearVelocityGradient.py, the 1-D case (two NodeLists side by side, jitter 0.2 dx, tolerance 5e-5) ------------
def test_linear_velocity_gradient_is_exact_with_M_correction_1d(oracle):
    nx1 = nx2 = 10
    nPerh = 2.01
    p1, m1, H1, d1 = ng.lattice(1, nx1, [0.0], [0.5], 1.0, nPerh)
    p2, m2, H2, d2 = ng.lattice(1, nx2, [0.5], [1.0], 1.0, nPerh)
    pos = np.concatenate([p1, p2]); mass = np.concatenate([m1, m2]); H = np.concatenate([H1, H2])
    pos = ng.jitter_python_random(pos, 0.2, d1, seed=14892042)
    N = len(pos)
    vel = 1.0 + 1.0*pos
    rho = np.ones(N); eps = np.zeros(N)
    P, cs = ng.gamma_law(rho, eps)
    st = dict(pos=pos, vel=vel, H=H, mass=mass, rho=rho, P=P, cs=cs, omega=np.ones(N))
    WT = oracle.TableKernel(oracle.KERNEL_BSPLINE, 1, 1000)
    pi, pj, cnt = oracle.pairs(1, N, 0, pos, H, WT.kext)
    err = {}
    for corr in (0, 1):
        o = oracle.default_options(1, nPerh=nPerh, correctVelocityGradient=corr, Cl=2.0, Cq=2.0)
        d = oracle.evaluate_derivatives(o, WT, st, N, 0, pi, pj, cnt)
        interior = cnt > 2                                   # SPH.cc:500: the correction needs more than 2^nDim neighbours
        err[corr] = float(((np.asarray(d["DvDx"]).reshape(N) - 1.0)**2)[interior].max())
    assert err[1] <= 5.0e-5 and err[1] < 1.0e-20, err
    assert err[0] > 1.0e-4


def test_momentum_and_energy_contracts_1d(oracle):
    """Sum m DvDt = 0 and the pair-acceleration contract (SpecificThermalEnergyPolicy.cc:84-107) hold in the 1-D instantiation."""
    N = 80
    pos, mass, H, d = ng.lattice(1, N, nPerh=1.35)
    pos = ng.jitter(pos, 0.15, d, seed=3)
    rng = np.random.default_rng(8)
    vel = 0.3*np.sin(6.0*pos) + 0.05*rng.standard_normal((N, 1))
    rho = 1.0 + 0.1*np.cos(4.0*pos[:, 0]); eps = 1.0 + 0.2*rng.uniform(size=N)
    P, cs = ng.gamma_law(rho, eps)
    st = dict(pos=pos, vel=vel, H=H, mass=mass, rho=rho, P=P, cs=cs, omega=np.ones(N))
    WT = oracle.TableKernel(NBSPLINE5, 1, 1000)
    pi, pj, cnt = oracle.pairs(1, N, 0, pos, H, WT.kext)
    o = oracle.default_options(1, nPerh=1.35, compatibleEnergy=1)
    dd = oracle.evaluate_derivatives(o, WT, st, N, 0, pi, pj, cnt)
    a = np.asarray(dd["DvDt"]).reshape(N)
    assert abs(float((mass*a).sum())) < 1e-13*np.abs(mass*a).sum()
    chk = np.zeros(N)
    pa = np.asarray(dd["pairAccelerations"]).reshape(-1)
    np.add.at(chk, pi, pa)
    np.add.at(chk, pj, -pa*(mass[pi]/mass[pj]))
    assert np.abs(chk - a).max() < 1e-12*np.abs(a).max()
