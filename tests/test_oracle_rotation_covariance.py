"""Rotational covariance of the oracle's 2-D and 3-D code.  The one numerical golden the reference stores for this path is a 1-D run
(tests/test_oracle_noh_planar_1d_golden.py), which exercises the `#if D == 1` branches of oracle/*_dim.inc; the tensor algebra of the
2-D / 3-D branches (H.r, dyads, tensor products, the symmetric-tensor component order xx xy xz yy yz zz, the matrix inverse of the
velocity-gradient correction, the ASPH tensor derivative, the Q's velocity-gradient tensors, the RK corrections) is pinned here by the
property every one of those expressions must have: rotate the whole problem by R (x -> R x, v -> R v, H -> R H R^T, DvDx_Q -> R DvDx_Q R^T)
and every scalar output is unchanged, every vector output is rotated, every tensor output is conjugated -- to round-off.  A transposed
index, a swapped symmetric component or a row/column mix-up anywhere in those branches breaks it at O(1)."""
import numpy as np
import pytest

import common
from spheral_b200 import nodegen as ng

SCALARS = ("DrhoDt", "DepsDt", "rhoSum", "normalization", "maxViscousPressure", "effViscousPressure", "XSPHWeightSum", "massZerothMoment")
VECTORS = ("DxDt", "DvDt", "gradRho", "XSPHDeltaV")
TENSORS = ("DvDx", "localDvDx", "M", "localM")
SYMS = ("DHDt", "Hideal")


def _rotate_state(ndim, s, R):
    out = dict(s)
    out["pos"] = np.ascontiguousarray(s["pos"] @ R.T)
    out["vel"] = np.ascontiguousarray(s["vel"] @ R.T)
    F = ng.sym_to_full(ndim, s["H"])
    out["H"] = np.ascontiguousarray(ng.full_to_sym(ndim, R @ F @ R.T))
    if "DvDxQ" in s:
        T = np.asarray(s["DvDxQ"]).reshape(-1, ndim, ndim)
        out["DvDxQ"] = np.ascontiguousarray((R @ T @ R.T).reshape(-1, ndim*ndim))
    return out


def _check(ndim, d0, d1, R, nInt, names_scalar, names_vec, names_ten, names_sym, tol=2.0e-11):
    def rel(a, b):
        return float(np.abs(a - b).max())/max(float(np.abs(b).max()), 1e-300)
    for k in names_scalar:
        if k in d0:
            assert rel(np.asarray(d1[k])[:nInt], np.asarray(d0[k])[:nInt]) <= tol, k
    for k in names_vec:
        if k in d0:
            assert rel(np.asarray(d1[k])[:nInt], np.asarray(d0[k])[:nInt] @ R.T) <= tol, k
    for k in names_ten:
        if k in d0:
            T0 = np.asarray(d0[k])[:nInt].reshape(-1, ndim, ndim)
            assert rel(np.asarray(d1[k])[:nInt].reshape(-1, ndim, ndim), R @ T0 @ R.T) <= tol, k
    for k in names_sym:
        if k in d0 and np.abs(np.asarray(d0[k])[:nInt]).max() > 0.0:
            F0 = ng.sym_to_full(ndim, np.asarray(d0[k])[:nInt])
            assert rel(ng.sym_to_full(ndim, np.asarray(d1[k])[:nInt]), R @ F0 @ R.T) <= tol, k


@pytest.mark.parametrize("ndim,n", [(2, 22), (3, 8)])
@pytest.mark.parametrize("mode", ["sph", "asph", "limited_q_balsara", "asph_classic"])
def test_sph_derivatives_are_rotationally_covariant(oracle, ndim, n, mode):
    rng = np.random.default_rng(17)
    kind = "lattice" if mode == "sph" else "aniso"
    nPerh = 2.01 if ndim == 2 else (1.51 if kind == "lattice" else 1.3)
    st, nInt, nGhost = common.make_problem(ndim, n, nPerh=nPerh, kind=kind, seed=29)
    if mode == "limited_q_balsara":
        st["DvDxQ"] = rng.normal(size=(nInt, ndim*ndim))
    s0 = common.to_oracle_state(st)
    if "DvDxQ" in st:
        s0["DvDxQ"] = st["DvDxQ"]
    hb = 1.0/st["H"][:nInt, 0].mean()
    okw = dict(sph=dict(hEvolution=oracle.H_SPH), asph=dict(hEvolution=oracle.H_ASPH), limited_q_balsara=dict(hEvolution=oracle.H_ASPH, Qkind=1, balsara=1),
               asph_classic=dict(hEvolution=oracle.H_ASPH_CLASSIC, hmin=0.02*hb, hmax=50.0*hb, hminratio=0.1))[mode]
    o = oracle.default_options(ndim, nPerh=2.01 if kind == "aniso" else nPerh, Cl=1.0, Cq=1.5, **okw)
    WT = oracle.TableKernel(oracle.KERNEL_BSPLINE, ndim, 1000)
    R = ng.random_rotation(ndim, rng)
    s1 = _rotate_state(ndim, s0, R)
    res = []
    for s in (s0, s1):
        pi, pj, cnt = oracle.pairs(ndim, nInt, nGhost, s["pos"], s["H"], WT.kext)
        res.append((oracle.evaluate_derivatives(o, WT, s, nInt, nGhost, pi, pj, cnt), pi, pj))
    (d0, pi0, pj0), (d1, pi1, pj1) = res
    assert np.array_equal(pi0, pi1) and np.array_equal(pj0, pj1)            # the pair predicate is rotation invariant (up to the last ulp of eta)
    _check(ndim, d0, d1, R, nInt, SCALARS, VECTORS + ("massFirstMoment",), TENSORS, SYMS)
    a0, a1 = np.asarray(d0["pairAccelerations"]), np.asarray(d1["pairAccelerations"])
    assert np.abs(a1 - a0 @ R.T).max() <= 2e-11*np.abs(a0).max()


@pytest.mark.parametrize("ndim,n", [(2, 20), (3, 7)])
def test_crksph_derivatives_are_rotationally_covariant(oracle, ndim, n):
    rng = np.random.default_rng(19)
    st, nInt, nGhost = common.make_problem(ndim, n, nPerh=2.01 if ndim == 2 else 1.3, kind="aniso", seed=37)
    st["DvDxQ"] = rng.normal(size=(nInt, ndim*ndim))
    s0 = common.to_oracle_state(st); s0["DvDxQ"] = st["DvDxQ"]
    o = oracle.default_options(ndim, nPerh=2.01, Cl=2.0, Cq=1.0, Qkind=1, hEvolution=oracle.H_ASPH)
    WT = oracle.TableKernel(oracle.KERNEL_BSPLINE, ndim, 1000)
    R = ng.random_rotation(ndim, rng)
    s1 = _rotate_state(ndim, s0, R)
    res = []
    for s in (s0, s1):
        pi, pj, cnt = oracle.pairs(ndim, nInt, nGhost, s["pos"], s["H"], WT.kext)
        vol = oracle.crk_sum_volume(ndim, WT, nInt, nGhost, s["pos"], s["H"], pi, pj)
        corr = oracle.crk_corrections(ndim, WT, nInt, nGhost, s["pos"], s["H"], vol, pi, pj)
        res.append((oracle.crk_evaluate_derivatives(o, WT, s, vol, corr, nInt, nGhost, pi, pj), vol, corr))
    (d0, v0, c0), (d1, v1, c1) = res
    assert np.abs(v1 - v0).max() <= 1e-12*np.abs(v0).max()
    # RK coefficients of linear order {A, B | grad A, grad B}: A invariant, B and grad A vectors, grad B a tensor (RKUtilities.cc:637-715)
    P = ndim + 1
    C0, C1 = np.asarray(c0).reshape(-1, P, P), np.asarray(c1).reshape(-1, P, P)
    assert np.abs(C1[:, 0, 0] - C0[:, 0, 0]).max() <= 1e-10*np.abs(C0[:, 0, 0]).max()
    assert np.abs(C1[:, 0, 1:] - C0[:, 0, 1:] @ R.T).max() <= 1e-10*np.abs(C0[:, 0, 1:]).max()
    assert np.abs(C1[:, 1:, 0] - C0[:, 1:, 0] @ R.T).max() <= 1e-10*np.abs(C0[:, 1:, 0]).max()
    assert np.abs(C1[:, 1:, 1:] - R @ C0[:, 1:, 1:] @ R.T).max() <= 1e-10*np.abs(C0[:, 1:, 1:]).max()
    _check(ndim, d0, d1, R, nInt, ("DrhoDt", "DepsDt", "maxViscousPressure", "effViscousPressure"), ("DxDt", "DvDt", "XSPHDeltaV"),
           ("DvDx", "localDvDx"), ("DHDt",), tol=1.0e-9)


@pytest.mark.parametrize("ndim,n", [(2, 22), (3, 8)])
def test_step_loops_and_state_update_are_rotationally_covariant(oracle, ndim, n):
    """Sum density and the grad-h correction are invariants; IncrementASPHHtensor (eigenvalue bounds of the updated tensor) commutes with
    the rotation."""
    rng = np.random.default_rng(23)
    st, nInt, _ = common.make_problem(ndim, n, nPerh=2.01 if ndim == 2 else 1.3, kind="aniso", seed=41)
    s0 = common.to_oracle_state(st)
    WT = oracle.TableKernel(oracle.KERNEL_BSPLINE, ndim, 1000)
    R = ng.random_rotation(ndim, rng)
    s1 = _rotate_state(ndim, s0, R)
    out = []
    for s in (s0, s1):
        pi, pj, cnt = oracle.pairs(ndim, nInt, 0, s["pos"], s["H"], WT.kext)
        rho = oracle.sum_mass_density(ndim, WT, nInt, 0, s["pos"], s["mass"], s["H"], pi, pj)
        om = oracle.omega_gradh(ndim, WT, nInt, 0, s["pos"], s["H"], pi, pj, cnt)
        out.append((np.asarray(rho).copy(), np.asarray(om).copy()))
    assert np.abs(out[1][0] - out[0][0]).max() <= 1e-12*np.abs(out[0][0]).max()
    assert np.abs(out[1][1] - out[0][1]).max() <= 1e-12*np.abs(out[0][1]).max()
    # eigenvalue clamp of a symmetric tensor (the H policies): bound(R H R^T) = R bound(H) R^T
    F = ng.sym_to_full(ndim, s0["H"])
    lam = np.linalg.eigvalsh(F)
    lo, hi = float(np.quantile(lam, 0.3)), float(np.quantile(lam, 0.7))
    for k in range(0, nInt, max(1, nInt//40)):
        a = s0["H"][k].copy(); b = s1["H"][k].copy()
        oracle.sym_bound(ndim, a, lo, hi); oracle.sym_bound(ndim, b, lo, hi)
        Fa, Fb = ng.sym_to_full(ndim, a[None])[0], ng.sym_to_full(ndim, b[None])[0]
        assert np.abs(Fb - R @ Fa @ R.T).max() <= 1e-11*np.abs(Fa).max()


@pytest.mark.parametrize("ndim,n", [(2, 22), (3, 8)])
def test_translation_and_galilean_invariance(oracle, ndim, n):
    """Only differences of positions and velocities enter the pair loop: a shifted, uniformly moving copy of the problem has the same
    derivatives (DxDt moves with the frame)."""
    st, nInt, nGhost = common.make_problem(ndim, n, nPerh=2.01 if ndim == 2 else 1.3, kind="aniso", seed=43)
    s0 = common.to_oracle_state(st)
    WT = oracle.TableKernel(oracle.KERNEL_BSPLINE, ndim, 1000)
    o = oracle.default_options(ndim, nPerh=2.01, hEvolution=oracle.H_ASPH, Cl=1.0, Cq=1.5)
    shift, boost = np.array([0.375, -1.25, 2.5][:ndim]), np.array([0.5, -0.25, 0.125][:ndim])      # exactly representable: differences stay exact to an ulp
    s1 = dict(s0, pos=np.ascontiguousarray(s0["pos"] + shift), vel=np.ascontiguousarray(s0["vel"] + boost))
    res = []
    for s in (s0, s1):
        pi, pj, cnt = oracle.pairs(ndim, nInt, nGhost, s["pos"], s["H"], WT.kext)
        res.append((oracle.evaluate_derivatives(o, WT, s, nInt, nGhost, pi, pj, cnt), pi, pj))
    (d0, pi0, pj0), (d1, pi1, pj1) = res
    assert np.array_equal(pi0, pi1) and np.array_equal(pj0, pj1)
    for k in SCALARS + ("DvDt", "gradRho", "XSPHDeltaV") + TENSORS + ("DHDt",):
        a, b = np.asarray(d1[k])[:nInt], np.asarray(d0[k])[:nInt]
        assert np.abs(a - b).max() <= 1e-11*max(np.abs(b).max(), 1e-300), k
    assert np.abs(np.asarray(d1["DxDt"])[:nInt] - (np.asarray(d0["DxDt"])[:nInt] + boost)).max() <= 1e-12
