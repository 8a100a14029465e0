"""Pinning of the oracle's restatement of ASPHClassicSmoothingScale::evaluateDerivatives (SmoothingScale/ASPHClassicSmoothingScale.cc:130-378,
selected by ASPH = "Classic" in SPHHydros.py:133-134).  The reference stores no Cartesian golden for it (its ATS uses are RZ / solid runs), so
the C restatement is held against (i) an independent numpy restatement of the same formulae with LAPACK's eigen-solver in place of the oracle's
Jacobi rotations, from brute-force moments, and (ii) properties of the algorithm: on an undisturbed lattice the second moment is isotropic and
the ideal H is the SPH one; the size of the ideal H (its determinant) is the SPH ideal H's whatever the shape; a lattice stretched along x
gets an ideal H whose smoothing length is longest along x."""
import numpy as np
import pytest

import common
from spheral_b200 import nodegen as ng


def _table_grad(OT, eta):
    """|gradW| of the oracle's quadratic-interpolated table, Hdet = 1 (TableKernelViewInline.hh:118-127)."""
    eta = np.asarray(eta)
    k = np.minimum((np.maximum(0.0, eta - OT.xmin)/OT.xstep).astype(np.int64), OT.n1)
    c = OT.gradWcoef
    g = c[3*k] + (c[3*k + 1] + c[3*k + 2]*eta)*eta
    return np.where(eta < OT.kext, np.abs(g), 0.0)


def _sym(ndim, v):
    if ndim == 2:
        return np.array([[v[0], v[1]], [v[1], v[2]]])
    return np.array([[v[0], v[1], v[2]], [v[1], v[3], v[4]], [v[2], v[4], v[5]]])


def _unsym(ndim, M):
    return np.array([M[0, 0], M[0, 1], M[1, 1]]) if ndim == 2 else np.array([M[0, 0], M[0, 1], M[0, 2], M[1, 1], M[1, 2], M[2, 2]])


def numpy_classic(orc, OT, ndim, pos, H, pi, pj, nInt, nPerh, hmin, hmax, hminratio):
    """The whole package in numpy: pair moments, then the per-node ideal H (eigh instead of the oracle's eigen routine)."""
    n = pos.shape[0]
    m0 = np.zeros(n); m2 = np.zeros((n, ndim, ndim))
    Hm = np.array([_sym(ndim, h) for h in H])
    xij = pos[pi] - pos[pj]
    for (a, b, sgn) in ((pi, pj, 1.0), (pj, pi, -1.0)):
        eta = np.einsum("kab,kb->ka", Hm[a], xij)
        W = _table_grad(OT, np.sqrt((eta*eta).sum(axis=1)))
        np.add.at(m0, a, W)
        r = np.sqrt((xij*xij).sum(axis=1))
        dy = np.einsum("ka,kb->kab", xij, xij)/np.maximum(r**5, 1e-300)[:, None, None]
        np.add.at(m2, a, (W*W)[:, None, None]*dy)
    out = np.zeros((nInt, H.shape[1]))
    tiny = 1e-50
    for i in range(nInt):
        z0 = max(0.0, m0[i])**(1.0/ndim)
        cur = 0.5*nPerh if abs(z0) <= 1e-15*max(1.0, abs(z0)) else max(0.0, OT.equivalentNodesPerSmoothingScale(z0))
        s = min(4.0, max(0.25, nPerh/cur))
        psiweight = max(0.0, min(1.0, 2.0/s - 1.0))
        M = m2[i]
        if psiweight > 0.0 and np.linalg.det(M) > 0.0 and np.linalg.eigvalsh(M).min() > 0.0:
            psi = M/np.abs(M).max()
            d = np.linalg.det(psi)
            psi = psi/(abs(d) + tiny)**(1.0/ndim) if d > 1e-10 else np.eye(ndim)
            lam, V = np.linalg.eigh(psi)
            lam = 1.0/np.sqrt(lam)
            lam = np.maximum(lam, lam.max()*hminratio)
            psi = (V*lam) @ V.T
            psi = psi/(np.linalg.det(psi) + tiny)**(1.0/ndim)
            lam, V = np.linalg.eigh(psi)
            Hid = np.linalg.inv((V*np.sqrt(lam)) @ V.T)
        else:
            Hid = np.eye(ndim)
        a = 0.4*(1.0 + s*s) if s < 1.0 else 0.4*(1.0 + 1.0/(s**3 + tiny))
        Hid = Hid*np.linalg.det(Hm[i])**(1.0/ndim)/(1.0 - a + a*s)
        lam, V = np.linalg.eigh(Hid)
        hminEffInv = min(1.0/hmin, max(1.0/hmax, lam.min())/hminratio)
        lam = np.maximum(1.0/hmax, np.minimum(hminEffInv, lam))
        out[i] = _unsym(ndim, (V*lam) @ V.T)
    return out, m0


def _run(orc, ndim, st, nInt, nGhost, nPerh, mode, **kw):
    OT = orc.TableKernel(0, ndim, 1000)
    s = common.to_oracle_state(st)
    pi, pj, cnt = orc.pairs(ndim, nInt, nGhost, s["pos"], s["H"], OT.kext)
    oo = orc.default_options(ndim, nPerh=nPerh, hEvolution=mode, **kw)
    return orc.evaluate_derivatives(oo, OT, s, nInt, nGhost, pi, pj, cnt), OT, s, (pi, pj)


@pytest.mark.parametrize("ndim,n,kind", [(2, 18, "aniso"), (3, 7, "aniso"), (2, 20, "lattice"), (3, 8, "lattice")])
def test_oracle_equals_an_independent_numpy_restatement(oracle, ndim, n, kind):
    nPerh = 2.01 if ndim == 2 else 1.51
    st, nInt, nGhost = common.make_problem(ndim, n, nPerh=nPerh if kind == "lattice" else (2.01 if ndim == 2 else 1.3), kind=kind, seed=31)
    hb = 1.0/st["H"][:nInt, 0].mean()
    kw = dict(hmin=0.05*hb, hmax=20.0*hb, hminratio=0.1)
    d, OT, s, (pi, pj) = _run(oracle, ndim, st, nInt, nGhost, 2.01 if kind == "aniso" else nPerh, oracle.H_ASPH_CLASSIC, **kw)
    ref, m0 = numpy_classic(oracle, OT, ndim, s["pos"], s["H"], pi, pj, nInt, 2.01 if kind == "aniso" else nPerh, **kw)
    got = np.asarray(d["Hideal"])[:nInt]
    assert np.abs(got - ref).max() <= 1e-11*np.abs(ref).max()
    z0 = np.maximum(0.0, m0[:nInt])**(1.0/ndim)
    assert np.abs(np.asarray(d["massZerothMoment"])[:nInt] - z0).max() <= 1e-12*z0.max()
    if kind == "aniso":
        assert np.abs(got[:, 1]).max() > 1e-3*np.abs(got[:, 0]).max()          # a genuine tensor, not a disguised scalar


@pytest.mark.parametrize("ndim,n", [(2, 24), (3, 10)])
def test_lattice_gives_the_sph_ideal_h(oracle, ndim, n):
    """An undisturbed lattice has an isotropic second moment: the classic ideal H of an interior node is the SPH ideal H."""
    nPerh = 2.01 if ndim == 2 else 1.51
    pos, mass, H, dx = ng.lattice(ndim, n, nPerh=nPerh)
    st, nInt, nGhost = common.make_problem(ndim, n, nPerh=nPerh, seed=3)
    st["position"] = pos
    dS, _, s, _ = _run(oracle, ndim, st, nInt, nGhost, nPerh, oracle.H_SPH, hmin=1e-10, hmax=1e10)
    dC, _, _, _ = _run(oracle, ndim, st, nInt, nGhost, nPerh, oracle.H_ASPH_CLASSIC, hmin=1e-10, hmax=1e10, hminratio=0.1)
    ext = 2.0*nPerh*dx
    interior = np.all((pos > ext) & (pos < 1.0 - ext), axis=1)[:nInt]
    assert interior.sum() > 0
    a, b = np.asarray(dS["Hideal"])[:nInt][interior], np.asarray(dC["Hideal"])[:nInt][interior]
    assert np.abs(a - b).max() <= 1e-12*np.abs(a).max()


@pytest.mark.parametrize("ndim,n", [(2, 24), (3, 10)])
def test_size_follows_the_sph_rule_and_shape_follows_the_stretch(oracle, ndim, n):
    nPerh = 2.01 if ndim == 2 else 1.51
    st, nInt, nGhost = common.make_problem(ndim, n, nPerh=nPerh, seed=5)
    dS, _, _, _ = _run(oracle, ndim, st, nInt, nGhost, nPerh, oracle.H_SPH, hmin=1e-10, hmax=1e10)
    dC, _, _, _ = _run(oracle, ndim, st, nInt, nGhost, nPerh, oracle.H_ASPH_CLASSIC, hmin=1e-10, hmax=1e10, hminratio=1e-3)
    detS = np.array([np.linalg.det(_sym(ndim, h)) for h in np.asarray(dS["Hideal"])[:nInt]])
    detC = np.array([np.linalg.det(_sym(ndim, h)) for h in np.asarray(dC["Hideal"])[:nInt]])
    assert np.abs(detC/detS - 1.0).max() <= 1e-10                            # unit-determinant shape times the SPH size factor
    # the same nodes stretched by 1.6 along x: interior nodes see their neighbours further away along x, h must grow most along x
    st2 = {k: v.copy() for k, v in st.items()}
    st2["position"][:, 0] *= 1.6
    d2, _, s2, _ = _run(oracle, ndim, st2, nInt, nGhost, nPerh, oracle.H_ASPH_CLASSIC, hmin=1e-10, hmax=1e10, hminratio=1e-3)
    pos = st2["position"]
    ext = 2.5*nPerh/n
    interior = np.all((pos > ext*np.array([1.6] + [1.0]*(ndim - 1))) & (pos < (1.0 - ext)*np.array([1.6] + [1.0]*(ndim - 1))), axis=1)[:nInt]
    Hid = np.asarray(d2["Hideal"])[:nInt][interior]
    assert interior.sum() > 0
    yy = 2 if ndim == 2 else 3
    assert np.median(Hid[:, 0]/Hid[:, yy]) < 0.9                              # H_xx < H_yy: the smoothing length is longer along x


@pytest.mark.parametrize("ndim,n", [(2, 16), (3, 7)])
def test_the_package_is_the_same_behind_either_hydro(oracle, ndim, n):
    """The smoothing-scale sub-package does not know which hydro it follows: except for DHDt (which reads the hydro's DvDx) its outputs
    behind the CRKSPH restatement equal those behind the SPH one."""
    st, nInt, nGhost = common.make_problem(ndim, n, nPerh=2.01 if ndim == 2 else 1.3, kind="aniso", seed=33)
    hb = 1.0/st["H"][:nInt, 0].mean()
    kw = dict(hmin=0.05*hb, hmax=20.0*hb, hminratio=0.1)
    dS, OT, s, (pi, pj) = _run(oracle, ndim, st, nInt, nGhost, 2.01, oracle.H_ASPH_CLASSIC, **kw)
    vol = oracle.crk_sum_volume(ndim, OT, nInt, nGhost, s["pos"], s["H"], pi, pj)
    corr = oracle.crk_corrections(ndim, OT, nInt, nGhost, s["pos"], s["H"], vol, pi, pj)
    oo = oracle.default_options(ndim, nPerh=2.01, hEvolution=oracle.H_ASPH_CLASSIC, **kw)
    dC = oracle.crk_evaluate_derivatives(oo, OT, s, vol, corr, nInt, nGhost, pi, pj)
    for k in ("Hideal", "massZerothMoment", "massFirstMoment"):
        a, b = np.asarray(dS[k])[:nInt], np.asarray(dC[k])[:nInt]
        assert np.abs(a - b).max() <= 1e-13*max(np.abs(a).max(), 1e-300), k
