"""CPU tests that PIN the CRKSPH part of the oracle (oracle/crk_oracle_dim.inc) against the reference's own known-answer
properties (the reference holds no per-node golden vectors for this path either):

* tests/functional/RK/RKInterpolation.py:5-6,94 -- linear-order RK interpolation reproduces a linear function and its gradient
  to 1e-12 (2-D and 3-D; RKMassOverDensity and RKSumVolume volumes);
* CRKSPH.cc:376-384 -- the Type-III force is pair-antisymmetric, so sum(m DvDt) = 0;
* tests/functional/Hydro/Noh/Noh-cylindrical-2d.py:803-808 -- compatible energy conserves total energy to round-off.
"""
import numpy as np
import pytest

import common
from spheral_b200 import nodegen as ng


def _setup(oracle, ndim, n, nPerh, seed, kind=0):
    st, nInt, nGhost = common.make_problem(ndim, n, nPerh=nPerh, seed=seed)
    W = oracle.TableKernel(kind, ndim, 200)
    s = common.to_oracle_state(st)
    pi, pj, cnt = oracle.pairs(ndim, nInt, nGhost, s["pos"], s["H"], W.kext)
    return st, s, W, nInt, nGhost, pi, pj, cnt


def _adjacency(nInt, pi, pj):
    adj = [[] for _ in range(nInt)]
    for a, b in zip(pi.tolist(), pj.tolist()):
        adj[a].append(b)
        if b < nInt:
            adj[b].append(a)
    return adj


@pytest.mark.parametrize("ndim,n", [(2, 14), (3, 8)])
@pytest.mark.parametrize("volume", ["mass_over_density", "sum"])
def test_rk_linear_interpolation_is_exact(oracle, ndim, n, volume):
    st, s, W, nInt, nGhost, pi, pj, cnt = _setup(oracle, ndim, n, 2.01 if ndim == 2 else 1.7, seed=5)
    pos, H = s["pos"], s["H"]
    if volume == "sum":
        vol = oracle.crk_sum_volume(ndim, W, nInt, nGhost, pos, H, pi, pj)
    else:
        vol = s["mass"]/s["rho"]
    corr = oracle.crk_corrections(ndim, W, nInt, nGhost, pos, H, vol, pi, pj)
    rng = np.random.default_rng(1)
    a0, a = rng.standard_normal(), rng.standard_normal(ndim)
    f = a0 + pos @ a
    adj = _adjacency(nInt, pi, pj)
    # interior nodes only (full kernel support is not required for RK, but keep the check generic): all of them
    worst_v = worst_g = 0.0
    for i in rng.choice(nInt, size=min(40, nInt), replace=False):
        val = 0.0
        grad = np.zeros(ndim)
        for j in adj[i] + [i]:
            WR, gWR = oracle.rk_kernel_grad(ndim, W, pos[i] - pos[j], H[j], corr[i])
            val += vol[j]*f[j]*WR
            grad += vol[j]*f[j]*gWR
        worst_v = max(worst_v, abs(val - f[i])/max(1.0, abs(f[i])))
        worst_g = max(worst_g, np.abs(grad - a).max()/max(1.0, np.abs(a).max()))
    assert worst_v < 1.0e-12, worst_v          # RKInterpolation.py:94 tolerance
    assert worst_g < 1.0e-10, worst_g          # gradient carries 1/h: same relative accuracy on the moments


@pytest.mark.parametrize("ndim,n", [(2, 16), (3, 9)])
def test_sum_volume_is_bounded_and_near_lattice_volume(oracle, ndim, n):
    st, s, W, nInt, nGhost, pi, pj, cnt = _setup(oracle, ndim, n, 1.51 if ndim == 3 else 2.01, seed=9)
    vol = oracle.crk_sum_volume(ndim, W, nInt, nGhost, s["pos"], s["H"], pi, pj)
    assert np.all(vol > 0)
    # computeRKSumVolume.cc:46,113: never above the eta-space cap
    Hdet = np.array([np.linalg.det(ng.sym_to_full(ndim, h)) for h in s["H"]])
    cap = (0.5*W.kext)**ndim*(np.pi if ndim == 2 else 4.0/3.0*np.pi)/Hdet
    assert np.all(vol <= cap*(1 + 1e-14))
    # interior nodes of a jittered lattice: 1/sum(W) is close to the cell volume
    d = 1.0/n
    inner = np.all((s["pos"] > 0.3) & (s["pos"] < 0.7), axis=1)
    assert abs(np.median(vol[inner])/d**ndim - 1.0) < 0.05


@pytest.mark.parametrize("ndim,n,Qkind", [(2, 16, 0), (3, 9, 0), (3, 8, 1)])
def test_crk_momentum_and_compatible_energy_conservation(oracle, ndim, n, Qkind):
    st, s, W, nInt, nGhost, pi, pj, cnt = _setup(oracle, ndim, n, 1.51 if ndim == 3 else 2.01, seed=21)
    if Qkind == 1:
        st = common.add_q_fields(st, ndim)
        s = common.to_oracle_state(st)
    o = oracle.default_options(ndim, nPerh=1.51 if ndim == 3 else 2.01, Qkind=Qkind, Cl=1.0, Cq=0.75)
    vol = oracle.crk_sum_volume(ndim, W, nInt, nGhost, s["pos"], s["H"], pi, pj)
    corr = oracle.crk_corrections(ndim, W, nInt, nGhost, s["pos"], s["H"], vol, pi, pj)
    d = oracle.crk_evaluate_derivatives(o, W, s, vol, corr, nInt, nGhost, pi, pj)
    m = s["mass"]
    # momentum: the pair force is equal and opposite (CRKSPH.cc:376-381)
    mom = (m[:, None]*d["DvDt"]).sum(axis=0)
    scale = np.abs(m[:, None]*d["DvDt"]).sum()
    assert np.abs(mom).max() <= 1.0e-12*scale
    # pairAccelerations are the i-side accelerations of each pair
    acc = np.zeros_like(d["DvDt"])
    pa = d["pairAccelerations"]
    np.add.at(acc, pi, pa)
    np.add.at(acc, pj, -pa*(m[pi]/m[pj])[:, None])
    assert np.abs(acc - d["DvDt"]).max() <= 1.0e-12*np.abs(d["DvDt"]).max()
    # compatible energy: total energy change of one forward-Euler step is zero to round-off
    dt = 1.0e-3
    eps0 = st["specificThermalEnergy"]
    eps1 = oracle.update_energy_compatible(ndim, nInt, nGhost, m, s["vel"], d["DvDt"], d["DepsDt"], pi, pj, pa, dt, eps0)
    v1 = s["vel"] + dt*d["DvDt"]
    E0 = (m*(0.5*(s["vel"]**2).sum(axis=1) + eps0)).sum()
    E1 = (m*(0.5*(v1**2).sum(axis=1) + eps1)).sum()
    assert abs(E1 - E0) <= 1.0e-13*abs(E0)


def test_crk_sum_density_recovers_uniform_density(oracle):
    # computeCRKSPHSumMassDensity on an (unjittered) lattice of equal masses: interior density == m/V
    ndim, n = 3, 16        # kernel radius 0.19: nodes within 0.1 of the centre see only bulk volumes
    pos, mass, H, d = ng.lattice(ndim, n, nPerh=1.51)
    W = oracle.TableKernel(0, ndim, 200)
    pi, pj, cnt = oracle.pairs(ndim, len(mass), 0, pos, H, W.kext)
    vol = oracle.crk_sum_volume(ndim, W, len(mass), 0, pos, H, pi, pj)
    rho = oracle.crk_sum_density(ndim, W, len(mass), 0, pos, mass, vol, H, pi, pj)
    inner = np.all(np.abs(pos - 0.5) < 0.1, axis=1)
    assert inner.sum() > 8 and np.allclose(rho[inner], mass[inner]/vol[inner], rtol=1e-12)


@pytest.mark.parametrize("ndim,n", [(1, 40), (2, 12), (3, 6)])
def test_reflected_rk_coefficients_equal_those_of_the_mirror_image(oracle, ndim, n):
    """ReflectingBoundary::applyGhostBoundary(Field<RKCoefficients>) (Boundary/ReflectingBoundary.cc:403-432) transforms the copied
    coefficients with RKUtilities::getTransformationMatrix(R) (RK/RKUtilities.cc:637-715).  The restatement (nodegen.reflect_map,
    "corr") against the truth: in a mirror-symmetric node set the corrections computed for the image of a node equal the transformed
    corrections of the node itself -- a plain copy does not (B and grad A change sign along the normal)."""
    nPerh = {1: 1.35, 2: 2.01, 3: 1.51}[ndim]
    pos, mass, H, d = ng.lattice(ndim, n, nPerh=nPerh)
    pos = ng.jitter(pos, 0.15, d, seed=9)
    N = pos.shape[0]
    W = oracle.TableKernel(oracle.KERNEL_BSPLINE, ndim, 200)
    nhat = np.eye(ndim)[0]
    R = np.eye(ndim) - 2.0*np.outer(nhat, nhat)
    P2 = np.concatenate([pos, pos @ R.T])                            # the node set and its mirror image through x = 0
    H2 = np.concatenate([H, ng.reflect_map(ndim, "H", H, np.zeros(N), nhat)])
    pi, pj, cnt = oracle.pairs(ndim, 2*N, 0, P2, H2, W.kext)
    vol = oracle.crk_sum_volume(ndim, W, 2*N, 0, P2, H2, pi, pj)
    corr = oracle.crk_corrections(ndim, W, 2*N, 0, P2, H2, vol, pi, pj)
    mapped = ng.reflect_map(ndim, "corr", corr[:N], np.zeros(N), nhat)
    scale = np.abs(corr).max()
    assert np.abs(mapped - corr[N:]).max() <= 1e-9*scale
    near = pos[:, 0] < 2.0*d[0]*nPerh                                # nodes whose support crosses the plane are not symmetric themselves
    assert np.abs(corr[:N][near] - corr[N:][near]).max() > 1e-3*scale
