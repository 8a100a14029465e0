"""Synthetic node sets for the parity tests and bench.py (host-side numpy; not on the hot path).

The formulas restate the reference's Python node generators so that the inputs have the same shape as the
BASELINE.json configs:
  * lattice        : src/NodeGenerators/GenerateNodeDistribution3d.py:573-624 (and the 2-D analogue)
  * constantDTheta : src/NodeGenerators/GenerateNodeDistribution2d.py:376-424   (Noh-cylindrical-2d seed)
  * jitter         : tests/unit/SPH/testLinearVelocityGradient.py:224-244       (ranfrac*dx*U(-1,1), seed 14892042)
  * random rotated anisotropic H : tests/unit/Neighbor/NeighborTestBase.py:10-82
  * reflecting-plane ghosts      : src/Boundary/findNodesTouchingThroughPlanes.cc, PlanarBoundary.cc:102-117,
                                   ReflectingBoundary.cc:182-250, Utilities/planarReflectingOperator.hh:14-19
  * gamma-law EOS  : src/Material/GammaLawGas.cc:185-243
Layouts are the reference's AoS: Vector ndim doubles, SymTensor (xx,xy,xz,yy,yz,zz | xx,xy,yy), Tensor row major.
"""
import math
import random as _pyrandom
import numpy as np


def nsym(ndim):
    return {1: 1, 2: 3, 3: 6}[ndim]


def sym_from_diag(ndim, diag):
    """SymTensor AoS rows from per-node diagonal entries (n, ndim)."""
    diag = np.atleast_2d(diag)
    H = np.zeros((diag.shape[0], nsym(ndim)))
    if ndim == 3:
        H[:, 0], H[:, 3], H[:, 5] = diag[:, 0], diag[:, 1], diag[:, 2]
    elif ndim == 2:
        H[:, 0], H[:, 2] = diag[:, 0], diag[:, 1]
    else:
        H[:, 0] = diag[:, 0]
    return H


def sym_to_full(ndim, H):
    """(n, nsym) -> (n, ndim, ndim)."""
    H = np.atleast_2d(H)
    F = np.zeros((H.shape[0], ndim, ndim))
    if ndim == 3:
        idx = [(0, 0, 0), (0, 1, 1), (0, 2, 2), (1, 1, 3), (1, 2, 4), (2, 2, 5)]
    elif ndim == 2:
        idx = [(0, 0, 0), (0, 1, 1), (1, 1, 2)]
    else:
        idx = [(0, 0, 0)]
    for r, c, k in idx:
        F[:, r, c] = H[:, k]
        F[:, c, r] = H[:, k]
    return F


def full_to_sym(ndim, F):
    if ndim == 3:
        return np.stack([F[:, 0, 0], F[:, 0, 1], F[:, 0, 2], F[:, 1, 1], F[:, 1, 2], F[:, 2, 2]], axis=1)
    if ndim == 1:
        return F[:, 0, 0].reshape(-1, 1)
    return np.stack([F[:, 0, 0], F[:, 0, 1], F[:, 1, 1]], axis=1)


def lattice(ndim, n, xmin=None, xmax=None, rho0=1.0, nPerh=2.01):
    """Lattice in [xmin,xmax]; n = int or per-axis tuple.  Node order x fastest (iglobal % nx ...)."""
    nn = (n,)*ndim if np.isscalar(n) else tuple(n)
    xmin = np.zeros(ndim) if xmin is None else np.asarray(xmin, dtype=float)
    xmax = np.ones(ndim) if xmax is None else np.asarray(xmax, dtype=float)
    d = (xmax - xmin)/np.asarray(nn)
    axes = [xmin[a] + (np.arange(nn[a]) + 0.5)*d[a] for a in range(ndim)]
    if ndim == 3:
        Z, Y, X = np.meshgrid(axes[2], axes[1], axes[0], indexing="ij")
        pos = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    elif ndim == 1:                                              # 1-D lattices serve the CPU test suite only; the engine is 2-D / 3-D
        pos = axes[0].reshape(-1, 1)
    else:
        Y, X = np.meshgrid(axes[1], axes[0], indexing="ij")
        pos = np.stack([X.ravel(), Y.ravel()], axis=1)
    N = pos.shape[0]
    mass = np.full(N, float(np.prod(d))*rho0)
    H = sym_from_diag(ndim, np.tile(1.0/(nPerh*d), (N, 1)))
    return np.ascontiguousarray(pos), mass, H, d


def jitter_python_random(pos, frac, d, seed=14892042):
    """Exactly the reference unit test's jitter: python `random`, one uniform per axis per node, in node order."""
    rng = _pyrandom.Random(seed)
    pos = pos.copy()
    for i in range(pos.shape[0]):
        for a in range(pos.shape[1]):
            pos[i, a] += frac*d[a]*rng.uniform(-1.0, 1.0)
    return pos


def jitter(pos, frac, d, seed=14892042):
    rng = np.random.default_rng(seed)
    return pos + frac*np.asarray(d)*rng.uniform(-1.0, 1.0, size=pos.shape)


def random_rotation(ndim, rng):
    if ndim == 2:
        t = rng.uniform(0.0, math.pi)
        return np.array([[math.cos(t), -math.sin(t)], [math.sin(t), math.cos(t)]])
    t1, t2, t3 = rng.uniform(0.0, math.pi, 3)
    R1 = np.array([[math.cos(t1), -math.sin(t1), 0], [math.sin(t1), math.cos(t1), 0], [0, 0, 1.0]])
    R2 = np.array([[math.cos(t2), 0, -math.sin(t2)], [0, 1.0, 0], [math.sin(t2), 0, math.cos(t2)]])
    R3 = np.array([[1.0, 0, 0], [0, math.cos(t3), -math.sin(t3)], [0, math.sin(t3), math.cos(t3)]])
    return R1 @ R2 @ R3


def random_anisotropic(ndim, n, box, nPerh=2.01, seed=4599281940):
    """NeighborTestBase.randomDistribute: uniform positions, H = R diag(1/(nPerh*U(0.5,2)*dx0)) R^T."""
    rng = np.random.default_rng(seed)
    box = np.asarray(box, dtype=float)            # (ndim, 2)
    vol = float(np.prod(box[:, 1] - box[:, 0]))
    dx0 = (vol/n)**(1.0/ndim)
    pos = rng.uniform(box[:, 0], box[:, 1], size=(n, ndim))
    F = np.zeros((n, ndim, ndim))
    for i in range(n):
        dx = nPerh*rng.uniform(0.5, 2.0, ndim)*dx0
        R = random_rotation(ndim, rng)
        F[i] = R @ np.diag(1.0/dx) @ R.T
    return np.ascontiguousarray(pos), np.ascontiguousarray(full_to_sym(ndim, F))


def constant_dtheta_2d(nRadial, rho0=1.0, rmin=0.0, rmax=1.0, nPerh=2.01, theta=math.pi/2.0):
    dr = (rmax - rmin)/nRadial
    xs, ys, ms = [], [], []
    for i in range(nRadial):
        rInner, rOuter, ri = rmin + i*dr, rmin + (i + 1)*dr, rmin + (i + 0.5)*dr
        nTheta = max(1, int(theta*ri/dr))
        dTheta = theta/nTheta
        mi = (rOuter**2 - rInner**2)*theta/2.0*rho0/nTheta
        for j in range(nTheta):
            t = (j + 0.5)*dTheta
            xs.append(ri*math.cos(t)); ys.append(ri*math.sin(t)); ms.append(mi)
    pos = np.stack([np.array(xs), np.array(ys)], axis=1)
    h = 1.0/(nPerh*dr)
    H = sym_from_diag(2, np.full((len(xs), 2), h))
    return np.ascontiguousarray(pos), np.array(ms), H


def gamma_law(rho, eps, gamma=5.0/3.0):
    """P = (gamma-1) rho eps ; cs = sqrt(gamma (gamma-1) eps)   (GammaLawGas.cc:185-243)."""
    P = (gamma - 1.0)*rho*eps
    cs = np.sqrt(np.maximum(0.0, gamma*(gamma - 1.0)*eps))
    return P, cs


def _unit(v):
    v = np.asarray(v, dtype=float)
    return v/np.linalg.norm(v)


def expand_boundaries(boundaries):
    """("reflecting", (p, n)) / ("periodic", (p1, n1), (p2, n2)) -> planar boundaries (periodic, enter (p, n), exit (p, n));
    a PeriodicBoundary is two of them (Boundary/PeriodicBoundary.cc:60-62)."""
    out = []
    for b in boundaries:
        if b[0] == "reflecting":
            e = (np.asarray(b[1][0], dtype=float), _unit(b[1][1]))
            out.append((False, e, e))
        elif b[0] == "periodic":
            e1 = (np.asarray(b[1][0], dtype=float), _unit(b[1][1]))
            e2 = (np.asarray(b[2][0], dtype=float), _unit(b[2][1]))
            out += [(True, e1, e2), (True, e2, e1)]
        else:
            raise ValueError(b[0])
    return out


def reflect_map(ndim, name, c, sd, nhat, periodic=False, sd_exit=None):
    """Ghost values of one field from its control values c (ReflectingBoundary::applyGhostBoundary,
    Boundary/ReflectingBoundary.cc:182-250; periodic boundaries copy).  Positions: mapPosition(r, exit, enter) =
    closestPointOnPlane_enter(r) - signedDistance_exit(r) n_enter (PlanarBoundary.cc:318-320); sd = distance to the enter plane."""
    if name == "pos":
        sx = sd if sd_exit is None else sd_exit
        return (c - np.outer(sd, nhat)) - np.outer(sx, nhat)
    if periodic:
        return c.copy()
    R = np.eye(ndim) - 2.0*np.outer(nhat, nhat)
    if name == "H":
        Fr = np.einsum("ab,nbc,cd->nad", R, sym_to_full(ndim, c), R)
        return full_to_sym(ndim, 0.5*(Fr + np.transpose(Fr, (0, 2, 1))))
    if ndim == 1 and name.startswith("DvDx"):                   # a 1-D tensor has the width of a vector: R.(T.R) = T
        return c.copy()
    if name in ("corr", "rkCorrections"):
        # RK coefficients (LinearOrder): ReflectingBoundary::applyGhostBoundary(Field<RKCoefficients>) (Boundary/ReflectingBoundary.cc:403-432)
        # applies RKUtilities::getTransformationMatrix(R) (RK/RKUtilities.cc:637-715): with the layout {A, B_k | dA/dx_d, dB_k/dx_d}
        # A' = A, B' = R.B, (grad A)' = R.grad A, (grad B)' = R.(grad B).R
        n = c.shape[0]
        C = np.asarray(c, dtype=float).reshape(n, 1 + ndim, 1 + ndim)
        out = C.copy()
        out[:, 0, 1:] = C[:, 0, 1:] @ R.T
        out[:, 1:, 0] = C[:, 1:, 0] @ R.T
        out[:, 1:, 1:] = np.einsum("ab,nbc,cd->nad", R, C[:, 1:, 1:], R)
        return out.reshape(c.shape)
    if c.ndim == 2 and c.shape[1] == ndim:                      # vectors
        return c @ R.T
    if c.ndim == 2 and c.shape[1] == ndim*ndim:                 # tensors R.(T.R)
        return np.einsum("ab,nbc,cd->nad", R, c.reshape(-1, ndim, ndim), R).reshape(-1, ndim*ndim)
    return c.copy()                                             # scalars (and any other width): copy


def boundary_apply(ndim, fields, boundaries, ctl_per_plane, n0):
    """Refresh the ghost entries of `fields` (arrays of n0 internal + ghosts, ghosts laid out boundary after boundary) from
    their control nodes, in order, so that later boundaries see the refreshed ghosts of earlier ones."""
    first = n0
    for (periodic, (pe, ne), (px, nx)), ctl in zip(expand_boundaries(boundaries), ctl_per_plane):
        sd = (fields["pos"][ctl] - pe) @ ne
        sx = (fields["pos"][ctl] - px) @ nx
        for k, v in fields.items():
            v[first:first + len(ctl)] = reflect_map(ndim, k, v[ctl], sd, ne, periodic, sx)
        first += len(ctl)
    return fields


def boundary_ghosts(ndim, fields, boundaries, kext):
    """Ghost nodes of planar boundaries applied one after the other (Integrator.cc:415-424): control nodes by
    findNodesTouchingThroughPlanes (active branch): hmax = largest 1/lambda_min(H_i) among nodes closer than kext*hmax_i to either
    plane; controls are the nodes with 0 <= signedDistance_exit/hmax <= kext.  Returns (fields with ghosts, control lists, n0)."""
    out = {k: np.array(v, dtype=float, copy=True) for k, v in fields.items()}
    n0 = out["pos"].shape[0]
    lists = []
    for periodic, (pe, ne), (px, nx) in expand_boundaries(boundaries):
        pos = out["pos"]
        hmax_i = 1.0/np.linalg.eigvalsh(sym_to_full(ndim, out["H"]))[:, 0]
        sde, sdx = (pos - pe) @ ne, (pos - px) @ nx
        near = np.minimum(np.abs(sde), np.abs(sdx)) < kext*hmax_i
        if not near.any():
            lists.append(np.zeros(0, dtype=np.int64))
            continue
        hmax = hmax_i[near].max()
        ctl = np.nonzero((sdx/hmax >= 0.0) & (sdx/hmax <= kext))[0]
        new = {k: reflect_map(ndim, k, v[ctl], sde[ctl], ne, periodic, sdx[ctl]) for k, v in out.items()}
        for k in out:
            out[k] = np.concatenate([out[k], new[k]], axis=0)
        lists.append(ctl)
    return out, lists, n0


def reflect_apply(ndim, fields, planes, ctl_per_plane, n0):
    """boundary_apply for reflecting planes given as (point, inward normal)."""
    return boundary_apply(ndim, fields, [("reflecting", pl) for pl in planes], ctl_per_plane, n0)


def reflect_ghosts(ndim, fields, planes, kext, per_plane=False):
    """Reflecting-boundary ghosts for planes given as (point, inward normal); see boundary_ghosts.
    Returns (fields_with_ghosts, control index array -- or the per-plane lists --, n0)."""
    out, lists, n0 = boundary_ghosts(ndim, fields, [("reflecting", pl) for pl in planes], kext)
    if per_plane:
        return out, lists, n0
    return out, np.concatenate(lists).astype(np.int64) if lists else np.zeros(0, dtype=np.int64), n0
