"""ctypes front-end of the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may import
this module.  Nothing under spheral_b200/ does.  See oracle/sph_oracle.h for what is restated and how the
restatement is pinned.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsph_oracle.so")

KERNEL_BSPLINE, KERNEL_WENDLANDC4, KERNEL_WENDLANDC2 = 0, 1, 2
KERNEL_NBSPLINE = 100        # + order: NBSplineKernel(order)
Q_MG, Q_LIMITED_MG = 0, 1
H_SPH, H_ASPH, H_NONE, H_ASPH_CLASSIC = 0, 1, 2, 3


def build(force=False):
    """Compile the C restatement (gcc only)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class Options(C.Structure):
    _fields_ = [("ndim", C.c_int), ("compatibleEnergy", C.c_int), ("evolveTotalEnergy", C.c_int),
                ("XSPH", C.c_int), ("correctVelocityGradient", C.c_int),
                ("epsTensile", C.c_double), ("nTensile", C.c_double), ("nPerh", C.c_double),
                ("Qkind", C.c_int), ("Cl", C.c_double), ("Cq", C.c_double), ("eps2", C.c_double),
                ("negligibleSoundSpeed", C.c_double),
                ("balsara", C.c_int), ("linearInExpansion", C.c_int), ("quadraticInExpansion", C.c_int),
                ("etaCritFrac", C.c_double), ("etaFoldFrac", C.c_double),
                ("hEvolution", C.c_int), ("hmin", C.c_double), ("hmax", C.c_double), ("hminratio", C.c_double)]


_dp = C.POINTER(C.c_double)


class Table(C.Structure):
    _fields_ = [("kext", C.c_double), ("xmin", C.c_double), ("xstep", C.c_double), ("n1", C.c_size_t),
                ("Wcoef", _dp), ("gradWcoef", _dp), ("grad2Wcoef", _dp),
                ("nperhN", C.c_size_t), ("nperhXmin", C.c_double), ("nperhXmax", C.c_double),
                ("nperhXstep", C.c_double), ("nperhVals", _dp),
                ("wsumN", C.c_size_t), ("wsumXmin", C.c_double), ("wsumXmax", C.c_double),
                ("wsumXstep", C.c_double), ("wsumVals", _dp)]


class State(C.Structure):
    _fields_ = [(k, _dp) for k in ("pos", "vel", "H", "mass", "rho", "P", "cs", "omega", "DvDxQ", "fCl", "fCq")]


DERIV_FIELDS = ("DxDt", "DrhoDt", "DvDt", "DepsDt", "DvDx", "localDvDx", "gradRho", "M", "localM",
                "rhoSum", "normalization", "maxViscousPressure", "effViscousPressure",
                "XSPHWeightSum", "XSPHDeltaV", "DHDt", "Hideal", "massZerothMoment", "massFirstMoment",
                "pairAccelerations")


class Derivs(C.Structure):
    _fields_ = [(k, _dp) for k in DERIV_FIELDS]


class StepOptions(C.Structure):
    _fields_ = [("gamma", C.c_double), ("minimumPressure", C.c_double), ("maximumPressure", C.c_double),
                ("externalPressure", C.c_double), ("minPressureType", C.c_int), ("rhoMin", C.c_double), ("rhoMax", C.c_double),
                ("hminratio", C.c_double), ("HEvolution", C.c_int), ("cfl", C.c_double), ("useVelocityMagnitudeForDt", C.c_int)]


HEVOLUTION_IDEALH, HEVOLUTION_INTEGRATEH, HEVOLUTION_FIXEDH = 0, 1, 2
DT_REASONS = ("sound speed", "artificial viscosity", "velocity divergence", "acceleration", "velocity magnitude",
              "pairwise velocity difference")

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_kernel_extent.restype = C.c_double
        L.orc_kernel_extent.argtypes = [C.c_int, C.c_int]
        L.orc_kernel_analytic.argtypes = [C.c_int, C.c_int, C.c_double, _dp, _dp, _dp]
        L.orc_table_ncoef.restype = C.c_size_t
        L.orc_table_ncoef.argtypes = [C.c_size_t]
        L.orc_table_build.argtypes = [C.c_int, C.c_int, C.c_size_t, _dp, _dp, _dp, _dp, C.POINTER(C.c_size_t), _dp]
        L.orc_table_build_nperh.argtypes = [C.POINTER(Table), C.c_int, C.c_size_t, C.c_double, C.c_double,
                                            _dp, _dp, _dp, _dp]
        L.orc_table_eval.argtypes = [C.POINTER(Table), C.c_double, C.c_double, _dp, _dp]
        L.orc_cubic_hermite_eval.restype = C.c_double
        L.orc_cubic_hermite_eval.argtypes = [C.c_size_t, C.c_double, C.c_double, C.c_double, _dp, C.c_double]
        u32p = C.POINTER(C.c_uint32)
        for f in (L.orc_pairs_bruteforce, L.orc_pairs_cells):
            f.restype = C.c_size_t
            f.argtypes = [C.c_int, C.c_size_t, C.c_size_t, _dp, _dp, C.c_double, u32p, u32p, C.c_size_t, u32p]
        L.orc_evaluate_derivatives.argtypes = [C.POINTER(Options), C.POINTER(Table), C.POINTER(Table),
                                               C.c_size_t, C.c_size_t, C.POINTER(State), C.c_size_t, u32p, u32p,
                                               u32p, C.POINTER(Derivs), C.c_int]
        L.orc_update_energy_compatible.argtypes = [C.c_int, C.c_size_t, C.c_size_t, _dp, _dp, _dp, _dp, C.c_size_t,
                                                   u32p, u32p, _dp, C.c_double, _dp]
        L.orc_crk_sum_volume.argtypes = [C.c_int, C.POINTER(Table), C.c_size_t, C.c_size_t, _dp, _dp, C.c_size_t, u32p, u32p, _dp]
        L.orc_crk_corrections.argtypes = [C.c_int, C.POINTER(Table), C.c_size_t, C.c_size_t, _dp, _dp, _dp, C.c_size_t,
                                          u32p, u32p, _dp]
        L.orc_crk_sum_density.argtypes = [C.c_int, C.POINTER(Table), C.c_size_t, C.c_size_t, _dp, _dp, _dp, _dp, C.c_size_t,
                                          u32p, u32p, C.c_double, C.c_double, _dp]
        L.orc_crk_evaluate_derivatives.argtypes = [C.POINTER(Options), C.POINTER(Table), C.c_size_t, C.c_size_t,
                                                   C.POINTER(State), _dp, _dp, C.c_size_t, u32p, u32p, C.POINTER(Derivs)]
        L.orc_rk_kernel_grad.argtypes = [C.c_int, C.POINTER(Table), _dp, _dp, _dp, _dp, _dp]
        L.orc_rk_kernel_grad.restype = None
        L.orc_sum_mass_density.argtypes = [C.c_int, C.POINTER(Table), C.c_size_t, C.c_size_t, _dp, _dp, _dp, C.c_size_t, u32p, u32p, _dp]
        L.orc_omega_gradh.argtypes = [C.c_int, C.POINTER(Table), C.c_size_t, C.c_size_t, _dp, _dp, C.c_size_t, u32p, u32p, u32p, _dp]
        L.orc_eos_gamma_law.argtypes = [C.POINTER(StepOptions), C.c_size_t, _dp, _dp, _dp, _dp]
        L.orc_eos_gamma_law.restype = None
        L.orc_state_update.argtypes = [C.POINTER(Options), C.POINTER(StepOptions), C.c_size_t, C.c_size_t, C.c_double, C.c_int, C.c_int,
                                       C.POINTER(Derivs), _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_hydro_dt.restype = C.c_double
        L.orc_hydro_dt.argtypes = [C.POINTER(Options), C.POINTER(StepOptions), C.c_size_t, _dp, _dp, _dp, _dp, C.POINTER(Derivs),
                                   C.c_size_t, u32p, u32p, C.POINTER(C.c_int), u32p]
        L.orc_sym_bound.argtypes = [C.c_int, _dp, C.c_double, C.c_double]
        L.orc_sym_bound.restype = None
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _u(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_uint32))


def _c(a, dtype=np.float64):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def nsym(ndim):
    return {1: 1, 2: 3, 3: 6}[ndim]


class TableKernel:
    """TableKernel(kind, numPoints) restated (Kernel/TableKernel.cc:169-209)."""

    def __init__(self, kind, ndim, numPoints=100, minNperh=0.25, maxNperh=64.0, with_nperh=True):
        L = lib()
        nc = L.orc_table_ncoef(numPoints)
        self.kind, self.ndim, self.numPoints = kind, ndim, numPoints
        self.Wcoef = np.zeros(nc)
        self.gradWcoef = np.zeros(nc)
        self.grad2Wcoef = np.zeros(nc)
        kext = C.c_double()
        n1 = C.c_size_t()
        xstep = C.c_double()
        rc = L.orc_table_build(kind, ndim, numPoints, _p(self.Wcoef), _p(self.gradWcoef), _p(self.grad2Wcoef),
                               C.byref(kext), C.byref(n1), C.byref(xstep))
        assert rc == 0
        self.kext, self.n1, self.xstep, self.xmin = kext.value, n1.value, xstep.value, 0.0
        self.wsumVals = self.nperhVals = None
        self.wsumRange = self.nperhRange = None
        if with_nperh:
            self.wsumVals = np.zeros(2*numPoints)
            self.nperhVals = np.zeros(2*numPoints)
            self.wsumRange = np.zeros(2)
            self.nperhRange = np.zeros(2)
            t = self.ctable()
            L.orc_table_build_nperh(C.byref(t), ndim, numPoints, minNperh, maxNperh,
                                    _p(self.wsumVals), _p(self.wsumRange), _p(self.nperhVals), _p(self.nperhRange))

    @classmethod
    def from_arrays(cls, ndim, kext, xmin, xstep, n1, Wcoef, gradWcoef, grad2Wcoef=None,
                    nperh=None, wsum=None):
        """Wrap an externally built table (e.g. the product's) so GPU and oracle share ONE table."""
        self = cls.__new__(cls)
        self.kind, self.ndim = -1, ndim
        self.kext, self.xmin, self.xstep, self.n1 = kext, xmin, xstep, n1
        self.Wcoef, self.gradWcoef = _c(Wcoef), _c(gradWcoef)
        self.grad2Wcoef = _c(grad2Wcoef) if grad2Wcoef is not None else np.zeros_like(self.Wcoef)
        self.numPoints = 0
        self.wsumVals = self.nperhVals = self.wsumRange = self.nperhRange = None
        if nperh is not None:       # (vals[2n], xmin, xmax)
            self.nperhVals, self.nperhRange = _c(nperh[0]), np.array(nperh[1:3], dtype=float)
            self.numPoints = len(self.nperhVals)//2
        if wsum is not None:
            self.wsumVals, self.wsumRange = _c(wsum[0]), np.array(wsum[1:3], dtype=float)
        return self

    def ctable(self):
        t = Table()
        t.kext, t.xmin, t.xstep, t.n1 = self.kext, self.xmin, self.xstep, self.n1
        t.Wcoef, t.gradWcoef, t.grad2Wcoef = _p(self.Wcoef), _p(self.gradWcoef), _p(self.grad2Wcoef)
        if self.nperhVals is not None and self.nperhRange is not None and self.nperhRange[1] != self.nperhRange[0]:
            n = len(self.nperhVals)//2
            t.nperhN, t.nperhXmin, t.nperhXmax = n, self.nperhRange[0], self.nperhRange[1]
            t.nperhXstep = (self.nperhRange[1] - self.nperhRange[0])/(n - 1)
            t.nperhVals = _p(self.nperhVals)
        if self.wsumVals is not None and self.wsumRange is not None and self.wsumRange[1] != self.wsumRange[0]:
            n = len(self.wsumVals)//2
            t.wsumN, t.wsumXmin, t.wsumXmax = n, self.wsumRange[0], self.wsumRange[1]
            t.wsumXstep = (self.wsumRange[1] - self.wsumRange[0])/(n - 1)
            t.wsumVals = _p(self.wsumVals)
        return t

    def kernelAndGradValue(self, eta, Hdet=1.0):
        W, g = C.c_double(), C.c_double()
        t = self.ctable()
        lib().orc_table_eval(C.byref(t), eta, Hdet, C.byref(W), C.byref(g))
        return W.value, g.value

    def equivalentNodesPerSmoothingScale(self, Wsum):
        t = self.ctable()
        return max(0.0, lib().orc_cubic_hermite_eval(t.nperhN, t.nperhXmin, t.nperhXmax, t.nperhXstep,
                                                     t.nperhVals, Wsum))

    def equivalentWsum(self, nPerh):
        t = self.ctable()
        return max(0.0, lib().orc_cubic_hermite_eval(t.wsumN, t.wsumXmin, t.wsumXmax, t.wsumXstep,
                                                     t.wsumVals, nPerh))


def kernel_analytic(kind, ndim, eta):
    W, g, g2 = C.c_double(), C.c_double(), C.c_double()
    lib().orc_kernel_analytic(kind, ndim, eta, C.byref(W), C.byref(g), C.byref(g2))
    return W.value, g.value, g2.value


def pairs(ndim, nInt, nGhost, pos, H, kext, method="cells"):
    """Sorted (i<j) pair list + numNeighborsForNode for internal nodes."""
    L = lib()
    pos, H = _c(pos), _c(H)
    f = L.orc_pairs_bruteforce if method == "brute" else L.orc_pairs_cells
    counts = np.zeros(nInt, dtype=np.uint32)
    dummy = np.zeros(1, dtype=np.uint32)
    npairs = f(ndim, nInt, nGhost, _p(pos), _p(H), kext, _u(dummy), _u(dummy), 0, _u(counts))
    pi = np.zeros(max(npairs, 1), dtype=np.uint32)
    pj = np.zeros(max(npairs, 1), dtype=np.uint32)
    n2 = f(ndim, nInt, nGhost, _p(pos), _p(H), kext, _u(pi), _u(pj), npairs, _u(counts))
    assert n2 == npairs
    return pi[:npairs], pj[:npairs], counts


def default_options(ndim, **kw):
    o = Options()
    o.ndim = ndim
    o.compatibleEnergy, o.evolveTotalEnergy, o.XSPH, o.correctVelocityGradient = 1, 0, 1, 1
    o.epsTensile, o.nTensile, o.nPerh = 0.0, 4.0, 2.01
    o.Qkind, o.Cl, o.Cq, o.eps2, o.negligibleSoundSpeed = Q_MG, 1.0, 1.0, 1.0e-2, 1.0e-10
    o.balsara = o.linearInExpansion = o.quadraticInExpansion = 0
    o.etaCritFrac, o.etaFoldFrac = 1.0, 0.2
    o.hEvolution, o.hmin, o.hmax, o.hminratio = H_SPH, 1.0e-20, 1.0e20, 0.1
    for k, v in kw.items():
        if not hasattr(o, k):
            raise KeyError(k)
        setattr(o, k, v)
    return o


def evaluate_derivatives(opts, W, state, nInt, nGhost, pi, pj, counts, WQ=None, nthreads=1):
    """state: dict of AoS numpy arrays (pos, vel, H, mass, rho, P, cs, omega[, DvDxQ, fCl, fCq]).
    Returns dict of derivative arrays (DERIV_FIELDS)."""
    L = lib()
    nd = opts.ndim
    n = nInt + nGhost
    ns, nt = nsym(nd), nd*nd
    keep = {k: _c(state.get(k)) for k in ("pos", "vel", "H", "mass", "rho", "P", "cs", "omega", "DvDxQ", "fCl", "fCq")}
    s = State(**{k: _p(v) for k, v in keep.items()})
    npairs = len(pi)
    shapes = dict(DxDt=nd, DrhoDt=1, DvDt=nd, DepsDt=1, DvDx=nt, localDvDx=nt, gradRho=nd, M=nt, localM=nt,
                  rhoSum=1, normalization=1, maxViscousPressure=1, effViscousPressure=1, XSPHWeightSum=1,
                  XSPHDeltaV=nd, DHDt=ns, Hideal=ns, massZerothMoment=1, massFirstMoment=nd)
    scalars = ("DrhoDt", "DepsDt", "rhoSum", "normalization", "maxViscousPressure", "effViscousPressure", "XSPHWeightSum",
               "massZerothMoment")
    # vectors and tensors keep a component axis in 1-D too (their ghosts are reflected, scalars are copied)
    out = {k: np.zeros(n if k in scalars else (n, w)) for k, w in shapes.items()}
    out["pairAccelerations"] = np.zeros((npairs, nd))
    d = Derivs(**{k: _p(v) for k, v in out.items()})
    pi, pj, counts = _c(pi, np.uint32), _c(pj, np.uint32), _c(counts, np.uint32)
    tW = W.ctable()
    tQ = WQ.ctable() if WQ is not None else None
    rc = L.orc_evaluate_derivatives(C.byref(opts), C.byref(tW), C.byref(tQ) if tQ is not None else None,
                                    nInt, nGhost, C.byref(s), npairs, _u(pi), _u(pj), _u(counts), C.byref(d),
                                    nthreads)
    if rc != 0:
        raise RuntimeError("SPH error : you cannot simultaneously use both compatibleEnergyEvolution and "
                           "evolveTotalEnergy" if rc == 2 else "oracle error %d" % rc)
    return out


def update_energy_compatible(ndim, nInt, nGhost, mass, vel, DvDt, DepsDt0, pi, pj, pacc, multiplier, eps):
    eps = np.array(eps, dtype=np.float64, copy=True)
    mass, vel, DvDt, DepsDt0, pacc = map(_c, (mass, vel, DvDt, DepsDt0, pacc))
    pi, pj = _c(pi, np.uint32), _c(pj, np.uint32)
    lib().orc_update_energy_compatible(ndim, nInt, nGhost, _p(mass), _p(vel), _p(DvDt), _p(DepsDt0), len(pi),
                                       _u(pi), _u(pj), _p(pacc), multiplier, _p(eps))
    return eps


# ---- CRKSPH (LinearOrder reproducing kernels, RKSumVolume) -------------------------------------------------
def crk_ncorr(ndim):
    return (1 + ndim)*(1 + ndim)


def crk_sum_volume(ndim, W, nInt, nGhost, pos, H, pi, pj, vol=None):
    """computeRKSumVolume: internal entries are written, ghost entries keep the caller's values."""
    pos, H = _c(pos), _c(H)
    vol = np.zeros(nInt + nGhost) if vol is None else np.array(vol, dtype=np.float64, copy=True)
    pi, pj = _c(pi, np.uint32), _c(pj, np.uint32)
    t = W.ctable()
    lib().orc_crk_sum_volume(ndim, C.byref(t), nInt, nGhost, _p(pos), _p(H), len(pi), _u(pi), _u(pj), _p(vol))
    return vol


def crk_corrections(ndim, W, nInt, nGhost, pos, H, vol, pi, pj, corr=None):
    """RKUtilities::computeCorrections (LinearOrder): (n, (1+ndim)^2); internal rows written."""
    pos, H, vol = _c(pos), _c(H), _c(vol)
    nc = crk_ncorr(ndim)
    corr = np.zeros((nInt + nGhost, nc)) if corr is None else np.array(corr, dtype=np.float64, copy=True).reshape(-1, nc)
    pi, pj = _c(pi, np.uint32), _c(pj, np.uint32)
    t = W.ctable()
    lib().orc_crk_corrections(ndim, C.byref(t), nInt, nGhost, _p(pos), _p(H), _p(vol), len(pi), _u(pi), _u(pj), _p(corr))
    return corr


def crk_sum_density(ndim, W, nInt, nGhost, pos, mass, vol, H, pi, pj, rhoMin=0.0, rhoMax=1.0e300, rho=None):
    pos, mass, vol, H = _c(pos), _c(mass), _c(vol), _c(H)
    rho = np.zeros(nInt + nGhost) if rho is None else np.array(rho, dtype=np.float64, copy=True)
    pi, pj = _c(pi, np.uint32), _c(pj, np.uint32)
    t = W.ctable()
    lib().orc_crk_sum_density(ndim, C.byref(t), nInt, nGhost, _p(pos), _p(mass), _p(vol), _p(H), len(pi), _u(pi), _u(pj),
                              rhoMin, rhoMax, _p(rho))
    return rho


def rk_kernel_grad(ndim, W, x, H, corr):
    x, H, corr = _c(x), _c(H), _c(corr)
    WR = C.c_double()
    g = np.zeros(ndim)
    t = W.ctable()
    lib().orc_rk_kernel_grad(ndim, C.byref(t), _p(x), _p(H), _p(corr), C.byref(WR), _p(g))
    return WR.value, g


def crk_evaluate_derivatives(opts, W, state, vol, corr, nInt, nGhost, pi, pj):
    """CRKSPH::evaluateDerivatives + smoothing-scale sub-package.  Same output dict as evaluate_derivatives."""
    L = lib()
    nd = opts.ndim
    n = nInt + nGhost
    ns, nt = nsym(nd), nd*nd
    keep = {k: _c(state.get(k)) for k in ("pos", "vel", "H", "mass", "rho", "P", "cs", "omega", "DvDxQ", "fCl", "fCq")}
    s = State(**{k: _p(v) for k, v in keep.items()})
    npairs = len(pi)
    shapes = dict(DxDt=nd, DrhoDt=1, DvDt=nd, DepsDt=1, DvDx=nt, localDvDx=nt, gradRho=nd, M=nt, localM=nt,
                  rhoSum=1, normalization=1, maxViscousPressure=1, effViscousPressure=1, XSPHWeightSum=1,
                  XSPHDeltaV=nd, DHDt=ns, Hideal=ns, massZerothMoment=1, massFirstMoment=nd)
    scalars = ("DrhoDt", "DepsDt", "rhoSum", "normalization", "maxViscousPressure", "effViscousPressure", "XSPHWeightSum",
               "massZerothMoment")
    # vectors and tensors keep a component axis in 1-D too (their ghosts are reflected, scalars are copied)
    out = {k: np.zeros(n if k in scalars else (n, w)) for k, w in shapes.items()}
    out["pairAccelerations"] = np.zeros((npairs, nd))
    d = Derivs(**{k: _p(v) for k, v in out.items()})
    pi, pj = _c(pi, np.uint32), _c(pj, np.uint32)
    vol, corr = _c(vol), _c(corr)
    tW = W.ctable()
    rc = L.orc_crk_evaluate_derivatives(C.byref(opts), C.byref(tW), nInt, nGhost, C.byref(s), _p(vol), _p(corr),
                                        npairs, _u(pi), _u(pj), C.byref(d))
    if rc != 0:
        raise RuntimeError("oracle error %d" % rc)
    return out


# ---- per-step callers of the derivative path ------------------------------------------------
def default_step_options(**kw):
    """GammaLawGas(gamma=5/3) with no pressure limits, rho in [1e-10, 1e10] (the FluidNodeList defaults),
    hminratio 0.1 (NodeList default), IdealH, cfl 0.25 (GenericHydro default)."""
    so = StepOptions()
    so.gamma = 5.0/3.0
    so.minimumPressure, so.maximumPressure, so.externalPressure, so.minPressureType = -1.0e200, 1.0e200, 0.0, 0
    so.rhoMin, so.rhoMax, so.hminratio = 1.0e-10, 1.0e10, 0.1
    so.HEvolution, so.cfl, so.useVelocityMagnitudeForDt = HEVOLUTION_IDEALH, 0.25, 0
    for k, v in kw.items():
        if not hasattr(so, k):
            raise KeyError(k)
        setattr(so, k, v)
    return so


def sum_mass_density(ndim, W, nInt, nGhost, pos, mass, H, pi, pj, rho=None):
    """computeSPHSumMassDensity: internal entries written, ghost entries keep the caller's values."""
    pos, mass, H = _c(pos), _c(mass), _c(H)
    rho = np.zeros(nInt + nGhost) if rho is None else np.array(rho, dtype=np.float64, copy=True)
    pi, pj = _c(pi, np.uint32), _c(pj, np.uint32)
    t = W.ctable()
    rc = lib().orc_sum_mass_density(ndim, C.byref(t), nInt, nGhost, _p(pos), _p(mass), _p(H), len(pi), _u(pi), _u(pj), _p(rho))
    assert rc == 0
    return rho


def omega_gradh(ndim, W, nInt, nGhost, pos, H, pi, pj, counts, omega=None):
    """computeSPHOmegaGradhCorrection: internal entries written."""
    pos, H = _c(pos), _c(H)
    omega = np.ones(nInt + nGhost) if omega is None else np.array(omega, dtype=np.float64, copy=True)
    pi, pj, counts = _c(pi, np.uint32), _c(pj, np.uint32), _c(counts, np.uint32)
    t = W.ctable()
    rc = lib().orc_omega_gradh(ndim, C.byref(t), nInt, nGhost, _p(pos), _p(H), len(pi), _u(pi), _u(pj), _u(counts), _p(omega))
    assert rc == 0
    return omega


def eos_gamma_law(so, rho, eps):
    rho, eps = _c(rho), _c(eps)
    P, cs = np.zeros_like(rho), np.zeros_like(rho)
    lib().orc_eos_gamma_law(C.byref(so), len(rho), _p(rho), _p(eps), _p(P), _p(cs))
    return P, cs


def _derivs_struct(derivs):
    keep = {k: _c(derivs[k]) for k in DERIV_FIELDS if k in derivs and derivs[k] is not None}
    return Derivs(**{k: _p(v) for k, v in keep.items()}), keep


def state_update(opts, so, nInt, nGhost, multiplier, timeAdvanceOnly, derivs, state, epsDone=False):
    """State::update over the internal nodes.  state: dict pos, vel, H, rho, eps, P, cs (copied, returned updated)."""
    out = {k: np.array(state[k], dtype=np.float64, copy=True) for k in ("pos", "vel", "H", "rho", "eps", "P", "cs")}
    d, keep = _derivs_struct(derivs)
    rc = lib().orc_state_update(C.byref(opts), C.byref(so), nInt, nGhost, multiplier, int(timeAdvanceOnly), int(epsDone), C.byref(d),
                                _p(out["pos"]), _p(out["vel"]), _p(out["H"]), _p(out["rho"]), _p(out["eps"]), _p(out["P"]), _p(out["cs"]))
    assert rc == 0
    return out


def hydro_dt(opts, so, nInt, vel, H, rho, cs, derivs, pi, pj):
    """GenericHydro::dt -> (dt, reason string, node)."""
    vel, H, rho, cs = _c(vel), _c(H), _c(rho), _c(cs)
    pi, pj = _c(pi, np.uint32), _c(pj, np.uint32)
    d, keep = _derivs_struct(derivs)
    reason, node = C.c_int(), C.c_uint32()
    dt = lib().orc_hydro_dt(C.byref(opts), C.byref(so), nInt, _p(vel), _p(H), _p(rho), _p(cs), C.byref(d), len(pi), _u(pi), _u(pj),
                            C.byref(reason), C.byref(node))
    return dt, DT_REASONS[reason.value], node.value


def sym_bound(ndim, H, minv, maxv):
    H = np.array(H, dtype=np.float64, copy=True)
    lib().orc_sym_bound(ndim, _p(H), minv, maxv)
    return H
