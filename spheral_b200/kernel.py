"""Host-side kernels: the Python face of src/Kernel for the hot path.

    WT = TableKernel(BSplineKernel(3), 1000)          # reference: TableKernel3d(BSplineKernel3d(), 1000)

The table payload is built by the C++ host code of libsphb200 (csrc/tablekernel_host.cpp, a restatement of
Kernel/TableKernel.cc:169-209); no GPU is needed for that.  Property names follow PYB11/Kernel/Kernel.py.
"""
import ctypes as C
import numpy as np

from . import _lib as L


class _AnalyticKernel:
    kind = None

    def __init__(self, ndim=3):
        self.ndim = ndim

    @property
    def kernelExtent(self):
        if self.kind >= L.KERNEL_NBSPLINE:
            return float((self.kind - L.KERNEL_NBSPLINE + 1)//2)      # NBSplineKernel.cc:116, integer arithmetic
        return 2.0 if self.kind == L.KERNEL_BSPLINE else 1.0


class BSplineKernel(_AnalyticKernel):
    kind = L.KERNEL_BSPLINE


class WendlandC4Kernel(_AnalyticKernel):
    kind = L.KERNEL_WENDLANDC4


class WendlandC2Kernel(_AnalyticKernel):
    kind = L.KERNEL_WENDLANDC2


class NBSplineKernel(_AnalyticKernel):
    """NBSplineKernel(order) -- PYB11/Kernel/Kernel.py; Kernel/NBSplineKernel.cc.  The stock Noh scripts build their TableKernel from
    NBSplineKernel(5) (tests/functional/Hydro/Noh/Noh-spherical-3d.py:25,205)."""

    def __init__(self, ndim=3, order=5):
        super().__init__(ndim)
        if not 1 <= int(order) <= 11:
            raise ValueError("NBSplineKernel: order must be in 1..11")
        self.order = int(order)
        self.kind = L.KERNEL_NBSPLINE + self.order


def BSplineKernel2d():
    return BSplineKernel(2)


def BSplineKernel3d():
    return BSplineKernel(3)


def NBSplineKernel1d(order=5):
    return NBSplineKernel(1, order)


def NBSplineKernel2d(order=5):
    return NBSplineKernel(2, order)


def NBSplineKernel3d(order=5):
    return NBSplineKernel(3, order)


def WendlandC4Kernel2d():
    return WendlandC4Kernel(2)


def WendlandC4Kernel3d():
    return WendlandC4Kernel(3)


class TableKernel:
    """TableKernel(kernel, numPoints=100, minNperh=0.25, maxNperh=64.0) -- Kernel/TableKernel.hh:34-37."""

    def __init__(self, kernel, numPoints=100, minNperh=0.25, maxNperh=64.0):
        lib = L.lib()
        self.baseKernel = kernel
        self.ndim = kernel.ndim
        self.numPoints = numPoints
        nc = lib.sphb200_table_ncoef(numPoints)
        self.Wcoef = np.zeros(nc)
        self.gradWcoef = np.zeros(nc)
        self.grad2Wcoef = np.zeros(nc)
        self.nperhVals = np.zeros(2*numPoints)
        self.wsumVals = np.zeros(2*numPoints)
        self.nperhRange = np.zeros(2)
        self.wsumRange = np.zeros(2)
        kext, xstep, n1 = C.c_double(), C.c_double(), C.c_size_t()
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        rc = lib.sphb200_table_kernel_build(kernel.kind, kernel.ndim, numPoints, minNperh, maxNperh,
                                            C.byref(kext), C.byref(xstep), C.byref(n1),
                                            dp(self.Wcoef), dp(self.gradWcoef), dp(self.grad2Wcoef),
                                            dp(self.nperhVals), dp(self.nperhRange), dp(self.wsumVals), dp(self.wsumRange))
        if rc != 0:
            raise ValueError("TableKernel ERROR: bad arguments")
        self.kernelExtent, self.xstep, self.n1, self.xmin = kext.value, xstep.value, n1.value, 0.0
        self.minNperhLookup, self.maxNperhLookup = self.wsumRange

    # --- QuadraticInterpolatorView::lowerBound / operator() (Utilities/QuadraticInterpolatorViewInline.hh) ---
    def _lookup(self, coef, eta):
        k = min(self.n1, int(max(0.0, eta - self.xmin)/self.xstep))
        return coef[3*k] + (coef[3*k + 1] + coef[3*k + 2]*eta)*eta

    def kernelValue(self, etaij, Hdet=1.0):
        return Hdet*self._lookup(self.Wcoef, etaij) if etaij < self.kernelExtent else 0.0

    def gradValue(self, etaij, Hdet=1.0):
        return Hdet*self._lookup(self.gradWcoef, etaij) if etaij < self.kernelExtent else 0.0

    def grad2Value(self, etaij, Hdet=1.0):
        return Hdet*self._lookup(self.grad2Wcoef, etaij) if etaij < self.kernelExtent else 0.0

    def kernelAndGradValue(self, etaij, Hdet=1.0):
        return self.kernelValue(etaij, Hdet), self.gradValue(etaij, Hdet)

    def __call__(self, etaij, Hdet=1.0):
        return self.kernelValue(etaij, Hdet)

    def _hermite(self, vals, rng, x):
        n = len(vals)//2
        xmin, xmax = rng
        xstep = (xmax - xmin)/(n - 1)
        if x < xmin:
            return vals[0] + vals[n]*(x - xmin)
        if x > xmax:
            return vals[n - 1] + vals[2*n - 1]*(x - xmin)
        i0 = min(n - 2, int(max(0.0, x - xmin)/xstep))
        t = max(0.0, min(1.0, (x - xmin - i0*xstep)/xstep))
        t2, t3 = t*t, t*t*t
        return ((2.0*t3 - 3.0*t2 + 1.0)*vals[i0] + (-2.0*t3 + 3.0*t2)*vals[i0 + 1] +
                xstep*((t3 - 2.0*t2 + t)*vals[n + i0] + (t3 - t2)*vals[n + i0 + 1]))

    def equivalentNodesPerSmoothingScale(self, Wsum):
        return max(0.0, self._hermite(self.nperhVals, self.nperhRange, Wsum))

    def equivalentWsum(self, nPerh):
        return max(0.0, self._hermite(self.wsumVals, self.wsumRange, nPerh))

    def __eq__(self, other):
        return (isinstance(other, TableKernel) and self.n1 == other.n1 and self.kernelExtent == other.kernelExtent and
                np.array_equal(self.Wcoef, other.Wcoef) and np.array_equal(self.gradWcoef, other.gradWcoef))

    __hash__ = object.__hash__
