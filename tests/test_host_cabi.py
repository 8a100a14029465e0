"""CPU tests of the product's host side: the C-ABI library loads, exports every symbol of include/sphb200.h, the host
TableKernel builder agrees with the oracle's independent restatement, and compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re
import numpy as np
import pytest

from spheral_b200 import _lib, kernel as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(sphlib):
    hdr = open(os.path.join(ROOT, "include", "sphb200.h")).read()
    declared = set(re.findall(r"\b(sphb200_[a-zA-Z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(sphlib, name), "missing export " + name
    assert declared == set(_lib.EXPORTS)
    assert sphlib.sphb200_abi_version() == 5


def test_struct_layouts_match_header(sphlib):
    # field order of the ctypes mirrors follows the header
    hdr = open(os.path.join(ROOT, "include", "sphb200.h")).read()
    body = hdr[hdr.index("typedef struct {\n  int    ndim;"):hdr.index("} sphb200_options;")]
    names = re.findall(r"\b([A-Za-z0-9_]+)\s*(?:,|;)", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    assert names == [f for f, _ in _lib.Options._fields_]


@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_host_table_builder_matches_oracle(sphlib, oracle, ndim, kind):
    kern = {0: K.BSplineKernel, 1: K.WendlandC4Kernel, 2: K.WendlandC2Kernel}[kind](ndim)
    for npts in (100, 1000):
        WT = K.TableKernel(kern, npts)
        OT = oracle.TableKernel(kind, ndim, npts)
        assert WT.kernelExtent == OT.kext and WT.n1 == OT.n1 and WT.xstep == OT.xstep
        scale = np.abs(OT.Wcoef).max()
        assert np.abs(WT.Wcoef - OT.Wcoef).max() <= 1e-9*scale        # Vandermonde cond ~1e6: coefficients agree to ~1e-10
        assert np.abs(WT.gradWcoef - OT.gradWcoef).max() <= 1e-9*np.abs(OT.gradWcoef).max()
        # evaluated values agree far better than the raw coefficients
        for eta in np.linspace(0, WT.kernelExtent*0.999, 300):
            a, b = WT.kernelAndGradValue(float(eta)), OT.kernelAndGradValue(float(eta))
            assert abs(a[0] - b[0]) < 1e-13*scale and abs(a[1] - b[1]) < 1e-12*np.abs(OT.gradWcoef).max()
        assert np.allclose(WT.nperhVals, OT.nperhVals, rtol=1e-9, atol=1e-12)
        assert np.allclose(WT.wsumVals, OT.wsumVals, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("ndim,order", [(1, 5), (2, 3), (3, 5), (3, 7)])
def test_nbspline_table_matches_oracle(sphlib, oracle, ndim, order):
    """NBSplineKernel(order) in the PRODUCT's host table builder (missing 4) against the oracle's restatement of
    Kernel/NBSplineKernel.cc -- the kernel the stock Noh scripts use, and the one the reference's stored golden was produced with."""
    WT = K.TableKernel(K.NBSplineKernel(ndim, order), 1000)
    OT = oracle.TableKernel(oracle.KERNEL_NBSPLINE + order, ndim, 1000)
    assert WT.kernelExtent == OT.kext == float((order + 1)//2) and WT.n1 == OT.n1 and WT.xstep == OT.xstep
    sW, sG = np.abs(OT.Wcoef).max(), np.abs(OT.gradWcoef).max()
    assert np.abs(WT.Wcoef - OT.Wcoef).max() <= 1e-8*sW and np.abs(WT.gradWcoef - OT.gradWcoef).max() <= 1e-8*sG
    for eta in np.linspace(0, WT.kernelExtent*0.999, 200):
        a, b = WT.kernelAndGradValue(float(eta)), OT.kernelAndGradValue(float(eta))
        # the alternating sum cancels ~(k/2)^(k-1) : 1, so two orderings of the same formula differ by that much round-off
        assert abs(a[0] - b[0]) < 2e-11*sW and abs(a[1] - b[1]) < 2e-11*sG
    assert np.allclose(WT.nperhVals, OT.nperhVals, rtol=1e-8, atol=1e-10)
    # unit volume integral in ndim dimensions (the normalisation is Simpson's, 10000 bins: ~1e-9)
    x = np.linspace(0.0, WT.kernelExtent, 20001)
    w = np.array([WT.kernelValue(float(e)) for e in x])
    shell = {1: 2.0*np.ones_like(x), 2: 2.0*np.pi*x, 3: 4.0*np.pi*x*x}[ndim]
    f = shell*w
    assert abs(float(np.sum(0.5*(f[1:] + f[:-1])*np.diff(x))) - 1.0) < 2e-6
    with pytest.raises(ValueError):
        K.NBSplineKernel(3, 0)


def test_table_kernel_python_face(sphlib):
    WT = K.TableKernel(K.BSplineKernel3d(), 1000)
    assert WT.kernelExtent == 2.0
    assert abs(WT.kernelValue(0.0, 1.0) - 1.0/np.pi) < 1e-9
    assert WT.kernelValue(2.0, 1.0) == 0.0 and WT.gradValue(2.5, 1.0) == 0.0
    assert WT == K.TableKernel(K.BSplineKernel3d(), 1000)
    assert not (WT == K.TableKernel(K.WendlandC4Kernel3d(), 1000))
    assert abs(WT.equivalentNodesPerSmoothingScale(WT.equivalentWsum(2.01)) - 2.01) < 0.02


def test_bad_options_are_rejected_without_gpu(sphlib):
    from spheral_b200 import engine
    o = engine.make_options(3, compatibleEnergy=1, evolveTotalEnergy=1)
    h = C.c_void_p()
    assert sphlib.sphb200_create(C.byref(h), 0, C.byref(o)) != 0
    assert b"cannot simultaneously" in sphlib.sphb200_last_error(None)
    o = engine.make_options(4)
    assert sphlib.sphb200_create(C.byref(h), 0, C.byref(o)) != 0


def test_no_cpu_fallback(sphlib):
    """Without a CUDA device the engine refuses to come up (the product never routes through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from spheral_b200 import engine
    with pytest.raises(engine.SPHB200Error, match="no CUDA device"):
        engine.Engine(3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "spheral_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "nodegen.py" and "oracle" not in src.lower().replace("the oracle harness", ""), f
