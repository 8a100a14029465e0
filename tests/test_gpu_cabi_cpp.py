"""The C ABI called from C++ (tests/cabi_smoke.cpp: plain C++17, AoS std::vector<double> fields, the call order of the Spheral-side
adapter in INTEGRATION.md section 2).  The program checks the invariants that need no second implementation and writes its inputs and
outputs; the same inputs then go through the ctypes binding and must give the SAME BITS -- one library, two front ends."""
import os
import struct
import subprocess

import numpy as np
import pytest

from spheral_b200 import kernel as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp):
    exe = os.path.join(tmp, "cabi_smoke")
    libdir = os.path.join(ROOT, "spheral_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cabi_smoke.cpp"),
                           "-o", exe, "-L", libdir, "-lsphb200", "-Wl,-rpath," + libdir])
    return exe


def test_cabi_smoke_compiles_and_links_against_the_library(sphlib, tmp_path):
    """CPU part: the header is valid C++17 with -Wall, every symbol the program uses resolves against libsphb200.so."""
    assert os.path.exists(_build(str(tmp_path)))


def _read(path):
    out, raw = {}, open(path, "rb").read()
    off = 0
    while off < len(raw):
        tag = raw[off:off + 32].split(b"\0")[0].decode(); n = struct.unpack_from("<Q", raw, off + 32)[0]
        out[tag] = raw[off + 40:off + 40 + n]; off += 40 + n
    return out


@pytest.mark.gpu
def test_cpp_and_ctypes_front_ends_give_the_same_bits(sphlib, tmp_path):
    from spheral_b200 import engine
    exe = _build(str(tmp_path))
    path = str(tmp_path/"cabi_smoke.bin")
    r = subprocess.run([exe, path, "12"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cabi_smoke ok" in r.stdout
    d = _read(path)
    N, npairs, ndim, n = struct.unpack("<4Q", d["meta"])
    f64 = lambda k, w=1: np.frombuffer(d[k], dtype=np.float64).reshape((-1, w) if w > 1 else (-1,))
    u32 = lambda k: np.frombuffer(d[k], dtype=np.uint32)
    e = engine.Engine(3, nPerh=1.51, Cl=2.0, Cq=2.0, XSPH=1)
    e.set_kernel_table(K.TableKernel(K.BSplineKernel(3), 1000))
    e.set_nodes(N, 0)
    e.upload_state(position=f64("pos", 3), H=f64("H", 6))
    assert e.build_pairs() == npairs
    gi, gj = e.download_pairs()
    assert np.array_equal(gi, u32("pi")) and np.array_equal(gj, u32("pj")) and np.array_equal(e.download_neighbor_counts(), u32("cnt"))
    e.upload_state(velocity=f64("vel", 3), mass=f64("mass"), massDensity=f64("rho"), specificThermalEnergy=f64("eps"), pressure=f64("P"),
                   soundSpeed=f64("cs"), omegaGradh=f64("omega"))
    e.evaluate_derivatives(0.0, 1.0)
    got = e.download_derivs()
    names = dict(DxDt=("DxDt", 3), DrhoDt=("DrhoDt", 1), DvDt=("DvDt", 3), DepsDt=("DepsDt", 1), DvDx=("DvDx", 9), gradRho=("gradRho", 3),
                 M=("M", 9), rhoSum=("rhoSum", 1), normalization=("norm", 1), maxViscousPressure=("maxQ", 1), effViscousPressure=("effQ", 1),
                 XSPHWeightSum=("XW", 1), XSPHDeltaV=("XdV", 3), DHDt=("DHDt", 6), Hideal=("Hideal", 6), massZerothMoment=("m0", 1),
                 massFirstMoment=("m1", 3))
    for k, (tag, w) in names.items():
        assert np.array_equal(np.asarray(got[k]).reshape(-1), f64(tag).reshape(-1)), k
    assert np.array_equal(e.download_pair_accelerations().reshape(-1), f64("pacc"))
    e.update_energy_compatible(1.0e-3)
    assert np.array_equal(e.download_state("specificThermalEnergy")["specificThermalEnergy"], f64("eps1"))
