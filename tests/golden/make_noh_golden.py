#!/usr/bin/env python
"""Golden radial profile of the Noh-cylindrical-2d problem (BASELINE configs[0], scaled to nRadial = 50, t = 0.6), produced by the
ORACLE-driven CheapSynchronousRK2 (tests/common.py::OracleRK2) with two reflecting planes.  The GPU test
tests/test_gpu_boundary.py::test_noh_cylindrical_2d_device_resident runs the same problem with the state resident on the device
and compares.  Regenerate with:  python tests/golden/make_noh_golden.py   (about 40 s of CPU)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common                                   # noqa: E402
from oracle import oracle                       # noqa: E402
from spheral_b200 import kernel as K, nodegen as ng   # noqa: E402

NRADIAL, TEND, NPERH = 50, 0.6, 4.01
BINS = np.linspace(0.0, 0.7, 15)


def noh_2d(nRadial, nPerh):
    pos, mass, H = ng.constant_dtheta_2d(nRadial, nPerh=nPerh)
    N = len(pos)
    r = np.linalg.norm(pos, axis=1)
    st = dict(position=pos, velocity=-pos/r[:, None], H=H, mass=mass, massDensity=np.ones(N), specificThermalEnergy=np.zeros(N),
              pressure=np.zeros(N), soundSpeed=np.zeros(N), omegaGradh=np.ones(N))
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in st.items()}, N


def profile(pos, rho):
    r = np.linalg.norm(pos, axis=1)
    which = np.digitize(r, BINS) - 1
    return [float(rho[which == b].mean()) if np.any(which == b) else 0.0 for b in range(len(BINS) - 1)]


def main():
    oracle.build()
    st, N = noh_2d(NRADIAL, NPERH)
    WT = K.TableKernel(K.WendlandC4Kernel(2), 1000)
    OT = common.oracle_table(oracle, WT)
    oo = oracle.default_options(2, nPerh=NPERH, Cl=1.0, Cq=1.0, hmin=1.0e-4, hmax=0.1)
    so = oracle.default_step_options(hminratio=0.1)
    planes = [(np.zeros(2), np.eye(2)[a]) for a in range(2)]
    rk = common.OracleRK2(oracle, oo, so, OT, st, densityUpdate=1, planes=planes, dtMin=1.0e-8, dtMax=0.1)
    m = st["mass"]
    E0 = float(np.sum(m*0.5*np.sum(st["velocity"]**2, axis=1)))
    rk.initializeDerivatives()
    while rk.t < TEND:
        rk.step(TEND)
    out = dict(nRadial=NRADIAL, tend=TEND, nPerh=NPERH, nodes=N, cycles=rk.cycle, dE_over_E=(rk.total_energy() - E0)/E0,
               bins=BINS.tolist(), rho_profile=profile(rk.s["pos"][:N], rk.s["rho"][:N]))
    json.dump(out, open(os.path.join(HERE, "noh2d_nr50_t06.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
