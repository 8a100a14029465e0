#!/usr/bin/env python
"""Generate the golden input/output vectors of tests/golden/*.npz with the CPU oracle.

    python tests/golden/make_golden.py

The reference holds no per-node derivative vectors for this path and cannot be built here, so these
vectors are ORACLE outputs: they freeze the oracle (any later change of oracle/ that moves a value is caught by
tests/test_golden.py) and give the GPU tests a fixture that does not depend on the oracle being rebuilt on the GPU box.
Inputs are regenerated from the seeds stored in each file; the state arrays are stored as well so the fixture is
self-contained."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    # name: (ndim, n, nPerh, kind, kernel, options)
    "sph3d_lattice": dict(ndim=3, n=7, nPerh=1.51, kind="lattice", kernel="BSpline", seed=41, ghosts=False,
                          opts=dict(Cl=2.0, Cq=2.0)),
    "asph3d_aniso": dict(ndim=3, n=6, nPerh=1.51, kind="aniso", kernel="BSpline", seed=42, ghosts=False,
                         opts=dict(Cl=1.0, Cq=1.0, hEvolution=1, XSPH=0)),
    "sph2d_ghosts_wendland": dict(ndim=2, n=14, nPerh=4.01, kind="lattice", kernel="WendlandC4", seed=43, ghosts=True,
                                  opts=dict(Cl=1.0, Cq=1.0)),
    "sph3d_limitedq_tensile": dict(ndim=3, n=6, nPerh=1.51, kind="lattice", kernel="BSpline", seed=44, ghosts=False, negP=True, q=True,
                                   opts=dict(Cl=2.0, Cq=2.0, Qkind=1, balsara=1, epsTensile=0.3, compatibleEnergy=0, evolveTotalEnergy=1)),
    # CRKSPH (hydro="crk"): RK sum volumes and linear corrections are part of the fixture (ghost entries: m/rho, identity)
    "crk3d_lattice_limitedq": dict(ndim=3, n=6, nPerh=1.51, kind="lattice", kernel="BSpline", seed=45, ghosts=False, q=True, hydro="crk",
                                   opts=dict(Cl=1.0, Cq=0.25, Qkind=1)),
    "crk2d_ghosts": dict(ndim=2, n=14, nPerh=2.01, kind="lattice", kernel="BSpline", seed=46, ghosts=True, hydro="crk",
                         opts=dict(Cl=1.0, Cq=1.0)),
}


def build_case(name):
    import common
    from spheral_b200 import kernel as K
    c = CASES[name]
    ndim = c["ndim"]
    kern = {"BSpline": K.BSplineKernel, "WendlandC4": K.WendlandC4Kernel}[c["kernel"]](ndim)
    WT = K.TableKernel(kern, 1000)
    st, nInt, nGhost = common.make_problem(ndim, c["n"], nPerh=c["nPerh"], kind=c["kind"], seed=c["seed"], ghosts=c["ghosts"],
                                           negP=c.get("negP", False), kext=WT.kernelExtent)
    if c.get("q"):
        st = common.add_q_fields(st, ndim, seed=c["seed"])
    opts = dict(c["opts"], nPerh=c["nPerh"])
    return c, WT, st, nInt, nGhost, opts


def crk_ghost_defaults(ndim, st):
    """What the fixture uses as boundary-condition values of volume / RK corrections on ghost nodes."""
    n = st["mass"].shape[0]
    corr0 = np.zeros((n, (ndim + 1)**2))
    corr0[:, 0] = 1.0
    return st["mass"]/st["massDensity"], corr0


STEP_MULT = 1.0e-3


def step_outputs(orc, oo, OT, ndim, st, s, nInt, nGhost, pi, pj, cnt, ref):
    """The per-step callers on the same inputs: sum density, grad-h correction, the dt vote and one
    State::update (full policies, default step options) with the derivatives of this fixture."""
    so = orc.default_step_options()
    out = {"step_sum_density": orc.sum_mass_density(ndim, OT, nInt, nGhost, s["pos"], s["mass"], s["H"], pi, pj, rho=s["rho"]),
           "step_omega": orc.omega_gradh(ndim, OT, nInt, nGhost, s["pos"], s["H"], pi, pj, cnt, omega=s["omega"])}
    dt, why, node = orc.hydro_dt(oo, so, nInt, s["vel"], s["H"], s["rho"], s["cs"], ref, pi, pj)
    out["step_dt"] = np.array([dt, float(orc.DT_REASONS.index(why)), float(node)])
    s_in = dict(s, eps=st["specificThermalEnergy"])
    epsDone = False
    if oo.compatibleEnergy:
        s_in["eps"] = orc.update_energy_compatible(ndim, nInt, nGhost, s["mass"], s["vel"], ref["DvDt"], ref["DepsDt"], pi, pj,
                                                   ref["pairAccelerations"], STEP_MULT, s_in["eps"])
        epsDone = True
    upd = orc.state_update(oo, so, nInt, nGhost, STEP_MULT, False, ref, s_in, epsDone=epsDone)
    out.update({"step_state_" + k: v for k, v in upd.items()})
    return out


def oracle_outputs(name):
    import common
    from oracle import oracle as orc
    c, WT, st, nInt, nGhost, opts = build_case(name)
    ndim = c["ndim"]
    oo = orc.default_options(ndim, **opts)
    s = common.to_oracle_state(st)
    pi, pj, cnt = orc.pairs(ndim, nInt, nGhost, s["pos"], s["H"], WT.kernelExtent)
    OT = common.oracle_table(orc, WT)
    extra = {}
    if c.get("hydro") == "crk":
        vol0, corr0 = crk_ghost_defaults(ndim, st)
        vol = orc.crk_sum_volume(ndim, OT, nInt, nGhost, s["pos"], s["H"], pi, pj, vol=vol0)
        corr = orc.crk_corrections(ndim, OT, nInt, nGhost, s["pos"], s["H"], vol, pi, pj, corr=corr0)
        ref = orc.crk_evaluate_derivatives(oo, OT, s, vol, corr, nInt, nGhost, pi, pj)
        extra = {"crk_volume": vol, "crk_corrections": corr}
    else:
        ref = orc.evaluate_derivatives(oo, OT, s, nInt, nGhost, pi, pj, cnt)
        extra = step_outputs(orc, oo, OT, ndim, st, s, nInt, nGhost, pi, pj, cnt, ref)
    out = {"pairs_i": pi, "pairs_j": pj, "counts": cnt, "nInt": np.int64(nInt), "nGhost": np.int64(nGhost)}
    out.update(extra)
    out.update({"state_" + k: v for k, v in st.items()})
    out.update({"deriv_" + k: v for k, v in ref.items()})
    return out


if __name__ == "__main__":
    for name in CASES:
        out = oracle_outputs(name)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-28s nodes %5d+%4d pairs %7d  %6.1f kB" % (name, out["nInt"], out["nGhost"], len(out["pairs_i"]), os.path.getsize(path)/1e3))
