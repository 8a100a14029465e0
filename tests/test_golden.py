"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle must keep reproducing them
(CPU, every round) and the CUDA path must match them through the C ABI (GPU) -- pairs bit-exact, fields within 1e-10."""
import importlib.util
import os

import numpy as np
import pytest

import common

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)
NAMES = sorted(mg.CASES)


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(oracle, sphlib, name):
    g = _load(name)
    out = mg.oracle_outputs(name)
    assert np.array_equal(out["pairs_i"], g["pairs_i"]) and np.array_equal(out["pairs_j"], g["pairs_j"])
    assert np.array_equal(out["counts"], g["counts"])
    for k in g.files:
        if k.startswith("state_"):
            assert np.array_equal(out[k], g[k]), "input generator drifted: " + k
        if k.startswith("deriv_") or k.startswith("crk_") or k.startswith("step_"):
            scale = max(float(np.abs(g[k]).max()), 1e-300)
            assert np.abs(out[k] - g[k]).max() <= 1e-13*scale, k


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_matches_golden(sphlib, name):
    from spheral_b200 import engine
    g = _load(name)
    c, WT, st, nInt, nGhost, opts = mg.build_case(name)
    st = {k[6:]: g[k] for k in g.files if k.startswith("state_")}          # the stored inputs, not regenerated ones
    crk = c.get("hydro") == "crk"
    e = engine.Engine(c["ndim"], options=engine.make_options(c["ndim"], hydro=1 if crk else 0, **opts))
    e.set_kernel_table(WT)
    e.set_nodes(nInt, nGhost)
    e.upload_state(**st)
    npairs = e.build_pairs()
    gi, gj = e.download_pairs()
    assert npairs == len(g["pairs_i"]) and np.array_equal(gi, g["pairs_i"]) and np.array_equal(gj, g["pairs_j"])
    assert np.array_equal(e.download_neighbor_counts(), g["counts"])
    if crk:
        vol0, corr0 = mg.crk_ghost_defaults(c["ndim"], st)
        e.upload_state(volume=vol0, rkCorrections=corr0)
        e.crk_compute_volume()
        e.crk_compute_corrections()
        vol = e.download_state("volume")["volume"]
        corr = e.download_state("rkCorrections")["rkCorrections"]
        assert np.abs(vol - g["crk_volume"]).max() <= 1e-10*np.abs(g["crk_volume"]).max()
        ps = c["ndim"] + 1
        for b in range(ps):                       # C, dC_x, dC_y[, dC_z] blocks carry different powers of 1/h
            blk = slice(b*ps, (b + 1)*ps)
            assert np.abs(corr[:, blk] - g["crk_corrections"][:, blk]).max() <= 1e-10*np.abs(g["crk_corrections"][:, blk]).max()
    e.evaluate_derivatives(0.0, 1.0)
    got = e.download_derivs()
    floors = common.physical_floors(st, nInt, c["ndim"])
    for k, f in floors.items():
        err = common.field_err(got[k], g["deriv_" + k], nInt, f)
        assert err <= 1.0e-10, (k, err)
    if opts.get("compatibleEnergy", 1):
        pa = e.download_pair_accelerations()
        assert np.abs(pa - g["deriv_pairAccelerations"]).max() <= 1e-10*max(np.abs(g["deriv_pairAccelerations"]).max(), 1e-300)
    if crk:
        return
    # the per-step callers against the same fixture
    def rel(a, b):
        return float(np.abs(np.asarray(a)[:nInt] - np.asarray(b)[:nInt]).max()/max(np.abs(np.asarray(b)[:nInt]).max(), 1e-300))
    dt, why, node = e.compute_dt(0.25, False)
    assert abs(dt - g["step_dt"][0]) <= 1e-12*g["step_dt"][0]
    assert engine.L.DT_REASONS.index(why) == int(g["step_dt"][1]) and node == int(g["step_dt"][2])
    e.state_copy()
    e.state_update(engine.make_step_options(), mg.STEP_MULT, False)
    got = e.download_state("position", "velocity", "H", "massDensity", "specificThermalEnergy", "pressure", "soundSpeed")
    names = dict(position="pos", velocity="vel", H="H", massDensity="rho", specificThermalEnergy="eps", pressure="P", soundSpeed="cs")
    for k, o in names.items():
        assert rel(got[k], g["step_state_" + o]) <= 1e-10, k
    e.state_assign()
    e.sum_mass_density()
    e.compute_omega_gradh()
    got = e.download_state("massDensity", "omegaGradh")
    assert rel(got["massDensity"], g["step_sum_density"]) <= 1e-10
    assert rel(got["omegaGradh"], g["step_omega"]) <= 1e-10
