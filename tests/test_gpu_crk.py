"""GPU parity tests of the CRKSPH path (BASELINE config 4) against the CPU oracle, through the C ABI.

Bars: RK volumes, RK corrections and every derivative field of one CRKSPH evaluateDerivatives call within 1e-10 relative
(field-wise max-norm metric of SURVEY.md 8c over internal nodes).  The corrections are compared coefficient block by
coefficient block because C, dC carry different powers of 1/h.
"""
import numpy as np
import pytest

import common
from spheral_b200 import kernel as K, _lib as L

pytestmark = pytest.mark.gpu
TOL = 1.0e-10


@pytest.fixture(scope="module")
def eng_mod(sphlib):
    from spheral_b200 import engine
    return engine


def corr_err(got, ref, nInt, ndim):
    """Worst block-relative error: blocks are the columns of the (1+ndim)^2 coefficient record."""
    got, ref = got[:nInt], ref[:nInt]
    worst = 0.0
    for q in range(ref.shape[1]):
        scale = max(np.abs(ref[:, q]).max(), 1e-300)
        # columns that are identically ~0 on symmetric lattices are measured against the block of the same order
        ps = ndim + 1
        blk = ref[:, (q//ps)*ps:(q//ps + 1)*ps] if q % ps else ref[:, q:q + 1]
        scale = max(scale, np.abs(blk).max()*1e-3) if q % ps else scale
        worst = max(worst, np.abs(got[:, q] - ref[:, q]).max()/scale)
    return worst


def run_crk(oracle, eng_mod, ndim, st, nInt, nGhost, WT, ghost_fill=True, **kw):
    oo = oracle.default_options(ndim, **kw)
    po = eng_mod.make_options(ndim, hydro=L.HYDRO_CRKSPH, **kw)
    OT = common.oracle_table(oracle, WT)
    s = common.to_oracle_state(st)
    pi, pj, cnt = oracle.pairs(ndim, nInt, nGhost, s["pos"], s["H"], WT.kernelExtent)
    n = nInt + nGhost
    # ghost volumes / corrections come from the "boundary condition": here simply mass/rho and the identity correction
    vol0 = s["mass"]/s["rho"]
    rvol = oracle.crk_sum_volume(ndim, OT, nInt, nGhost, s["pos"], s["H"], pi, pj, vol=vol0)
    corr0 = np.zeros((n, (ndim + 1)**2)); corr0[:, 0] = 1.0
    rcorr = oracle.crk_corrections(ndim, OT, nInt, nGhost, s["pos"], s["H"], rvol, pi, pj, corr=corr0)
    ref = oracle.crk_evaluate_derivatives(oo, OT, s, rvol, rcorr, nInt, nGhost, pi, pj)
    rrho = oracle.crk_sum_density(ndim, OT, nInt, nGhost, s["pos"], s["mass"], rvol, s["H"], pi, pj,
                                  rhoMin=1e-10, rhoMax=1e10, rho=s["rho"])

    e = eng_mod.Engine(ndim, options=po)
    e.set_kernel_table(WT)
    e.set_nodes(nInt, nGhost)
    e.upload_state(volume=vol0, rkCorrections=corr0, **st)          # ghost entries of volume / corrections
    npairs = e.build_pairs()
    assert npairs == len(pi)
    e.crk_compute_volume()
    gvol = e.download_state("volume")["volume"]
    e.crk_compute_corrections()
    gcorr = e.download_state("rkCorrections")["rkCorrections"]
    e.evaluate_derivatives(0.0, 1.0)
    got = e.download_derivs()
    if po.compatibleEnergy:
        got["pairAccelerations"] = e.download_pair_accelerations()
    return dict(ref=ref, got=got, rvol=rvol, gvol=gvol, rcorr=rcorr, gcorr=gcorr, rrho=rrho, engine=e, pairs=(pi, pj))


CRK_FIELDS = ("DxDt", "DrhoDt", "DvDt", "DepsDt", "DvDx", "localDvDx", "maxViscousPressure", "effViscousPressure",
              "XSPHDeltaV", "DHDt", "Hideal", "massZerothMoment", "massFirstMoment")


def assert_crk_parity(r, st, nInt, nGhost, ndim):
    assert np.abs(r["gvol"][:nInt] - r["rvol"][:nInt]).max() <= TOL*np.abs(r["rvol"][:nInt]).max()
    assert np.array_equal(r["gvol"][nInt:], r["rvol"][nInt:])           # ghost entries untouched
    ce = corr_err(r["gcorr"], r["rcorr"], nInt, ndim)
    assert ce <= TOL, "RK corrections differ: %.3e" % ce
    floors = common.physical_floors(st, nInt, ndim)
    worst = {k: common.field_err(r["got"][k], r["ref"][k], nInt, floors[k]) for k in CRK_FIELDS}
    bad = {k: v for k, v in worst.items() if not v <= TOL}
    assert not bad, "fields out of tolerance: %s (all: %s)" % (bad, worst)
    if "pairAccelerations" in r["got"]:
        a, b = r["got"]["pairAccelerations"], r["ref"]["pairAccelerations"]
        assert a.shape == b.shape
        assert np.abs(a - b).max() <= TOL*max(np.abs(b).max(), floors["DvDt"]*1e-3)
    return worst


@pytest.mark.parametrize("ndim,n,nPerh", [(3, 12, 1.51), (2, 36, 2.01)])
def test_crk_lattice_mg(oracle, eng_mod, ndim, n, nPerh):
    st, nInt, nGhost = common.make_problem(ndim, n, nPerh=nPerh, seed=31)
    WT = K.TableKernel(K.BSplineKernel(ndim), 1000)
    r = run_crk(oracle, eng_mod, ndim, st, nInt, nGhost, WT, nPerh=nPerh, Cl=1.0, Cq=0.75)
    assert_crk_parity(r, st, nInt, nGhost, ndim)


@pytest.mark.parametrize("ndim", [2, 3])
@pytest.mark.parametrize("qflags", [dict(Qkind=1), dict(Qkind=1, balsara=1)])
def test_crk_limited_mg_default_q(oracle, eng_mod, ndim, qflags):
    # the CRKSPH factory default: LimitedMonaghanGingoldViscosity (CRKSPHHydros.py:65-68)
    st, nInt, nGhost = common.make_problem(ndim, 10 if ndim == 3 else 30, nPerh=1.51 if ndim == 3 else 2.01, seed=37)
    st = common.add_q_fields(st, ndim)
    WT = K.TableKernel(K.BSplineKernel(ndim), 1000)
    r = run_crk(oracle, eng_mod, ndim, st, nInt, nGhost, WT, nPerh=1.51, Cl=1.0, Cq=0.25, **qflags)
    assert_crk_parity(r, st, nInt, nGhost, ndim)


@pytest.mark.parametrize("flags", [dict(XSPH=0), dict(compatibleEnergy=0, evolveTotalEnergy=1), dict(hEvolution=1), dict(hEvolution=2)])
def test_crk_flag_variants(oracle, eng_mod, flags):
    st, nInt, nGhost = common.make_problem(3, 10, nPerh=1.51, seed=41)
    WT = K.TableKernel(K.WendlandC4Kernel(3), 1000)
    r = run_crk(oracle, eng_mod, 3, st, nInt, nGhost, WT, nPerh=4.01/2.0, **flags)
    assert_crk_parity(r, st, nInt, nGhost, 3)


def test_crk_anisotropic_H(oracle, eng_mod):
    st, nInt, nGhost = common.make_problem(3, 11, nPerh=1.3, kind="aniso", seed=43)
    WT = K.TableKernel(K.BSplineKernel(3), 1000)
    r = run_crk(oracle, eng_mod, 3, st, nInt, nGhost, WT, nPerh=2.01, hEvolution=1)
    assert_crk_parity(r, st, nInt, nGhost, 3)


@pytest.mark.parametrize("ndim,n", [(3, 11), (2, 30)])
def test_crk_classic_asph_ideal_H(oracle, eng_mod, ndim, n):
    """ASPHClassicSmoothingScale behind the CRKSPH hydro (CRKSPHHydros.py with ASPH = "Classic"): k_crk_derivs followed by k_asph_classic."""
    st, nInt, nGhost = common.make_problem(ndim, n, nPerh=1.3 if ndim == 3 else 2.01, kind="aniso", seed=47)
    WT = K.TableKernel(K.BSplineKernel(ndim), 1000)
    hb = 1.0/st["H"][:nInt, 0].mean()
    r = run_crk(oracle, eng_mod, ndim, st, nInt, nGhost, WT, nPerh=2.01, hEvolution=3, hmin=0.02*hb, hmax=50.0*hb, hminratio=0.1)
    assert_crk_parity(r, st, nInt, nGhost, ndim)
    assert np.abs(r["got"]["Hideal"][:nInt, 1]).max() > 0.0


def test_crk_with_ghosts(oracle, eng_mod):
    st, nInt, nGhost = common.make_problem(2, 24, nPerh=2.01, seed=47, ghosts=True)
    assert nGhost > 0
    WT = K.TableKernel(K.BSplineKernel(2), 1000)
    r = run_crk(oracle, eng_mod, 2, st, nInt, nGhost, WT, nPerh=2.01)
    assert_crk_parity(r, st, nInt, nGhost, 2)
    assert np.all(r["got"]["DvDt"][nInt:] == 0.0)


def test_crk_sum_density_and_compatible_energy(oracle, eng_mod):
    ndim = 3
    st, nInt, nGhost = common.make_problem(ndim, 11, nPerh=1.51, seed=53)
    WT = K.TableKernel(K.BSplineKernel(ndim), 1000)
    r = run_crk(oracle, eng_mod, ndim, st, nInt, nGhost, WT, nPerh=1.51)
    e = r["engine"]
    pi, pj = r["pairs"]
    # compatible energy update on the device vs the oracle's pair sweep over the oracle's own derivatives
    dt = 1.0e-3
    ref = r["ref"]
    eps_ref = oracle.update_energy_compatible(ndim, nInt, nGhost, st["mass"], st["velocity"], ref["DvDt"], ref["DepsDt"],
                                              pi, pj, ref["pairAccelerations"], dt, st["specificThermalEnergy"])
    e.update_energy_compatible(dt)
    eps_got = e.download_state("specificThermalEnergy")["specificThermalEnergy"]
    assert np.abs(eps_got - eps_ref).max() <= TOL*np.abs(eps_ref).max()
    # total energy is conserved by the device path itself
    m, v = st["mass"], st["velocity"]
    v1 = v + dt*r["got"]["DvDt"]
    E0 = (m*(0.5*(v**2).sum(axis=1) + st["specificThermalEnergy"])).sum()
    E1 = (m*(0.5*(v1**2).sum(axis=1) + eps_got)).sum()
    assert abs(E1 - E0) <= 1.0e-12*abs(E0)
    # sum density (overwrites the device mass density)
    e.crk_sum_mass_density(1e-10, 1e10)
    rho = e.download_state("massDensity")["massDensity"]
    assert np.abs(rho[:nInt] - r["rrho"][:nInt]).max() <= TOL*np.abs(r["rrho"][:nInt]).max()


def test_crk_calls_fail_on_sph_context(eng_mod):
    e = eng_mod.Engine(3)
    with pytest.raises(eng_mod.SPHB200Error):
        e.crk_compute_volume()
