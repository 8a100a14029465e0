"""GPU parity at sizes where the persistent loops stride, tiles span many cells and the capacity-redo paths fire:
the bench workloads themselves (bench.make_inputs: Sedov SPH, Noh ASPH with compressed H and reflecting planes, CRKSPH) against the
oracle, a build that outgrows every buffer sized by the previous one, tiles with more candidate runs than the shared-memory run
table holds, and all of it again with the round-1 neighbour kernel (SPHB200_NBR_V2=0) so that both builders stay pinned."""
import numpy as np
import pytest

import common
from spheral_b200 import kernel as K

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(sphlib):
    import bench
    from spheral_b200 import engine
    return bench, engine


@pytest.fixture(params=["1", "0"], ids=["nbr_v2", "nbr_v1"])
def nbr_version(request, monkeypatch):
    monkeypatch.setenv("SPHB200_NBR_V2", request.param)
    return request.param


@pytest.mark.parametrize("workload,n", [("sedov1m", 56), ("noh8m", 48), ("crksph4m", 32), ("glass:64:32", 40)])
def test_bench_workloads_match_the_oracle(oracle, mods, nbr_version, workload, n):
    """bench.parity_check: the workload's inputs at n^3 (plane ghosts included), pair sets memcmp-equal, every derivative field of
    one evaluateDerivatives call within 1e-10 (SURVEY 8c metric).  56^3 = 176 k nodes = 5.5 k tiles: the 296 persistent CTAs of
    k_sph_derivs stride their tile loop, k_nbr_build runs 1.4 k CTAs."""
    bench, _ = mods
    spec = bench.workload_spec(workload)
    res = bench.parity_check(spec, n, 0, 0)
    assert res["pairs_equal"], res
    assert res["worst_field_err"] <= 1.0e-10, res


def test_xsph_variant_of_the_bench_workload(oracle, mods, nbr_version):
    bench, _ = mods
    res = bench.parity_check(bench.workload_spec("noh8m"), 32, 1, 0)
    assert res["ok"], res


def _pairs_equal(orc, e, st, nInt, nGhost, kext):
    s = common.to_oracle_state(st)
    pi, pj, cnt = orc.pairs(st["position"].shape[1], nInt, nGhost, s["pos"], s["H"], kext)
    npairs = e.build_pairs()
    gi, gj = e.download_pairs()
    assert npairs == len(pi) and np.array_equal(gi, pi) and np.array_equal(gj, pj)
    assert np.array_equal(e.download_neighbor_counts(), cnt[:nInt])
    return len(pi)


def test_capacity_redo_path(oracle, mods, nbr_version):
    """Every variable-size buffer (candidate runs, list staging / chunk records, the sliced-ELL array) is sized by the previous build.
    A small, sparse problem first, then a 40x larger one with 4x the neighbours on the same context: the second build overflows them
    all, the kernels skip the overflowing tiles and the host redoes the build -- the result must be the oracle's."""
    _, engine = mods
    WT = K.TableKernel(K.BSplineKernel(3), 1000)
    e = engine.Engine(3, nPerh=1.0)
    e.set_kernel_table(WT)
    st, nInt, _ = common.make_problem(3, 8, nPerh=1.0, seed=2)
    e.set_nodes(nInt, 0)
    e.upload_state(**st)
    _pairs_equal(oracle, e, st, nInt, 0, WT.kernelExtent)
    st, nInt, _ = common.make_problem(3, 28, nPerh=2.2, seed=3)          # ~340 neighbours per node
    e.set_nodes(nInt, 0)
    e.upload_state(**st)
    np2 = _pairs_equal(oracle, e, st, nInt, 0, WT.kernelExtent)
    assert np2 > 100*nInt
    # and back down: stale large capacities, small problem
    st, nInt, _ = common.make_problem(3, 9, nPerh=1.51, seed=4)
    e.set_nodes(nInt, 0)
    e.upload_state(**st)
    _pairs_equal(oracle, e, st, nInt, 0, WT.kernelExtent)


def test_more_runs_per_tile_than_the_run_table(oracle, mods, nbr_version):
    """Few nodes per cell (nPerh = 0.55: ~1.3 nodes per cell) make a tile of 32 nodes span ~25 cells, whose stencils add up to more than
    RUN_CAP = 128 candidate runs: k_tile_runs walks twice, the builders read run starts past JB_CAP from global memory."""
    _, engine = mods
    WT = K.TableKernel(K.BSplineKernel(3), 1000)
    st, nInt, _ = common.make_problem(3, 30, nPerh=0.55, seed=9)
    e = engine.Engine(3, nPerh=0.55)
    e.set_kernel_table(WT)
    e.set_nodes(nInt, 0)
    e.upload_state(**st)
    _pairs_equal(oracle, e, st, nInt, 0, WT.kernelExtent)
    oo, po = common.opts_pair(oracle, engine, 3, nPerh=0.55)
    s = common.to_oracle_state(st)
    pi, pj, cnt = oracle.pairs(3, nInt, 0, s["pos"], s["H"], WT.kernelExtent)
    ref = oracle.evaluate_derivatives(oo, common.oracle_table(oracle, WT), s, nInt, 0, pi, pj, cnt)
    e.evaluate_derivatives(0.0, 1.0)
    got = e.download_derivs()
    floors = common.physical_floors(st, nInt, 3)
    assert max(common.field_err(got[k], ref[k], nInt, f) for k, f in floors.items()) <= 1.0e-10


def test_morton_jump_tiles_and_ragged_sizes(oracle, mods, nbr_version):
    """Node counts that are no multiple of 32, two far-apart clumps (tiles straddling a jump of the Morton curve: the multi-pass
    path of k_nbr_build2) and random anisotropic H."""
    _, engine = mods
    WT = K.TableKernel(K.BSplineKernel(3), 1000)
    st, nInt, _ = common.make_problem(3, 11, nPerh=1.51, kind="aniso", seed=21)
    st2, n2, _ = common.make_problem(3, 9, nPerh=1.51, kind="aniso", seed=22)
    st2["position"] = st2["position"]*0.5 + np.array([7.3, -2.1, 3.9])
    st2["H"] = st2["H"]*2.0
    both = {k: np.ascontiguousarray(np.concatenate([st[k], st2[k]], axis=0)) for k in st}
    N = nInt + n2
    e = engine.Engine(3, nPerh=1.51, hEvolution=1)
    e.set_kernel_table(WT)
    e.set_nodes(N, 0)
    e.upload_state(**both)
    _pairs_equal(oracle, e, both, N, 0, WT.kernelExtent)


@pytest.mark.parametrize("order", ["lattice", "shuffled"])
@pytest.mark.parametrize("workload,n", [("noh8m", 40), ("sedov1m", 36)])
def test_evaluate_derivatives_to_host_is_bit_identical(mods, monkeypatch, workload, n, order):
    """sphb200_evaluate_derivatives_to_host (pair loop in chunks of the host index range, each chunk's download overlapped with the next
    chunk's computation) against evaluate_derivatives + download_derivs: every field bit-identical, pair accelerations included.
    `lattice`: the generator's order (chunks are slabs, few tiles straddle two chunks).  `shuffled`: a random host order with the
    chunking forced (SPHB200_E2H_FORCE=1) -- every tile holds nodes of every chunk and is visited by all of them -- and, unforced, the
    fallback to the two separate calls."""
    bench, engine = mods
    spec = bench.workload_spec(workload)
    st, N = bench.make_inputs(spec, n=n)
    if order == "shuffled":
        p = np.random.default_rng(5).permutation(N)
        st = {k: np.ascontiguousarray(v[p]) for k, v in st.items()}
    planes = bench.plane_list(spec) if spec.get("planes") else []

    def run(fused, env):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        e = engine.Engine(3, **bench.options_kwargs(spec, 0))
        e.set_kernel_table(K.TableKernel(K.BSplineKernel(3), 1000))
        e.set_nodes(N, 0)
        if planes:
            e.reflect_configure(planes)
        e.upload_state(**st)
        if planes:
            e.reflect_set_ghost_nodes()
        e.build_pairs()
        if fused:
            d = e.evaluate_derivatives_to_host()
            again = e.download_derivs()                # the derivatives also stay on the device
            for k in d:
                assert np.array_equal(d[k], again[k]), k
        else:
            e.evaluate_derivatives(0.0, 1.0)
            d = e.download_derivs()
        pacc = e.download_pair_accelerations()
        launches = e.stats()["launches"]
        e.close()
        return d, pacc, launches

    ref, paccRef, _ = run(False, {})
    for env in ({"SPHB200_E2H_FORCE": "1", "SPHB200_E2H_CHUNKS": "4"}, {"SPHB200_E2H_FORCE": "1", "SPHB200_E2H_CHUNKS": "7"},
                {"SPHB200_E2H_FORCE": "0", "SPHB200_E2H_CHUNKS": "4"}):
        got, pacc, _ = run(True, env)
        for k in ref:
            assert np.array_equal(got[k], ref[k]), (k, env)
        assert np.array_equal(pacc, paccRef), env
