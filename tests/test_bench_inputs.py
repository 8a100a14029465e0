"""bench.py's strong-scaling claim rests on its input generator: `--gpus N` gives rank r the lattice planes ix in [r n/N, (r+1) n/N) of THE
single-GPU problem (every random number is drawn for the whole lattice), so that the N-GPU checksum can be compared with the 1-GPU one.  CPU
check of exactly that, for every workload kind, plus the properties the workload descriptions state."""
import numpy as np
import pytest

import bench


@pytest.mark.parametrize("workload", ["noh8m", "sedov1m", "crksph4m", "glass:64:32"])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_the_union_of_the_slabs_is_the_single_gpu_problem(workload, world):
    spec = bench.workload_spec(workload)
    n = 12
    whole, N = bench.make_inputs(spec, n=n)
    assert N == n**3
    parts = [bench.make_inputs(spec, n=n, slab=(r, world)) for r in range(world)]
    assert sum(M for _, M in parts) == N
    ix = np.tile(np.arange(n), n*n)
    for r, (st, M) in enumerate(parts):
        sel = np.nonzero((ix >= (r*n)//world) & (ix < ((r + 1)*n)//world))[0]
        assert M == len(sel)
        for k, v in whole.items():
            assert np.array_equal(st[k], v[sel]), (workload, world, r, k)          # bit-identical, in the single-GPU order
        lo, hi = ((r*n)//world)/float(n), (((r + 1)*n)//world)/float(n)
        x0 = (sel % n + 0.5)/n                                                     # unjittered x of the slab's nodes
        assert np.all((x0 > lo) & (x0 < hi))


def test_workload_descriptions_hold():
    spec = bench.workload_spec("noh8m")
    st, N = bench.make_inputs(spec, n=10)
    r = np.linalg.norm(st["position"], axis=1)
    assert np.allclose(st["velocity"], -st["position"]/r[:, None])                  # v = -r_hat (Noh-spherical-3d.py)
    assert spec["asph"] and np.abs(st["H"][:, 1]).max() > 0.0                        # anisotropic H tensors
    assert len(bench.plane_list(spec)) == 3                                          # the octant's three reflecting planes
    assert np.all(st["position"] > 0.0)
    o = bench.options_kwargs(spec, 0)
    assert o["XSPH"] == 0 and o["compatibleEnergy"] == 1 and o["correctVelocityGradient"] == 1
    iso, _ = bench.make_inputs(bench.workload_spec("sedov1m"), n=10)
    assert np.all(iso["H"][:, [1, 2, 4]] == 0.0) and np.all(iso["H"][:, 0] == iso["H"][:, 3])      # SPH: H = I/h
