"""compute-sanitizer racecheck target: every ring-walk kernel (sum density, grad-h correction, dt vote, compatible energy, the CRKSPH loops, the pair kernel)
on a problem large enough that a warp owns several tiles (SPHB200_RACE_N^3 nodes; the persistent grids are 148 x resident CTAs).
compute-sanitizer --tool racecheck python scripts/race_loops.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from spheral_b200 import engine, kernel as K, _lib as L
n = int(os.environ.get("SPHB200_RACE_N", "40"))
for workload in ("noh8m", "crksph4m"):
    spec = bench.workload_spec(workload)
    st, N = bench.make_inputs(spec, n=n)
    crk = spec.get("hydro") == "crksph"
    e = engine.Engine(3, hydro=(L.HYDRO_CRKSPH if crk else L.HYDRO_SPH), **bench.options_kwargs(spec, 0))
    e.set_kernel_table(K.TableKernel(K.BSplineKernel(3), 1000))
    e.set_nodes(N, 0)
    planes = bench.plane_list(spec) if spec.get("planes") else []
    if planes:
        e.reflect_configure(planes)
    e.upload_state(**st)
    if planes:
        e.reflect_set_ghost_nodes()
    e.build_pairs()
    if crk:
        e.crk_compute_volume(); e.crk_compute_corrections(); e.crk_sum_mass_density(1e-10, 1e10)
    else:
        e.sum_mass_density(); e.compute_omega_gradh()
    e.evaluate_derivatives(0.0, 1.0)
    e.compute_dt(0.25, False)
    e.update_energy_compatible(1.0e-6)
    e.sync()
    print("race_loops ok:", workload, N, "nodes,", e.stats()["launches"], "launches")
    e.close()
