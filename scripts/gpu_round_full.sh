#!/bin/bash
# One GPU call: GPU parity suite, then the bench as the driver runs it plus the RK2 leg, on the 8 M target and the 1 M Sedov case.
# usage: bash scripts/gpu_round_full.sh <tag>
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -x -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
for W in noh8m sedov1m; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --rk2 --workload $W > $OUT/$W.json 2> $OUT/$W.err
  python - "$W" $OUT/$W.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); b=d["breakdown_ms"]; r=d.get("rk2_step_resident") or {}
    print("[%s] step %.3f ms  build %.3f  nbr %.3f  pair %.3f  value %.1f M/s  e2e %.1f M/s  rk2 %s / lazy %s  parity %s"%(sys.argv[1], d["ms_per_step"], b["build_pairs"], b["neighbor_kernels"], b["pair_kernel"], d["value"]/1e6, d["e2e"]["value"]/1e6, r.get("ms_per_step"), r.get("ms_per_step_lazy_omega"), (d.get("parity") or {})))
except Exception as e:
    print("[%s] failed: %s"%(sys.argv[1], e))
PY
done
