#!/bin/bash
# One GPU call of the tuning loop for run-time switches: GPU parity suite on the default settings, then bench --quick on the 8 M target
# and the 1 M Sedov case once per environment setting.
# usage: bash scripts/gpu_round_env.sh <tag> "VAR=1" "VAR=0" ...
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -x -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
for W in noh8m sedov1m; do
  k=0
  for E in "$@"; do
    k=$((k+1))
    env $E timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --quick --workload $W > $OUT/${W}_v$k.json 2> $OUT/${W}_v$k.err
    python - "$E $W" $OUT/${W}_v$k.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); b=d["breakdown_ms"]
    print("[%s] step %.3f ms  build %.3f  nbr %.3f  pair %.3f  value %.1f M/s  edges %s"%(sys.argv[1], d["ms_per_step"], b["build_pairs"], b["neighbor_kernels"], b["pair_kernel"], d["value"]/1e6, d["checksum"]["directed_edges"]))
except Exception as e:
    print("[%s] failed: %s"%(sys.argv[1], e))
PY
  done
done
