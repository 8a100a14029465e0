#!/bin/bash
# A/B of run-time switches of the library on one GPU: bench --quick once per environment setting.
# usage: bash scripts/gpu_env_ab.sh <tag> "<bench args>" "VAR=1" "VAR=0 OTHER=2" ...
TAG=$1; ARGS=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
k=0
for E in "$@"; do
  k=$((k+1))
  env $E timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --quick $ARGS > $OUT/v$k.json 2> $OUT/v$k.err
  python - "$E" $OUT/v$k.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); b=d["breakdown_ms"]
    print("[%s] step %.3f ms  build %.3f  nbr %.3f  pair %.3f  value %.1f M/s  fine_walk %s"%(sys.argv[1], d["ms_per_step"], b["build_pairs"], b["neighbor_kernels"], b["pair_kernel"], d["value"]/1e6, d["config"].get("grid_fine_walk")))
except Exception as e:
    print("[%s] failed: %s"%(sys.argv[1], e))
PY
done
