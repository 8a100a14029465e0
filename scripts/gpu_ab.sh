#!/bin/bash
# A/B of prebuilt library variants (scripts/build_variants.py) on one GPU: bench each through SPHB200_LIB.
# usage: bash scripts/gpu_ab.sh <tag> [bench args...]      (runs every spheral_b200/variants/libsphb200_*.so)
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for LIB in spheral_b200/variants/libsphb200_*.so; do
  V=$(basename $LIB .so); V=${V#libsphb200_}
  SPHB200_LIB=$PWD/$LIB timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --quick "$@" > $OUT/$V.json 2> $OUT/$V.err
  python - "$V" $OUT/$V.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); b=d["breakdown_ms"]
    print("[%s] step %.3f ms  build %.3f  nbr %.3f  pair %.3f  value %.1f M/s"%(sys.argv[1], d["ms_per_step"], b["build_pairs"], b["neighbor_kernels"], b["pair_kernel"], d["value"]/1e6))
except Exception as e:
    print("[%s] failed: %s"%(sys.argv[1], e))
PY
done
