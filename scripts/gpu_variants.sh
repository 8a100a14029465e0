#!/bin/bash
# Build the library with different -D flag sets ON THE GPU BOX and bench each (quick A/B of tuning constants).
# usage: bash scripts/gpu_variants.sh <tag> "<flags1>" "<flags2>" ...
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for F in "$@"; do
  SPHB200_NVCC_FLAGS="$F" python -m spheral_b200.build --force > $OUT/build.log 2>&1 || { echo "build failed for [$F]"; tail -5 $OUT/build.log; continue; }
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/b.json 2>$OUT/b.err
  python - "$F" $OUT/b.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); b=d["breakdown_ms"]
    print("[%s] step %.3f ms  build %.3f  nbr %.3f  pair %.3f  value %.1f M/s"%(sys.argv[1], d["ms_per_step"], b["build_pairs"], b["neighbor_kernels"], b["pair_kernel"], d["value"]/1e6))
except Exception as e:
    print("[%s] failed: %s"%(sys.argv[1], e))
PY
done
python -m spheral_b200.build --force > /dev/null 2>&1
