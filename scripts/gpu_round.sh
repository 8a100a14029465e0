#!/bin/bash
# One GPU call of the tuning loop: GPU parity suite on the default library, then bench --quick of the default library and of every
# prebuilt variant (scripts/build_variants.py) on the 8 M target and the 1 M Sedov case.
# usage: bash scripts/gpu_round.sh <tag> [pytest-args]
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -x -m gpu "$@" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); b=d["breakdown_ms"]
    print("[%s] step %.3f ms  build %.3f  nbr %.3f  pair %.3f  value %.1f M/s  parity %s"%(sys.argv[1], d["ms_per_step"], b["build_pairs"], b["neighbor_kernels"], b["pair_kernel"], d["value"]/1e6, (d.get("parity") or {}).get("worst_field_error")))
except Exception as e:
    print("[%s] failed: %s"%(sys.argv[1], e))
PY
}
for W in noh8m sedov1m; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --quick --workload $W > $OUT/default_$W.json 2> $OUT/default_$W.err; show "default $W" $OUT/default_$W.json
  for LIB in spheral_b200/variants/libsphb200_*.so; do
    [ -f "$LIB" ] || continue
    V=$(basename $LIB .so); V=${V#libsphb200_}
    SPHB200_LIB=$PWD/$LIB timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --quick --workload $W > $OUT/${V}_$W.json 2> $OUT/${V}_$W.err; show "$V $W" $OUT/${V}_$W.json
  done
done
