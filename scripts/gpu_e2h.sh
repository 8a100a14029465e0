#!/bin/bash
# A/B of the end-to-end leg: evaluate_derivatives_to_host (chunked, download overlapped) against evaluate + download.
# usage: bash scripts/gpu_e2h.sh <tag>
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest e2h"; timeout 600 python -m pytest tests/test_gpu_scale.py -q -x -m gpu -k "to_host" > $OUT/pytest_e2h.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_e2h.log
for V in "fused4:1:4" "plain:0:4" "fused2:1:2" "fused8:1:8"; do
  IFS=: read NAME F Q <<< "$V"
  for W in noh8m sedov1m; do
    SPHB200_E2E_FUSED=$F SPHB200_E2H_CHUNKS=$Q timeout 600 python bench.py --steps 8 --warmup 3 --quick --workload $W > $OUT/${W}_$NAME.json 2> $OUT/${W}_$NAME.err
    python - "$W $NAME" $OUT/${W}_$NAME.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); b=d["breakdown_ms"]
    print("[%s] step %.3f ms  pair %.3f  value %.1f M/s  e2e %.1f M/s (%.2f ms)"%(sys.argv[1], d["ms_per_step"], b["pair_kernel"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["config"]["particles"]/d["e2e"]["value"]*1e3))
except Exception as e:
    print("[%s] failed: %s"%(sys.argv[1], e))
PY
  done
done
