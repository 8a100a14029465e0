#!/bin/bash
# full GPU suite + end-to-end leg with the default chunking (A/B against the two separate calls)
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -x -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
for V in "fused:1" "plain:0"; do
  IFS=: read NAME F <<< "$V"
  for W in noh8m sedov1m; do
    SPHB200_E2E_FUSED=$F timeout 600 python bench.py --steps 8 --warmup 3 --quick --workload $W > $OUT/${W}_$NAME.json 2> $OUT/${W}_$NAME.err
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
