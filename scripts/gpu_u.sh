#!/bin/bash
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest boundary/steps"; timeout 900 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_steps.py -q -x -m gpu > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
for W in noh8m sedov1m; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --quick --workload $W > $OUT/default_$W.json 2> $OUT/default_$W.err
python - "default $W" $OUT/default_$W.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); b=d["breakdown_ms"]
    print("[%s] step %.3f ms  build %.3f  nbr %.3f  pair %.3f  value %.1f M/s e2e %.1f"%(sys.argv[1], d["ms_per_step"], b["build_pairs"], b["neighbor_kernels"], b["pair_kernel"], d["value"]/1e6, d["e2e"]["value"]/1e6))
except Exception as e:
    print("[%s] failed: %s"%(sys.argv[1], e))
PY
bash scripts/gpu_ab.sh $TAG/ab_$W --workload $W
done
