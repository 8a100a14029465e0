#!/bin/bash
TAG=${1:-san3}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONDONTWRITEBYTECODE=1
echo "== memcheck to_host"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scale.py -q -m gpu -x -k "to_host and sedov1m" > $OUT/memcheck_to_host.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" $OUT/memcheck_to_host.log | head -5
echo "== racecheck to_host"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_scale.py -q -m gpu -x -k "to_host and sedov1m and lattice" > $OUT/racecheck_to_host.log 2>&1; echo "rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $OUT/racecheck_to_host.log | head -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --quick > $OUT/noh8m.json 2> $OUT/noh8m.err
python - $OUT/noh8m.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); b=d["breakdown_ms"]
print("[noh8m] step %.3f ms  build %.3f  nbr %.3f  pair %.3f  value %.1f M/s e2e %.1f"%(d["ms_per_step"], b["build_pairs"], b["neighbor_kernels"], b["pair_kernel"], d["value"]/1e6, d["e2e"]["value"]/1e6))
PY
