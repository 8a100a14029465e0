#!/bin/bash
TAG=$1
OUT=gpurun_out/$TAG; mkdir -p $OUT
bash scripts/gpu_ab.sh $TAG/noh --workload noh8m
bash scripts/gpu_ab.sh $TAG/sedov --workload sedov1m
bash scripts/gpu_ab.sh $TAG/noh_again --workload noh8m
export PYTHONDONTWRITEBYTECODE=1
echo "== racecheck to_host (default build)"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_scale.py -q -m gpu -x -k "to_host and sedov1m and lattice" > $OUT/racecheck_to_host.log 2>&1; echo "rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $OUT/racecheck_to_host.log | head -5
