#!/bin/bash
# compute-sanitizer passes over small problems: memcheck on smoke() and on a few parity tests, racecheck + synccheck on the ring kernels.
TAG=${1:-san}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONDONTWRITEBYTECODE=1
echo "== memcheck smoke"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/memcheck_smoke.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|smoke ok|Invalid|Error" $OUT/memcheck_smoke.log | head -5
echo "== memcheck steps/boundary/crk subset"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_steps.py tests/test_gpu_boundary.py tests/test_gpu_crk.py -q -m gpu -x -k "not noh" > $OUT/memcheck_tests.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" $OUT/memcheck_tests.log | head -5
echo "== racecheck smoke"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/racecheck_smoke.log 2>&1; echo "rc=$?"; grep -E "RACECHECK SUMMARY|smoke ok|hazard" $OUT/racecheck_smoke.log | head -5
echo "== racecheck step kernels"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_steps.py -q -m gpu -x -k "sum_density or compute_dt or crksph" > $OUT/racecheck_steps.log 2>&1; echo "rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $OUT/racecheck_steps.log | head -5
