#!/bin/bash
# memcheck over the code added at the end of round 2: the chunked evaluate-to-host path (forced on small problems), the fused ghost fill, the C++ ABI smoke
TAG=${1:-san2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
export PYTHONDONTWRITEBYTECODE=1
echo "== memcheck to_host"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scale.py -q -m gpu -x -k "to_host and sedov1m" > $OUT/memcheck_to_host.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" $OUT/memcheck_to_host.log | head -5
echo "== memcheck boundary"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_cabi_cpp.py -q -m gpu -x > $OUT/memcheck_boundary.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" $OUT/memcheck_boundary.log | head -5
echo "== racecheck to_host"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_scale.py -q -m gpu -x -k "to_host and sedov1m and lattice" > $OUT/racecheck_to_host.log 2>&1; echo "rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $OUT/racecheck_to_host.log | head -5
