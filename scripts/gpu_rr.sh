#!/bin/bash
TAG=$1
OUT=gpurun_out/$TAG; mkdir -p $OUT
for L in default rr0; do
  if [ "$L" = default ]; then unset SPHB200_LIB; else export SPHB200_LIB=$PWD/spheral_b200/variants/libsphb200_$L.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --rk2 > $OUT/noh_$L.json 2> $OUT/noh_$L.err
  timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --quick --workload crksph4m > $OUT/crk_$L.json 2> $OUT/crk_$L.err
  python - $L $OUT/noh_$L.json $OUT/crk_$L.json <<'PY'
import json,sys
d=json.load(open(sys.argv[2])); r=d.get("rk2_step_resident") or {}
c=json.load(open(sys.argv[3]))
print("[%s] noh8m step %.3f  rk2 %.2f / lazy %.2f   crksph4m step %.3f pair %.3f"%(sys.argv[1], d["ms_per_step"], r.get("ms_per_step",0), r.get("ms_per_step_lazy_omega",0), c["ms_per_step"], c["breakdown_ms"]["pair_kernel"]))
PY
done
unset SPHB200_LIB
export PYTHONDONTWRITEBYTECODE=1
echo "== pytest steps/crk/parity"; timeout 600 python -m pytest tests/test_gpu_steps.py tests/test_gpu_crk.py tests/test_gpu_parity.py -q -x -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -2 $OUT/pytest.log
echo "== racecheck step kernels at a size where warps own several tiles"; SPHB200_RACE_N=40 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/race_loops.py > $OUT/racecheck_loops.log 2>&1; echo "rc=$?"; grep -E "RACECHECK SUMMARY|ok|hazard" $OUT/racecheck_loops.log | head -5
