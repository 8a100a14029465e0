#!/bin/bash
TAG=$1
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -x -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
for V in 1 0; do
  SPHB200_PACK_VALUES_ONLY=$V timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --rk2 > $OUT/noh_pv$V.json 2> $OUT/noh_pv$V.err
  python - "values_only=$V" $OUT/noh_pv$V.json <<'PY'
import json,sys
d=json.load(open(sys.argv[2])); r=d.get("rk2_step_resident") or {}
print("[%s] noh8m step %.3f  e2e %.2f ms (%.1f M/s)  rk2 %.2f / lazy %.2f"%(sys.argv[1], d["ms_per_step"], d["config"]["particles"]/d["e2e"]["value"]*1e3, d["e2e"]["value"]/1e6, r.get("ms_per_step",0), r.get("ms_per_step_lazy_omega",0)))
PY
done
