#!/bin/bash
TAG=$1
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -x -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 300 python scripts/time_classic.py noh8m 2>&1 | tail -1
timeout 300 python bench.py --steps 8 --warmup 3 --quick > $OUT/noh8m.json 2> $OUT/noh8m.err
python - $OUT/noh8m.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); b=d["breakdown_ms"]
print("[noh8m] step %.3f ms  build %.3f  nbr %.3f  pair %.3f  value %.1f M/s e2e %.1f"%(d["ms_per_step"], b["build_pairs"], b["neighbor_kernels"], b["pair_kernel"], d["value"]/1e6, d["e2e"]["value"]/1e6))
PY
