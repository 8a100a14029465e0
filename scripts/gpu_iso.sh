#!/bin/bash
TAG=${1:-iso}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu.log
bash scripts/gpu_variants.sh $TAG "" "-DSPHB200_PAIR_CTAS_ISO=3" "-DSPHB200_PAIR_CTAS_ISO=3 -DSPHB200_PAIR_STAGES=3"
echo "== crk 1M"; python bench.py --workload crksph4m --n 100 --steps 5 --no-cpu-baseline > $OUT/crk.json 2>$OUT/crk.err; python -c "import json;d=json.load(open('$OUT/crk.json'));print(d['ms_per_step'],d['breakdown_ms'])"
