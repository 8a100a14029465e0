#!/bin/bash
# Parity of everything + CRK timing after the ring changes. usage: bash scripts/gpu_steps.sh <tag>
TAG=${1:-steps}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 $OUT/pytest_gpu.log
echo "== bench crk 1M"; timeout 600 python bench.py --workload crksph4m --n 100 --steps 5 --no-cpu-baseline > $OUT/bench_crk1m.json 2> $OUT/bench_crk1m.err; echo "rc=$?"; python -c "import json;d=json.load(open('$OUT/bench_crk1m.json'));print(d['ms_per_step'],d['breakdown_ms'])"; tail -3 $OUT/bench_crk1m.err
echo "== ncu launch list crk (1M)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_crk1m.csv \
   python bench.py --workload crksph4m --n 100 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_list_crk.log 2>&1; echo "ncu list rc=$?"
python scripts/launch_summary.py $OUT/launches_crk1m.csv | tail -12
