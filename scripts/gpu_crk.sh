#!/bin/bash
# CRKSPH check: parity tests, 1M and 4M bench lines, ncu launch list. usage: bash scripts/gpu_crk.sh <tag>
TAG=${1:-crk}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== bench crk 1M"; timeout 600 python bench.py --workload crksph4m --n 100 --steps 5 --no-cpu-baseline > $OUT/bench_crk1m.json 2> $OUT/bench_crk1m.err; echo "rc=$?"; python -c "import json;d=json.load(open('$OUT/bench_crk1m.json'));print(d['ms_per_step'],d['breakdown_ms'])"; tail -3 $OUT/bench_crk1m.err
echo "== bench crksph4m"; timeout 600 python bench.py --workload crksph4m --steps 5 --no-cpu-baseline > $OUT/bench_crksph4m.json 2> $OUT/bench_crksph4m.err; echo "rc=$?"; python -c "import json;d=json.load(open('$OUT/bench_crksph4m.json'));print(d['value'],d['ms_per_step'],d['breakdown_ms'])"; tail -3 $OUT/bench_crksph4m.err
echo "== ncu launch list crk (1M)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_crk1m.csv \
   python bench.py --workload crksph4m --n 100 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_list_crk.log 2>&1; echo "ncu list rc=$?"
python scripts/launch_summary.py $OUT/launches_crk1m.csv | tail -12
