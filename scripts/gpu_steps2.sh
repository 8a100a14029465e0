#!/bin/bash
# usage: bash scripts/gpu_steps2.sh <tag> : parity, default bench (with the resident RK2 leg), launch list of an RK2 step
TAG=${1:-steps2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?"
python scripts/launch_tail.py $OUT/launches.csv 120
