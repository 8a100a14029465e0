#!/bin/bash
# ncu launch list (per-kernel device time) of one short bench run. usage: bash scripts/gpu_list.sh <tag> [bench args]
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?"
