#!/bin/bash
# ncu --set full capture of selected kernels during one bench step.  usage: bash scripts/gpu_prof.sh <tag> <kernel-regex> [skip] [count] [bench args...]
TAG=$1; RE=$2; SKIP=${3:-2}; CNT=${4:-1}; shift 4
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -f -o $OUT/prof \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline "$@" > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
