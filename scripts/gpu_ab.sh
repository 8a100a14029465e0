#!/bin/bash
# A/B of the aux-copy removal + parity of the changed kernels + step/CRK launch lists. usage: bash scripts/gpu_ab.sh <tag>
TAG=${1:-ab}
OUT=gpurun_out/$TAG; mkdir -p $OUT
bash scripts/gpu_variants.sh $TAG "-DSPHB200_PAIR_AUX=1" "-DSPHB200_PAIR_AUX=0" "-DSPHB200_PAIR_AUX=0 -DSPHB200_PAIR_STAGES=5" "-DSPHB200_PAIR_AUX=0 -DSPHB200_PAIR_STAGES=3"
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?"
python scripts/launch_tail.py $OUT/launches.csv 140 | awk '{a[$1]+=$(NF-1); n[$1]++} END {for (k in a) printf "%-50s %10.1f us  x%d\n", k, a[k], n[k]}' | sort -k2 -n -r | head -14
echo "== ncu launch list crk (1M)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_crk1m.csv \
   python bench.py --workload crksph4m --n 100 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_list_crk.log 2>&1; echo "ncu list rc=$?"
python scripts/launch_summary.py $OUT/launches_crk1m.csv | tail -9
