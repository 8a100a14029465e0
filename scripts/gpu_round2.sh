#!/bin/bash
# GPU-box visit covering every BASELINE config that fits one GPU: parity tests, smoke, the three bench workloads, ncu launch lists.
# Usage (under gpurun, repo root):  bash scripts/gpu_round2.sh <tag>
TAG=${1:-r01b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
echo "== bench sedov1m"; timeout 900 python bench.py > $OUT/bench_sedov1m.json 2> $OUT/bench_sedov1m.err; echo "rc=$?"; cat $OUT/bench_sedov1m.json; tail -3 $OUT/bench_sedov1m.err
echo "== bench noh8m"; timeout 900 python bench.py --workload noh8m --steps 5 --no-cpu-baseline > $OUT/bench_noh8m.json 2> $OUT/bench_noh8m.err; echo "rc=$?"; cat $OUT/bench_noh8m.json; tail -3 $OUT/bench_noh8m.err
echo "== bench crksph4m"; timeout 900 python bench.py --workload crksph4m --steps 5 --no-cpu-baseline > $OUT/bench_crksph4m.json 2> $OUT/bench_crksph4m.err; echo "rc=$?"; cat $OUT/bench_crksph4m.json; tail -3 $OUT/bench_crksph4m.err
echo "== ncu launch list crk (1M)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_crk1m.csv \
   python bench.py --workload crksph4m --n 100 --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_list_crk.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full crk"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_crk' -s 12 -c 4 -f -o $OUT/prof_crk \
   python bench.py --workload crksph4m --n 100 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_crk.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
