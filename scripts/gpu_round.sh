#!/bin/bash
# One GPU-box visit: parity tests, the bench lines, the ncu launch list and one full capture of the top kernels.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh <tag> [quick]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
if [ "$2" != "quick" ]; then
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_sph_derivs|k_nbr_build|k_tile_runs|k_pack' -s 8 -c 4 -f -o $OUT/prof \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
ls -la $OUT
