#!/bin/bash
# Evidence run: parity, smoke, every single-GPU bench line, the reference arm, launch lists, ncu full of the dominant kernels.
# usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-r01c}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "rc=$?"; tail -2 $OUT/bench_n1.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "rc=$?"
echo "== bench noh8m"; timeout 900 python bench.py --workload noh8m --steps 5 --no-cpu-baseline > $OUT/bench_noh8m.json 2> $OUT/bench_noh8m.err; echo "rc=$?"
echo "== bench crksph4m"; timeout 900 python bench.py --workload crksph4m --steps 5 --no-cpu-baseline > $OUT/bench_crksph4m.json 2> $OUT/bench_crksph4m.err; echo "rc=$?"
for f in bench_n1 bench_reference bench_noh8m bench_crksph4m; do python - $OUT/$f.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.2f M/s'%(d['value']/1e6), '%.3f ms'%d['ms_per_step'], d.get('breakdown_ms'), 'e2e %.1f'%(d['e2e']['value']/1e6), (d.get('rk2_step_resident') or {}).get('ms_per_step'))
except Exception as e: print(sys.argv[1], 'failed', e)
PY
done
echo "== ncu launch list (default bench)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sph_derivs|k_nbr_build|k_tile_runs|k_pack' -s 8 -c 4 -f -o $OUT/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_crk_derivs|k_crk_corrections|k_crk_volume' -s 9 -c 3 -f -o $OUT/prof_crk python bench.py --workload crksph4m --nside 100 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_crk.log 2>&1; echo "rc=$?"
ls -la $OUT | head -30
