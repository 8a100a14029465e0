#!/bin/bash
# Evidence run of the round-2 end state: parity suite, smoke, the bench as the driver runs it (+ RK2 leg), the reference arm, Sedov 1 M and CRKSPH 4 M lines,
# ncu launch list and ncu --set full of the two dominant kernels of a device-resident step.
# usage: bash scripts/gpu_final3.sh <tag>
TAG=${1:-fin3}
OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
echo "== bench (driver form)"; timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "rc=$?"
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "rc=$?"
echo "== bench noh8m + rk2"; timeout 500 python bench.py --workload noh8m --steps 10 --warmup 3 --no-cpu-baseline --rk2 > $OUT/bench_noh8m_rk2.json 2> $OUT/bench_noh8m_rk2.err; echo "rc=$?"
echo "== bench sedov1m + rk2"; timeout 400 python bench.py --workload sedov1m --steps 10 --warmup 3 --no-cpu-baseline --rk2 > $OUT/bench_sedov1m_rk2.json 2> $OUT/bench_sedov1m_rk2.err; echo "rc=$?"
echo "== bench crksph4m"; timeout 400 python bench.py --workload crksph4m --steps 5 --no-cpu-baseline > $OUT/bench_crksph4m.json 2> $OUT/bench_crksph4m.err; echo "rc=$?"
for f in bench_n1 bench_reference bench_noh8m_rk2 bench_sedov1m_rk2 bench_crksph4m; do python - $OUT/$f.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1].split('/')[-1], '%.2f M/s'%(d['value']/1e6), '%.3f ms'%d['ms_per_step'], d.get('breakdown_ms'), 'e2e %.1f'%(d['e2e']['value']/1e6), (d.get('rk2_step_resident') or {}).get('ms_per_step'), (d.get('rk2_step_resident') or {}).get('ms_per_step_lazy_omega'), (d.get('roofline') or {}).get('frac'), (d.get('roofline_fp64') or {}).get('frac'), (d.get('parity') or {}).get('ok'), (d.get('cpu_baseline') or {}).get('value'))
except Exception as e: print(sys.argv[1], 'failed', e)
PY
done
echo "== ncu launch list (device-resident steps, then end-to-end steps)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick > $OUT/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full (4th device-resident step: k_nbr_build2 + k_sph_derivs)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_sph_derivs|k_nbr_build' -s 6 -c 2 -f -o $OUT/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick > $OUT/ncu_full.log 2>&1; echo "rc=$?"
ls -la $OUT | head -30
