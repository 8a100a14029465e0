bash scripts/gpu_round_full.sh f10
for L in default spheral_b200/variants/libsphb200_*.so; do
  if [ "$L" = default ]; then unset SPHB200_LIB; else export SPHB200_LIB=$PWD/$L; fi
  timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --quick --workload crksph4m > gpurun_out/f10/crk.json 2> gpurun_out/f10/crk.err
  python - "$L" gpurun_out/f10/crk.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); b=d["breakdown_ms"]
    print("[%s crksph4m] step %.3f ms  build %.3f  pair %.3f  value %.1f M/s"%(sys.argv[1].split("_")[-1], d["ms_per_step"], b["build_pairs"], b["pair_kernel"], d["value"]/1e6))
except Exception as e:
    print("[%s] failed: %s"%(sys.argv[1], e))
PY
done
