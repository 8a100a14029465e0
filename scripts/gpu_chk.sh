#!/bin/bash
TAG=$1
OUT=gpurun_out/$TAG; mkdir -p $OUT
for W in noh8m sedov1m crksph4m; do
timeout 300 python bench.py --steps 6 --warmup 3 --quick --workload $W > $OUT/$W.json 2> $OUT/$W.err
python - $W $OUT/$W.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2])); b=d["breakdown_ms"]
    print("[%s] step %.3f ms value %.1f M/s e2e %.1f M/s  host == device-resident: %s"%(sys.argv[1], d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"].get("host_results_equal_device_resident")))
except Exception as e:
    print("[%s] failed: %s"%(sys.argv[1], e)); print(open(sys.argv[2].replace('.json','.err')).read()[-800:])
PY
done
