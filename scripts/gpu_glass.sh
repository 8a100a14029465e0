#!/bin/bash
# BASELINE configs[4]: synthetic glass sweep (jittered lattice), evaluateDerivatives + neighbour build, one GPU.
TAG=${1:-glass}
OUT=gpurun_out/$TAG; mkdir -p $OUT
for W in glass:100:32 glass:100:64 glass:100:128 glass:200:32 glass:200:64 glass:200:128 glass:317:32 glass:317:64; do
  timeout 600 python bench.py --workload $W --steps 5 --no-cpu-baseline > $OUT/b.json 2> $OUT/b.err || { echo "$W failed"; tail -3 $OUT/b.err; continue; }
  python - $W $OUT/b.json <<'PY'
import json,sys
d=json.load(open(sys.argv[2])); b=d["breakdown_ms"]
print(json.dumps({"workload": sys.argv[1], "particles": d["config"]["particles_per_gpu"], "neighbours": round(d["config"]["neighbours_per_particle"],1),
                  "M_updates_per_s": round(d["value"]/1e6,1), "ms_per_step": round(d["ms_per_step"],3), "build_pairs_ms": round(b["build_pairs"],3),
                  "evaluate_ms": round(b["evaluate"],3), "evaluate_only_M_per_s": round(d["config"]["particles_per_gpu"]/b["evaluate"]/1e3,1),
                  "fp64_frac": round(d["roofline_fp64"]["frac"],3), "e2e_M_per_s": round(d["e2e"]["value"]/1e6,1)}))
PY
done | tee $OUT/glass_sweep.jsonl
