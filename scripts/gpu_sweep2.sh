#!/bin/bash
# BASELINE configs[3] and [4] (round 2): bench lines for a list of workloads on N GPUs, one compact JSON line each.
# usage: bash scripts/gpu_sweep2.sh <tag> <ngpus> <workload> [<workload> ...]
TAG=$1; N=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
PORT=29600
for W in "$@"; do
  PORT=$((PORT+1))
  if [ "$N" = "1" ]; then timeout 900 python bench.py --workload $W --steps 5 --quick > $OUT/b.json 2> $OUT/b.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --workload $W --steps 5 --quick > $OUT/b.json 2> $OUT/b.err; fi
  python - $W $N $OUT/b.json <<'PY' || { echo "{\"workload\": \"$W\", \"gpus\": $N, \"failed\": true}"; tail -3 $OUT/b.err >&2; }
import json,sys
d=json.loads([l for l in open(sys.argv[3]) if l.startswith("{")][-1]); b=d["breakdown_ms"]
P=d["config"]["particles"]
print(json.dumps({"workload": sys.argv[1], "gpus": int(sys.argv[2]), "particles": P, "neighbours": round(d["details"]["neighbours_per_particle"],1),
                  "M_updates_per_s": round(d["value"]/1e6,1), "ms_per_step": round(d["ms_per_step"],3), "build_pairs_ms_rank0": round(b["build_pairs"],3),
                  "pair_kernel_ms_rank0": round(b["pair_kernel"],3), "evaluate_only_M_per_s": round(P/int(sys.argv[2])/b["evaluate"]/1e3*int(sys.argv[2]),1),
                  "fp64_frac_pair_kernel": round(d["roofline_fp64"]["frac"],3), "fp64_frac_whole_call": round(d["roofline_fp64"]["whole_call_frac"],3),
                  "hbm_frac_pair_kernel": round(d["roofline"]["frac"],3), "e2e_M_per_s": round(d["e2e"]["value"]/1e6,1), "directed_edges": d["checksum"]["directed_edges"]}))
PY
done | tee -a $OUT/sweep.jsonl
