#!/bin/bash
# Proxy for the 8-rank regime on fewer GPUs: the decomposed hot path with ~1 M particles per rank (nside chosen so), phase times + bench line.
# usage: bash scripts/gpu_small_ranks.sh <tag> <world> <nside>
TAG=$1; W=$2; NS=$3
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29531 scripts/mgpu_phase_times.py $NS > $OUT/phases.log 2>&1; echo "phases rc=$?"; grep -A12 '^{' $OUT/phases.log | head -14
for TP in 1 0; do
SPHB200_HALO_TWO_PHASE=$TP timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 2953$((2+TP)) bench.py --gpus $W --steps 20 --warmup 5 --quick --nside $NS > $OUT/bench_tp$TP.json 2> $OUT/bench_tp$TP.err; echo "bench rc=$?"
python - "two_phase=$TP" $OUT/bench_tp$TP.json <<'PY'
import json,sys
try:
    t=open(sys.argv[2]).read(); d=json.loads([l for l in t.splitlines() if l.startswith('{')][-1]); b=d["breakdown_ms"]
    print("[%s] step %.3f ms  build %.3f  nbr %.3f  pair %.3f  other %.3f  value %.1f M/s e2e %.1f"%(sys.argv[1], d["ms_per_step"], b["build_pairs"], b["neighbor_kernels"], b["pair_kernel"], d["ms_per_step"]-b["build_pairs"]-b["evaluate"], d["value"]/1e6, d["e2e"]["value"]/1e6))
except Exception as e:
    print("[%s] failed: %s"%(sys.argv[1], e))
PY
done
