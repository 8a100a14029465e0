#!/bin/bash
# strong-scaling line of the 8 M problem on W GPUs (no CPU baseline / weak leg / in-run parity: those are in the 1-GPU and 2-GPU records)
TAG=$1; W=$2
OUT=gpurun_out/$TAG; mkdir -p $OUT
NCCL_DEBUG=WARN timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $W --steps 20 --warmup 5 --no-cpu-baseline --no-weak --no-parity > $OUT/bench_n$W.json 2> $OUT/bench_n$W.err; echo "bench rc=$?"
python - $OUT/bench_n$W.json <<'PY'
import json,sys
try:
    t=open(sys.argv[1]).read(); d=json.loads([l for l in t.splitlines() if l.startswith('{')][-1]); b=d["breakdown_ms"]
    print("step %.3f ms  build %.3f  nbr %.3f  pair %.3f  other %.3f  value %.1f M/s e2e %.1f  ghosts %s checksum %s"%(d["ms_per_step"], b["build_pairs"], b["neighbor_kernels"], b["pair_kernel"], d["ms_per_step"]-b["build_pairs"]-b["evaluate"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["details"]["ghost_nodes_all_ranks"], d.get("checksum")))
except Exception as e:
    print("failed: %s"%e)
PY
