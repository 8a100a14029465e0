#!/bin/bash
# multi-GPU check: parity modes, then the bench as the driver launches it.  usage: bash scripts/gpu_mg2.sh <tag> <world>
TAG=$1; W=$2
bash scripts/gpu_mgpu.sh $TAG $W
OUT=gpurun_out/$TAG
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $W --steps 10 --warmup 3 > $OUT/bench_n$W.json 2> $OUT/bench_n$W.err; echo "bench rc=$?"
python - $OUT/bench_n$W.json <<'PY'
import json,sys
try:
    t=open(sys.argv[1]).read(); d=json.loads([l for l in t.splitlines() if l.startswith('{')][-1]); b=d["breakdown_ms"]
    print("step %.3f ms  build %.3f  nbr %.3f  pair %.3f  other %.3f  value %.1f M/s e2e %.1f  checksum %s weak %s"%(d["ms_per_step"], b["build_pairs"], b["neighbor_kernels"], b["pair_kernel"], d["ms_per_step"]-b["build_pairs"]-b["evaluate"], d["value"]/1e6, d["e2e"]["value"]/1e6, d.get("checksum"), (d.get("weak") or {}).get("value")))
except Exception as e:
    print("failed: %s"%e)
PY
