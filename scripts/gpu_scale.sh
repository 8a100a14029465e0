#!/bin/bash
# weak-scaling bench at N GPUs of one box. usage: bash scripts/gpu_scale.sh <tag> <N> [bench args]
TAG=$1; N=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 "$@" > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?"
cat $OUT/bench_n$N.json | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], 'GPUs', d['value']/1e6, 'M/s', d['ms_per_step'], 'ms/step', d['breakdown_ms'], d['config'].get('halo'))
"
tail -5 $OUT/bench_n$N.err
