#!/bin/bash
TAG=$1; W=2
OUT=gpurun_out/$TAG; mkdir -p $OUT
NCCL_DEBUG=WARN timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $W --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_n$W.json 2> $OUT/bench_n$W.err; echo "bench rc=$?"
python - $OUT/bench_n$W.json <<'PY'
import json,sys
try:
    t=open(sys.argv[1]).read(); d=json.loads([l for l in t.splitlines() if l.startswith('{')][-1]); b=d["breakdown_ms"]
    print("step %.3f ms  value %.1f M/s e2e %.1f (host==device %s) parity %s checksum edges %s weak %s"%(d["ms_per_step"], d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"].get("host_results_equal_device_resident"), (d.get("parity") or {}).get("ok"), d["checksum"]["directed_edges"], (d.get("weak") or {}).get("value")))
except Exception as e:
    print("failed: %s"%e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
MGPU_N=14 MGPU_RK2=1 MGPU_PLANES=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29562 tests/mgpu_parity.py > $OUT/rk2_planes.log 2>&1; echo "mgpu rk2_planes rc=$?"; grep -h '"rank"' $OUT/rk2_planes.log | cut -c1-200
