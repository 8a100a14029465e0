#!/bin/bash
# Multi-GPU parity (tests/mgpu_parity.py in its modes) on W ranks.  usage: bash scripts/gpu_mgpu.sh <tag> <world> [port]
TAG=$1; W=$2; PORT=${3:-29511}
OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { name=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $PORT tests/mgpu_parity.py > $OUT/$name.log 2>&1; echo "$name rc=$?"; grep -h '"rank"' $OUT/$name.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('   rank %d: ghosts %s  worst %.2e %s' % (d['rank'], d.get('ghosts'), d.get('worst_field_error', d.get('worst_state_error', -1)), 'counts_equal=%s' % d['counts_equal'] if 'counts_equal' in d else 'dt_err %.1e dE/E %.1e' % (d['dt_rel_err'], d['dE_over_E'])))
"; PORT=$((PORT+1)); }
run sph MGPU_N=20
run asph MGPU_N=20 MGPU_ASPH=1
run limitedq MGPU_N=20 MGPU_QKIND=1
run rk2 MGPU_N=14 MGPU_RK2=1
run rk2_planes MGPU_N=14 MGPU_RK2=1 MGPU_PLANES=1
run crk MGPU_N=12 MGPU_CRK=1
