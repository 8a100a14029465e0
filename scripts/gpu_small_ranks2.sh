#!/bin/bash
TAG=$1; W=$2; NS=$3
OUT=gpurun_out/$TAG; mkdir -p $OUT
SPHB200_HALO_TIMING=1 NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29531 scripts/mgpu_phase_times.py $NS > $OUT/phases.log 2>&1; echo "phases rc=$?"; grep -v "NCCL" $OUT/phases.log | tail -40
