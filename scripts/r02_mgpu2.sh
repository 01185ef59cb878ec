#!/bin/bash
# Multi-GPU parity only (scripts/mgpu_check.py incl. the practical set-up on ghost-layer patches).  Argument: number of ranks.
set -u
N=${1:-2}
out=gpurun_out/r02_mgpu2_n$N
mkdir -p $out
export NCCL_DEBUG=WARN STFEM_SYNC_TIMEOUT_S=45 STFEM_HALO_VERBOSE=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/mgpu_check.py 2 > $out/mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> $out/mgpu_check.log
tail -30 $out/mgpu_check.log
