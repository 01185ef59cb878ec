#!/bin/bash
# Multi-GPU parity with the NCCL fallback of the interface exchange forced (STFEM_HALO_NCCL=1).  Argument: number of ranks.
set -u
N=${1:-2}
out=gpurun_out/r02_mgpu_nccl_n$N
mkdir -p $out
export NCCL_DEBUG=WARN STFEM_SYNC_TIMEOUT_S=45 STFEM_HALO_VERBOSE=1 STFEM_HALO_NCCL=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 scripts/mgpu_check.py 2 > $out/mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> $out/mgpu_check.log
tail -6 $out/mgpu_check.log | cut -c1-1500
