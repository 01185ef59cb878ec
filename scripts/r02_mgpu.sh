#!/bin/bash
# Multi-GPU check: parity (partitioned vs global) and short bench runs.  Arguments: number of ranks, bench repetitions.
set -u
N=${1:-2}
REPS=${2:-1}
out=gpurun_out/r02_mgpu_n$N
mkdir -p $out
export NCCL_DEBUG=WARN STFEM_SYNC_TIMEOUT_S=45 STFEM_HALO_VERBOSE=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_check.py 2 > $out/mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> $out/mgpu_check.log
for rep in $(seq 1 $REPS); do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$rep bench.py --gpus $N --steps 20 --warmup 5 > $out/bench_rep$rep.json 2> $out/bench_rep$rep.err
  echo "bench rc=$?" >> $out/bench_rep$rep.err
done
ls -la $out
