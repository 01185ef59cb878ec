#!/bin/bash
# Round 2: the driver's scaling command at N GPUs (default legs), with the fail-fast timeout of the library armed.
set -u
N=${1:-8}
out=gpurun_out/r02_n$N
mkdir -p $out
export NCCL_DEBUG=WARN STFEM_SYNC_TIMEOUT_S=45 STFEM_HALO_VERBOSE=1
t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err
echo "bench rc=$? wall=$(( $(date +%s) - t0 )) s" >> $out/bench.err
t0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
echo "bench ref rc=$? wall=$(( $(date +%s) - t0 )) s" >> $out/bench_reference.err
ls -la $out
