#!/bin/bash
# Multi-GPU: parity (incl. host-buffer entry point and practical set-up) + the vmult / e2e legs of bench.py.  Argument: ranks.
set -u
N=${1:-2}
out=gpurun_out/r02_mgpu3_n$N
mkdir -p $out
export NCCL_DEBUG=WARN STFEM_SYNC_TIMEOUT_S=45 STFEM_HALO_VERBOSE=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/mgpu_check.py 2 > $out/mgpu_check.log 2>&1
echo "mgpu_check rc=$?" >> $out/mgpu_check.log
tail -5 $out/mgpu_check.log | cut -c1-1200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 --no-solve --no-extra --no-parity > $out/bench_short.json 2> $out/bench_short.err
echo "bench rc=$?" >> $out/bench_short.err
tail -3 $out/bench_short.err
python -c "
import json
d=json.loads(open('$out/bench_short.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], json.dumps(d['e2e']))
"
