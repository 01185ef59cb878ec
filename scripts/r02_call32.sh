#!/bin/bash
# Round 2, GPU call 32: the cell-batched CPU port on the GPU box's host (reference arm + its pin against the numpy oracle).
set -u
out=gpurun_out/r02_call32
mkdir -p $out
timeout 300 python -m pytest tests/test_oracle_cpu_ref.py -q -p no:cacheprovider > $out/pytest_cpu_ref.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
tail -2 $out/pytest_cpu_ref.log; cut -c1-260 $out/bench_reference.json; lscpu | grep -E "Model name|^CPU\(s\)"
