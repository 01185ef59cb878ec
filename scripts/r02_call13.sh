#!/bin/bash
# Round 2, GPU call 13: brick kernel timing after the chain split; full GPU suite; default bench run.
set -u
out=gpurun_out/r02_call13
mkdir -p $out
timeout 300 python scripts/tune_vmult.py 96 4 f64 3 0 86 > $out/tune_f64.log 2>&1
timeout 300 python scripts/tune_vmult.py 96 4 f32 3 0 > $out/tune_f32.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $out/pytest_gpu_all.log 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/bench_default.json 2> $out/bench_default.err
echo "bench rc=$?" >> $out/bench_default.err
ls -la $out
