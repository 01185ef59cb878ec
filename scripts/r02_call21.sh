#!/bin/bash
# Round 2, GPU call 21: state after the plain-product brick instance / conditional loop form of the plane kernel: timings
# (incl. the fast-diagonalisation variant 60 and the split-warp variant 77 for the record), full GPU suite, default bench.
set -u
out=gpurun_out/r02_call21
mkdir -p $out
timeout 300 python scripts/tune_vmult.py 96 4 f64 3 0 90 60 77 79 > $out/tune_f64.log 2>&1
timeout 300 python scripts/tune_vmult.py 96 4 f32 3 0 60 > $out/tune_f32.log 2>&1
DISTORT=0.15 timeout 300 python scripts/tune_vmult.py 96 4 f64 5 6 > $out/tune_perturbed_f64.log 2>&1
DISTORT=0.15 timeout 300 python scripts/tune_vmult.py 96 4 f32 5 6 > $out/tune_perturbed_f32.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $out/pytest_gpu_all.log 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu_all.log
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err
echo "bench rc=$?" >> $out/bench_default.err
ls -la $out
