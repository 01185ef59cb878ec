#!/bin/bash
# Round 2, GPU call 18: on-the-fly geometry kernel: parity + timing against the precomputed-metric kernel.
set -u
out=gpurun_out/r02_call18
mkdir -p $out
timeout 900 python -m pytest tests/test_vmult_gpu.py -x -q -p no:cacheprovider > $out/pytest_vmult.log 2>&1
echo "pytest rc=$?" >> $out/pytest_vmult.log
DISTORT=0.15 timeout 300 python scripts/tune_vmult.py 96 4 f64 5 0 11 > $out/tune_perturbed_f64.log 2>&1
DISTORT=0.15 timeout 300 python scripts/tune_vmult.py 96 4 f32 5 0 11 > $out/tune_perturbed_f32.log 2>&1
DISTORT=0.15 TT=DG TR=2 timeout 300 python scripts/tune_vmult.py 80 3 f64 5 0 11 > $out/tune_perturbed_q3.log 2>&1
ls -la $out
