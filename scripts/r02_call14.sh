#!/bin/bash
# Round 2, GPU call 14: per-level sizes, brick vs per-cell kernel (FP32 level operators, FP64).
set -u
out=gpurun_out/r02_call14
mkdir -p $out
for c in 12 24 48; do
  timeout 120 python scripts/tune_vmult.py $c 4 f32 3 72 >> $out/levels_f32.log 2>&1
  timeout 120 python scripts/tune_vmult.py $c 4 f64 3 72 >> $out/levels_f64.log 2>&1
done
TT=CGP TR=1 timeout 120 python scripts/tune_vmult.py 48 4 f32 3 72 >> $out/levels_f32_nb1.log 2>&1
ls -la $out
