#!/bin/bash
# Round 2, GPU call 24: tile 7 x 8 cells with one 16-warp CTA per SM (variant 75) against the default.
set -u
out=gpurun_out/r02_call24
mkdir -p $out
timeout 300 python scripts/tune_vmult.py 96 4 f64 0 75 0 75 > $out/tune_f64.log 2>&1
timeout 300 python scripts/tune_vmult.py 96 4 f32 0 75 > $out/tune_f32.log 2>&1
cat $out/tune_*.log
