#!/bin/bash
# Round 2, GPU call 30: the GPU product on all 96 rows of the reference's tests/tp_01.output.
set -u
out=gpurun_out/r02_call30
mkdir -p $out
timeout 1500 python scripts/gpu_tp01_all_rows.py > $out/gpu_tp01_all_rows.txt 2>&1
tail -12 $out/gpu_tp01_all_rows.txt
