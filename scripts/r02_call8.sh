#!/bin/bash
# Round 2, GPU call 8: brick kernel v2.6 (sleeping waits, odd pitch, zero row, vertex symmetry): parity, timing, ncu.
set -u
out=gpurun_out/r02_call8
mkdir -p $out
timeout 900 python -m pytest tests/test_brick_gpu.py -x -q -p no:cacheprovider > $out/pytest_brick.log 2>&1
echo "pytest rc=$?" >> $out/pytest_brick.log
timeout 300 python scripts/tune_vmult.py 96 4 f64 3 0 71 74 75 84 86 88 > $out/tune_f64.log 2>&1
timeout 300 python scripts/tune_vmult.py 96 4 f32 3 0 74 76 > $out/tune_f32.log 2>&1
timeout 200 python scripts/tune_vmult.py 128 3 f64 3 0 > $out/tune_q3.log 2>&1
timeout 200 python scripts/tune_vmult.py 96 4 f64 0 > $out/plain_for_ncu.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:st_vmult_brick -s 3 -c 1 -o $out/prof_brick \
  python scripts/tune_vmult.py 96 4 f64 0 > $out/ncu_brick.log 2>&1
ls -la $out
