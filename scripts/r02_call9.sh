#!/bin/bash
# Round 2, GPU call 9: brick kernel with hinted waits; split-warp variants 77/78.
set -u
out=gpurun_out/r02_call9
mkdir -p $out
timeout 900 python -m pytest tests/test_brick_gpu.py -x -q -p no:cacheprovider > $out/pytest_brick.log 2>&1
echo "pytest rc=$?" >> $out/pytest_brick.log
timeout 300 python scripts/tune_vmult.py 96 4 f64 3 0 77 78 71 74 86 > $out/tune_f64.log 2>&1
timeout 300 python scripts/tune_vmult.py 96 4 f32 3 0 77 78 > $out/tune_f32.log 2>&1
timeout 200 python scripts/tune_vmult.py 96 4 f64 77 > $out/plain_for_ncu.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:st_vmult_brick -s 3 -c 1 -o $out/prof_brick77 \
  python scripts/tune_vmult.py 96 4 f64 77 > $out/ncu_brick.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:st_vmult_brick -s 3 -c 1 -o $out/prof_brick0 \
  python scripts/tune_vmult.py 96 4 f64 0 > $out/ncu_brick0.log 2>&1
ls -la $out
