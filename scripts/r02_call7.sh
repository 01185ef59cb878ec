#!/bin/bash
# Round 2, GPU call 7: brick kernel v2.5 (mbarrier pipeline, no CTA barrier): parity, timing of tile shapes, ncu; full GPU suite.
set -u
out=gpurun_out/r02_call7
mkdir -p $out
timeout 900 python -m pytest tests/test_brick_gpu.py -x -q -p no:cacheprovider > $out/pytest_brick.log 2>&1
echo "pytest rc=$?" >> $out/pytest_brick.log
timeout 300 python scripts/tune_vmult.py 96 4 f64 3 0 71 73 74 75 76 84 86 88 > $out/tune_f64.log 2>&1
timeout 300 python scripts/tune_vmult.py 96 4 f32 3 0 73 74 76 > $out/tune_f32.log 2>&1
timeout 200 python scripts/tune_vmult.py 128 3 f64 3 0 > $out/tune_q3.log 2>&1
timeout 200 python scripts/tune_vmult.py 96 4 f64 0 > $out/plain_for_ncu.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:st_vmult_brick -s 3 -c 1 -o $out/prof_brick \
  python scripts/tune_vmult.py 96 4 f64 0 > $out/ncu_brick.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $out/pytest_gpu_all.log 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu_all.log
ls -la $out
