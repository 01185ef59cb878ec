#!/bin/bash
# Round 2, GPU call 3: TMA probe + brick kernel with plain loads (isolates the TMA path).
set -u
out=gpurun_out/r02_call3
mkdir -p $out
timeout 120 scripts/_bin/tma_probe > $out/tma_probe.log 2>&1
echo "probe rc=$?" >> $out/tma_probe.log
STFEM_BRICK_NO_TMA=1 timeout 900 python -m pytest tests/test_brick_gpu.py -x -q -p no:cacheprovider > $out/pytest_brick_plain.log 2>&1
echo "pytest rc=$?" >> $out/pytest_brick_plain.log
STFEM_BRICK_NO_TMA=1 timeout 300 python scripts/tune_vmult.py 96 4 f64 3 0 71 80 82 84 86 > $out/tune_f64_plain.log 2>&1
STFEM_BRICK_NO_TMA=1 timeout 300 python scripts/tune_vmult.py 96 4 f32 3 0 82 > $out/tune_f32_plain.log 2>&1
ls -la $out
