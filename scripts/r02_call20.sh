#!/bin/bash
# Round 2, GPU call 20: brick kernel with the plain-product instance (GEN = false) and the per-SM alternation of the warp roles
# (0: on, 90: off); plane kernel with a real loop over the x lines; parity tests of both + the multigrid tests.
set -u
out=gpurun_out/r02_call20
mkdir -p $out
timeout 300 python scripts/tune_vmult.py 96 4 f64 3 0 90 0 90 > $out/tune_f64.log 2>&1
timeout 300 python scripts/tune_vmult.py 96 4 f32 0 90 > $out/tune_f32.log 2>&1
DISTORT=0.15 timeout 300 python scripts/tune_vmult.py 96 4 f64 5 6 > $out/tune_perturbed_f64.log 2>&1
DISTORT=0.15 timeout 300 python scripts/tune_vmult.py 96 4 f32 5 6 > $out/tune_perturbed_f32.log 2>&1
timeout 1500 python -m pytest tests/test_brick_gpu.py tests/test_vmult_gpu.py tests/test_stmg_gpu.py -x -q -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
ls -la $out
