#!/bin/bash
# Round 2, GPU call 29: Kronecker Vanka with lanes ordered [block][cell][x] and the padded float exchange layout: level-kernel
# timing, solve, multigrid parity tests.
set -u
out=gpurun_out/r02_call29
mkdir -p $out
timeout 300 python scripts/level_kernels.py > $out/level_kernels.log 2>&1
timeout 300 python scripts/solve_3d.py 5 4 2 CGP 3 > $out/solve.log 2>&1
timeout 1500 python -m pytest tests/test_stmg_gpu.py tests/test_tp01_gpu.py -x -q -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
cat $out/level_kernels.log; grep "^step" $out/solve.log; tail -3 $out/pytest.log
