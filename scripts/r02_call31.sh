#!/bin/bash
# Round 2, GPU call 31: fused relaxation of the tiny multigrid levels (tiny_relax.cuh): parity tests, solve timing A/B.
set -u
out=gpurun_out/r02_call31
mkdir -p $out
timeout 1500 python -m pytest tests/test_stmg_gpu.py tests/test_tp01_gpu.py tests/test_zz_practical_gpu.py tests/test_cpp_facade.py -x -q -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
timeout 300 python scripts/solve_3d.py 5 4 2 CGP 3 > $out/solve_tiny.log 2>&1
STFEM_NO_TINY_RELAX=1 timeout 300 python scripts/solve_3d.py 5 4 2 CGP 3 > $out/solve_steps.log 2>&1
timeout 300 python scripts/solve_3d.py 3 2 1 DG 3 2 > $out/solve_small_tiny.log 2>&1
STFEM_NO_TINY_RELAX=1 timeout 300 python scripts/solve_3d.py 3 2 1 DG 3 2 > $out/solve_small_steps.log 2>&1
tail -3 $out/pytest.log; grep "^step" $out/solve_tiny.log $out/solve_steps.log $out/solve_small_tiny.log $out/solve_small_steps.log
