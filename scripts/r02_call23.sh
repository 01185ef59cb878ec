#!/bin/bash
# Round 2, GPU call 23: FGMRES with the second Gram-Schmidt pass only where needed: solve timing (A/B with STFEM_FGMRES_CGS2=1),
# parity of everything that solves.
set -u
out=gpurun_out/r02_call23
mkdir -p $out
timeout 300 python scripts/solve_3d.py 5 4 2 CGP 3 > $out/solve_single.log 2>&1
STFEM_FGMRES_CGS2=1 timeout 300 python scripts/solve_3d.py 5 4 2 CGP 3 > $out/solve_cgs2.log 2>&1
timeout 1500 python -m pytest tests/test_tp01_gpu.py tests/test_stmg_gpu.py tests/test_zz_practical_gpu.py tests/test_cpp_facade.py -x -q -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
grep step $out/solve_single.log $out/solve_cgs2.log; tail -3 $out/pytest.log
