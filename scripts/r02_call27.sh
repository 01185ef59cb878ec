#!/bin/bash
# Round 2, GPU call 27: specialised Gram-Schmidt kernels (compile-time vector count, two entries per thread): solve timing A/B,
# launch-list summary, solver parity tests.
set -u
out=gpurun_out/r02_call27
mkdir -p $out
timeout 300 python scripts/solve_3d.py 5 4 2 CGP 3 > $out/solve_new.log 2>&1
STFEM_GS_GENERIC=1 timeout 300 python scripts/solve_3d.py 5 4 2 CGP 3 > $out/solve_generic.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_solve.csv python scripts/solve_3d.py 5 4 2 CGP 2 > $out/solve_ncu.log 2>&1
python scripts/summarize_launches.py $out/launches_solve.csv > $out/summary_solve.txt 2>&1
gzip -f $out/launches_solve.csv
timeout 1500 python -m pytest tests/test_tp01_gpu.py tests/test_stmg_gpu.py tests/test_zz_practical_gpu.py tests/test_cpp_facade.py tests/test_cpp_tp01_main.py -x -q -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
grep "^step" $out/solve_new.log $out/solve_generic.log; head -14 $out/summary_solve.txt; tail -3 $out/pytest.log
