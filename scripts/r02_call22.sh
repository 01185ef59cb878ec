#!/bin/bash
# Round 2, GPU call 22: warp-level mbarrier operations (one arrival / one polling lane per warp): default flow against the
# barrier-pipeline flow (79); C++ tp_01 main; brick parity.
set -u
out=gpurun_out/r02_call22
mkdir -p $out
timeout 300 python scripts/tune_vmult.py 96 4 f64 0 79 0 79 > $out/tune_f64.log 2>&1
timeout 300 python scripts/tune_vmult.py 96 4 f32 0 79 > $out/tune_f32.log 2>&1
timeout 300 python scripts/tune_vmult.py 128 3 f64 0 79 > $out/tune_q3.log 2>&1
TT=DG TR=2 timeout 300 python scripts/tune_vmult.py 96 3 f64 0 79 > $out/tune_q3_nb3.log 2>&1
timeout 1500 python -m pytest tests/test_brick_gpu.py tests/test_cpp_tp01_main.py -x -q -p no:cacheprovider > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
cat $out/tune_*.log; tail -3 $out/pytest.log
