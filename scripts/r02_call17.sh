#!/bin/bash
# Round 2, GPU call 17: delayed CGS in FGMRES, fast assign stores: vmult + solve timing, full GPU suite.
set -u
out=gpurun_out/r02_call17
mkdir -p $out
timeout 300 python scripts/tune_vmult.py 96 4 f64 3 0 > $out/tune_f64.log 2>&1
timeout 300 python scripts/solve_3d.py 5 4 2 CGP 3 > $out/solve_plain.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $out/pytest_gpu_all.log 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu_all.log
ls -la $out
