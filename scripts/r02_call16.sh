#!/bin/bash
# Round 2, GPU call 16: residual in one pass (mode 2) + batched reads in the store phase: solve timing, full GPU suite.
set -u
out=gpurun_out/r02_call16
mkdir -p $out
timeout 300 python scripts/solve_3d.py 5 4 2 CGP 3 > $out/solve_plain.log 2>&1
timeout 300 python scripts/tune_vmult.py 96 4 f64 3 0 > $out/tune_f64.log 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $out/pytest_gpu_all.log 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu_all.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_brick.csv python scripts/solve_3d.py 5 4 2 CGP 2 > $out/solve_ncu.log 2>&1
python scripts/summarize_launches.py $out/launches_brick.csv > $out/summary_brick.txt 2>&1
gzip -f $out/launches_brick.csv
ls -la $out
