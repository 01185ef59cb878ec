#!/bin/bash
# Round 2, GPU call 15: per-kernel time of one STMG-FGMRES solve (96^3 cells), brick kernel vs per-cell kernel.
set -u
out=gpurun_out/r02_call15
mkdir -p $out
timeout 300 python scripts/solve_3d.py 5 4 2 CGP 2 > $out/solve_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_brick.csv python scripts/solve_3d.py 5 4 2 CGP 2 > $out/solve_ncu.log 2>&1
STFEM_NO_BRICK=1 timeout 300 python scripts/solve_3d.py 5 4 2 CGP 2 > $out/solve_plain_nobrick.log 2>&1 &&
STFEM_NO_BRICK=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_nobrick.csv python scripts/solve_3d.py 5 4 2 CGP 2 > $out/solve_ncu_nobrick.log 2>&1
python scripts/summarize_launches.py $out/launches_brick.csv > $out/summary_brick.txt 2>&1
python scripts/summarize_launches.py $out/launches_nobrick.csv > $out/summary_nobrick.txt 2>&1
gzip -f $out/launches_brick.csv $out/launches_nobrick.csv
ls -la $out
