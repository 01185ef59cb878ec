#!/bin/bash
# Round 2, GPU call 26: ncu --set full of the two level kernels on the finest level (profiling switched on after the set-up).
set -u
out=gpurun_out/r02_final
mkdir -p $out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:${LD_LIBRARY_PATH:-}
PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_vanka_fd -s 4 -c 1 -o $out/prof_vanka_fd \
  python scripts/level_kernels.py > $out/ncu_vanka_fd.log 2>&1
PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:st_vmult_brick -s 4 -c 1 -o $out/prof_brick_f32 \
  python scripts/level_kernels.py > $out/ncu_brick_f32.log 2>&1
for n in vanka_fd brick_f32; do
  ncu -i $out/prof_$n.ncu-rep --page details > $out/ncu_details_$n.txt 2>&1
  ncu -i $out/prof_$n.ncu-rep --page raw --csv > $out/ncu_raw_$n.csv 2>&1
  ncu -i $out/prof_$n.ncu-rep --page source --csv > $out/ncu_source_$n.csv 2>&1
  gzip -f $out/ncu_source_$n.csv
  rm -f $out/prof_$n.ncu-rep
done
tail -3 $out/ncu_vanka_fd.log $out/ncu_brick_f32.log
grep -E "Duration|Grid Size" $out/ncu_details_vanka_fd.txt $out/ncu_details_brick_f32.txt
