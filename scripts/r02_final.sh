#!/bin/bash
# Round 2, final evidence on one GPU: full GPU suite, smoke, default bench + reference arm, launch lists (bench vmult leg,
# one solve), ncu --set full of the headline kernel, the on-the-fly plane kernel and the two level kernels.
set -u
out=gpurun_out/r02_final
mkdir -p $out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $out/pytest_gpu_all.log 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu_all.log
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1
echo "smoke rc=$?" >> $out/smoke.log
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err
echo "bench rc=$?" >> $out/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
timeout 300 python scripts/level_kernels.py > $out/level_kernels.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-solve --no-perturbed --no-practical --no-extra --no-cpu-baseline > $out/bench_short.json 2> $out/bench_short.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_bench_short.csv \
  python bench.py --steps 20 --warmup 5 --no-solve --no-perturbed --no-practical --no-extra --no-cpu-baseline > $out/ncu_bench_short.log 2>&1
python scripts/summarize_launches.py $out/launches_bench_short.csv > $out/summary_bench_short.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_solve.csv python scripts/solve_3d.py 5 4 2 CGP 2 > $out/solve_ncu.log 2>&1
python scripts/summarize_launches.py $out/launches_solve.csv > $out/summary_solve.txt 2>&1
gzip -f $out/launches_solve.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:st_vmult_brick -s 3 -c 1 -o $out/prof_brick \
  python scripts/tune_vmult.py 96 4 f64 0 > $out/ncu_brick.log 2>&1
DISTORT=0.15 timeout 600 ncu --set full --clock-control none --import-source on -k regex:st_vmult_plane -s 3 -c 1 -o $out/prof_plane_otf \
  python scripts/tune_vmult.py 96 4 f64 6 > $out/ncu_plane_otf.log 2>&1
DISTORT=0.15 timeout 600 ncu --set full --clock-control none --import-source on -k regex:st_vmult_plane -s 3 -c 1 -o $out/prof_plane_stored \
  python scripts/tune_vmult.py 96 4 f64 5 > $out/ncu_plane_stored.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vanka_fd -s 30 -c 1 -o $out/prof_vanka_fd \
  python scripts/level_kernels.py > $out/ncu_vanka_fd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:st_vmult_brick_kernel.float -s 30 -c 1 -o $out/prof_brick_f32 \
  python scripts/level_kernels.py > $out/ncu_brick_f32.log 2>&1
for n in brick plane_otf plane_stored vanka_fd brick_f32; do
  ncu -i $out/prof_$n.ncu-rep --page details > $out/ncu_details_$n.txt 2>&1
  ncu -i $out/prof_$n.ncu-rep --page raw --csv > $out/ncu_raw_$n.csv 2>&1
  ncu -i $out/prof_$n.ncu-rep --page source --csv > $out/ncu_source_$n.csv 2>&1
  gzip -f $out/ncu_source_$n.csv
  rm -f $out/prof_$n.ncu-rep
done
ls -la $out
