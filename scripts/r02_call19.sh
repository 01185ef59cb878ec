#!/bin/bash
# Round 2, GPU call 19: evidence of the current tree: full GPU suite (incl. coarse-grid GMRES), default bench + reference arm,
# launch list of the bench command (vmult leg only), ncu --set full of the brick kernel (FP64 headline) and of the
# on-the-fly / stored-metric plane kernels.
set -u
out=gpurun_out/r02_call19
mkdir -p $out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $out/pytest_gpu_all.log 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu_all.log
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err
echo "bench rc=$?" >> $out/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
echo "bench ref rc=$?" >> $out/bench_reference.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-solve --no-perturbed --no-practical --no-extra --no-cpu-baseline > $out/bench_short.json 2> $out/bench_short.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_bench_short.csv \
  python bench.py --steps 20 --warmup 5 --no-solve --no-perturbed --no-practical --no-extra --no-cpu-baseline > $out/ncu_bench_short.log 2>&1
python scripts/summarize_launches.py $out/launches_bench_short.csv > $out/summary_bench_short.txt 2>&1
timeout 200 python scripts/tune_vmult.py 96 4 f64 0 > $out/plain_for_ncu.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:st_vmult_brick -s 3 -c 1 -o $out/prof_brick \
  python scripts/tune_vmult.py 96 4 f64 0 > $out/ncu_brick.log 2>&1
DISTORT=0.15 timeout 200 python scripts/tune_vmult.py 96 4 f64 5 6 > $out/plain_perturbed.log 2>&1 &&
DISTORT=0.15 timeout 600 ncu --set full --clock-control none --import-source on -k regex:st_vmult_plane -s 3 -c 1 -o $out/prof_plane_stored \
  python scripts/tune_vmult.py 96 4 f64 5 > $out/ncu_plane_stored.log 2>&1
DISTORT=0.15 timeout 600 ncu --set full --clock-control none --import-source on -k regex:st_vmult_plane -s 3 -c 1 -o $out/prof_plane_otf \
  python scripts/tune_vmult.py 96 4 f64 6 > $out/ncu_plane_otf.log 2>&1
for n in brick plane_stored plane_otf; do
  ncu -i $out/prof_$n.ncu-rep --page details > $out/ncu_details_$n.txt 2>&1
  ncu -i $out/prof_$n.ncu-rep --page raw --csv > $out/ncu_raw_$n.csv 2>&1
  ncu -i $out/prof_$n.ncu-rep --page source --csv > $out/ncu_source_$n.csv 2>&1
  gzip -f $out/ncu_source_$n.csv
done
rm -f $out/prof_plane_stored.ncu-rep $out/prof_plane_otf.ncu-rep
ls -la $out
