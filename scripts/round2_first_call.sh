#!/bin/bash
# First GPU call of round 2 (one B200): verify and time the experimental kernels prepared at the end of round 1.
#   gpurun --timeout 600 -- 'bash scripts/round2_first_call.sh'
# Results go to gpurun_out/r02_first_call/.
set -u
out=gpurun_out/r02_first_call
mkdir -p $out
# 1. parity of kernel_variant 60 (Cartesian operator in fast-diagonalisation form) on the GPU
STFEM_RUN_NEXT=1 timeout 300 python -m pytest tests/test_next_round_gpu.py -q -p no:cacheprovider > $out/pytest_next.log 2>&1
echo "pytest rc=$?" >> $out/pytest_next.log
# 2. headline operator: default kernel vs variant 60 (FP64, 96^3 cells, Q4 x cG(2)); FP32 as used on the multigrid levels
for v in 0 60; do
  timeout 120 python bench.py --variant $v --no-solve --no-perturbed --no-practical --no-cpu-baseline --steps 100 > $out/bench_vmult_v$v.json 2> $out/bench_vmult_v$v.err
  timeout 120 python scripts/tune_vmult.py 96 4 f32 $v > $out/tune_f32_v$v.log 2>&1
done
# 3. the same inside the solve: level operators in variant 60
for v in 0 60; do
  STFEM_LEVEL_VARIANT=$v timeout 200 python bench.py --no-perturbed --no-practical --no-cpu-baseline --steps 20 > $out/bench_solve_levelv$v.json 2> $out/bench_solve_levelv$v.err
done
# 4. ncu of the new kernel (only after the plain runs above exited)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_st_vmult_cart_fd -c 2 -o $out/prof_cart_fd \
  python bench.py --variant 60 --no-solve --no-perturbed --no-practical --no-cpu-baseline --steps 3 --warmup 1 > $out/ncu_cart_fd.log 2>&1
ls -la $out
