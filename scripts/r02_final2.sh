#!/bin/bash
# Round 2, final numbers after the specialised Gram-Schmidt kernels: full GPU suite, smoke, default bench, reference arm.
set -u
out=gpurun_out/r02_final2
mkdir -p $out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $out/pytest_gpu_all.log 2>&1
echo "pytest rc=$?" >> $out/pytest_gpu_all.log
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1
echo "smoke rc=$?" >> $out/smoke.log
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err
echo "bench rc=$?" >> $out/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
tail -3 $out/pytest_gpu_all.log; tail -1 $out/smoke.log; tail -1 $out/bench_default.err
