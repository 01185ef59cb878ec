#!/bin/bash
# Round 2, GPU call 28: thin first / last slab of the host-buffer pipeline (A/B), host-buffer parity tests.
set -u
out=gpurun_out/r02_call28
mkdir -p $out
timeout 120 python scripts/e2e_slabs.py > $out/e2e.log 2>&1
STFEM_HOST_EQUAL_SLABS=1 timeout 120 python scripts/e2e_slabs.py >> $out/e2e.log 2>&1
timeout 120 python scripts/e2e_slabs.py >> $out/e2e.log 2>&1
timeout 900 python -m pytest tests/test_vmult_gpu.py tests/test_brick_gpu.py -x -q -p no:cacheprovider -k "host or pipeline or slab" > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
cat $out/e2e.log; tail -3 $out/pytest.log
