#!/bin/bash
# Round 2, GPU call 25: slab count of the host-buffer pipeline against the copy floor (one GPU).
set -u
out=gpurun_out/r02_call25
mkdir -p $out
for s in 16 8 24 32 48; do STFEM_HOST_SLABS=$s timeout 120 python scripts/e2e_slabs.py >> $out/e2e_slabs.log 2>&1; done
cat $out/e2e_slabs.log
