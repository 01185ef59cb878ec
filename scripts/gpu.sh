#!/bin/bash
# Helper for this container: rebuild the library, refuse to go to the GPU with a stale or broken build, then run a script there.
#   scripts/gpu.sh <timeout seconds> <script> [gpurun options]
set -e
cd "$(dirname "$0")/.."
python dealii-stfem_b200/build.py > /tmp/build_gpu.log 2>&1 || { grep -m5 " error" /tmp/build_gpu.log; echo "BUILD FAILED"; exit 1; }
if grep -q " error" /tmp/build_gpu.log; then grep -m5 " error" /tmp/build_gpu.log; echo "BUILD FAILED"; exit 1; fi
t=$1; s=$2; shift 2
/usr/local/graft/bin/gpurun "$@" --timeout "$t" -- "bash $s"
