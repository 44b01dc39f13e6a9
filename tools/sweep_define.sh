#!/bin/bash
# usage: tools/sweep_define.sh MACRO v1 v2 ... : rebuild libct_gpu.so on the GPU box with -DMACRO=v and time the stage kernels.
M=$1; shift
for v in "$@"; do
  CT_NVCC_EXTRA="-D$M=$v" python -m cobbletrace_b200.build --force > /dev/null 2>&1
  echo "== $M=$v"
  python tools/perf_stages.py dragon4k pcbig1080 bunny1080 import640 2>&1 | grep -v "tests:"
done
python -m cobbletrace_b200.build --force > /dev/null 2>&1
