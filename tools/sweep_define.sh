#!/bin/bash
# usage: tools/sweep_define.sh "<python command to time>" MACRO v1 v2 ... : rebuild libct_gpu.so on the GPU box with -DMACRO=v
CMD=$1; M=$2; shift 2
for v in "$@"; do
  CT_NVCC_EXTRA="-D$M=$v" python -m cobbletrace_b200.build --force > /dev/null 2>&1
  echo "== $M=$v"
  $CMD 2>&1 | grep -v "tests:"
done
python -m cobbletrace_b200.build --force > /dev/null 2>&1
