#!/bin/bash
# usage: tools/sweep_flags.sh "<python command to time>" "<nvcc flags 1>" "<nvcc flags 2>" ... : rebuild libct_gpu.so on the GPU box with each flag set
CMD=$1; shift
for f in "$@"; do
  CT_NVCC_EXTRA="$f" python -m cobbletrace_b200.build --force > /dev/null 2>&1
  echo "== $f"
  $CMD 2>&1 | grep -v "tests:"
done
python -m cobbletrace_b200.build --force > /dev/null 2>&1
