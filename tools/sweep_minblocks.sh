#!/bin/bash
# Rebuild libct_gpu.so on the GPU box with different CT_MIN_BLOCKS and time the stage kernels.
for mb in "$@"; do
  CT_NVCC_EXTRA="-DCT_MIN_BLOCKS=$mb" python -m cobbletrace_b200.build --force > /dev/null 2>&1
  echo "== CT_MIN_BLOCKS=$mb"
  python tools/perf_stages.py dragon4k pcbig1080 bunny1080 2>&1 | grep -v "tests:"
done
python -m cobbletrace_b200.build --force > /dev/null 2>&1
