#!/bin/bash
for L in 1 2 4; do for N in 6 12 24; do
  CT_NVCC_EXTRA="-DCT_CLOSEST_LEAVES=$L -DCT_CLOSEST_LANES=$N" python -m cobbletrace_b200.build --force > /dev/null 2>&1
  echo "== leaves=$L lanes=$N"
  python tools/perf_stages.py dragon4k bunny1080 import640 2>&1 | grep -v "tests:" | sed -e 's/rays=.*primary=/primary=/' -e 's/ emit.*//'
done; done
python -m cobbletrace_b200.build --force > /dev/null 2>&1
