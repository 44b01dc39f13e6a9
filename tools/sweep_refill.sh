#!/bin/bash
# k_primary's refill threshold (CT_REFILL_T): stage times of the full frame and of a 1/8 share
for T in "$@"; do
  CT_NVCC_EXTRA="-DCT_REFILL_T=$T" python -m cobbletrace_b200.build --force > /dev/null 2>&1
  echo "== CT_REFILL_T=$T"
  python tools/perf_stages.py dragon4k bunny1080 import640 2>&1 | grep -v "tests:" | sed -e 's/rays=.*primary=/primary=/' -e 's/ emit.*//' -e 's/ shadow.*//'
  CT_RANKS=8 python tools/exp_emulate_ranks.py 2>&1 | grep "ranks=8" | sed -e 's/ emit.*//'
done
python -m cobbletrace_b200.build --force > /dev/null 2>&1
