#!/bin/bash
# GPU box: rebuild with the device-side bounds checks and run the GPU tests (then restore the normal build).
CT_NVCC_EXTRA="-DCT_DEBUG_BOUNDS=1" python -m cobbletrace_b200.build --force > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/sanitize_case.py 2>&1 | grep -c True
python -m cobbletrace_b200.build --force > /dev/null 2>&1
