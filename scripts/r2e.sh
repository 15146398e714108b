#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
PROBE_ROWS=32 timeout 600 python scripts/probe_win.py c3r8 3 2>&1 | tail -n 2
PROBE_ROWS=16 timeout 600 python scripts/probe_win.py c3r8 3 2>&1 | tail -n 2
PROBE_ROWS=32 timeout 600 python scripts/probe_win.py c3 3 2>&1 | tail -n 2
timeout 600 python scripts/probe_win.py c4 3,2 2>&1 | tail -n 4
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
