#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "wide or mid_radius or time_loop or diag_steps or c3 or c4 or row_blocks" 2>&1 | tail -4
timeout 900 python scripts/probe_win.py c3 3 2>&1 | tail -n 2 | head -1
