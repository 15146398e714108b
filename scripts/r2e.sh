#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "py_ or c4 or mid_radius or time_loop or row_blocks or batched" 2>&1 | tail -4
timeout 600 python scripts/probe_win.py c4 2,3 2>&1 | tail -n 4
