#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 600 python scripts/probe_win.py c3s 3 > gpurun_out/r2e_win_c3s.txt 2>&1; tail -n 2 gpurun_out/r2e_win_c3s.txt
timeout 600 python scripts/probe_win.py c4 3 > gpurun_out/r2e_win_c4.txt 2>&1; tail -n 2 gpurun_out/r2e_win_c4.txt
timeout 900 python scripts/probe_win.py c3 3 > gpurun_out/r2e_win_c3.txt 2>&1; tail -n 2 gpurun_out/r2e_win_c3.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide or mid_radius or time_loop or diag_steps" 2>&1 | tail -5
