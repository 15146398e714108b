#!/bin/bash
# round 2, call 2: k_win8 parity + timing vs k_win4
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 120 scripts/mb_pipes > gpurun_out/r2b_pipes.txt 2>&1; echo "pipes rc=$?"
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2b_pytest.txt
timeout 600 python scripts/probe_win.py c3s 3,2 > gpurun_out/r2b_win_c3s.txt 2>&1; tail -n 4 gpurun_out/r2b_win_c3s.txt
timeout 600 python scripts/probe_win.py c4 3,2 > gpurun_out/r2b_win_c4.txt 2>&1; tail -n 4 gpurun_out/r2b_win_c4.txt
timeout 900 python scripts/probe_win.py c3 3,2 > gpurun_out/r2b_win_c3.txt 2>&1; tail -n 4 gpurun_out/r2b_win_c3.txt
