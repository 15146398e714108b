#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:k_win8 -s 4 -c 1 -o gpurun_out/r2b_win8_c3 -f \
    python scripts/probe_win.py c3 3 > gpurun_out/r2b_ncu_c3.log 2>&1
ncu -i gpurun_out/r2b_win8_c3.ncu-rep --page details > gpurun_out/r2b_ncu_details_k_win8_c3.txt 2>/dev/null
grep -E "Duration|Issue Slots Busy|Executed Ipc Active|Active Warps Per|Eligible Warps|No Eligible|Executed Instructions  " gpurun_out/r2b_ncu_details_k_win8_c3.txt
