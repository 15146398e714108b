#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_win8 -s 4 -c 1 -o gpurun_out/r2f_win8_c3s python scripts/probe_win.py c3s 3 > gpurun_out/r2f_ncu_c3s.log 2>&1; echo "ncu c3s rc=$?"
