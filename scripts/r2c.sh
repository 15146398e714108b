#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_win4 -s 4 -c 1 -o gpurun_out/r2p_win4_c4 python scripts/probe_win.py c4 2 > gpurun_out/r2p_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
