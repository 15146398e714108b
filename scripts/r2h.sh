#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
for c in 1 2; do
echo "CFG=$c"
SBD_WIN8_CFG=$c timeout 900 python scripts/probe_win.py c3 3 2>&1 | tail -n 2
SBD_WIN8_CFG=$c timeout 600 python scripts/probe_win.py c4 3 2>&1 | tail -n 2
done
