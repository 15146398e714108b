#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
for sp in 1 2 3 4 8; do echo "split max $sp"; PROBE_SPLIT=$sp timeout 600 python scripts/probe_win.py c3r8 3 2>&1 | tail -n 2 | head -1; done
