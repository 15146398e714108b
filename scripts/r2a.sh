#!/bin/bash
# round 2, call 1: pipe-rate and scan microbenchmarks, new parity tests, static-chain timings
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv | tee gpurun_out/r2a_smi.txt
timeout 120 scripts/mb_pipes > gpurun_out/r2a_pipes.txt 2>&1; echo "pipes rc=$?"
timeout 200 scripts/mb_scan 16 0.44 > gpurun_out/r2a_scan.txt 2>&1; echo "scan rc=$?"
timeout 900 python -m pytest tests/test_gpu_r2.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2a_pytest.txt
timeout 300 python scripts/probe_static.py c2 > gpurun_out/r2a_static_c2.txt 2>&1
timeout 600 python scripts/probe_static.py c4 > gpurun_out/r2a_static_c4.txt 2>&1
timeout 900 python scripts/probe_static.py c3 > gpurun_out/r2a_static_c3.txt 2>&1
tail -3 gpurun_out/r2a_static_c*.txt
