#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_r2.py -m gpu -x -q 2>&1 | tail -6
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_step_quad -s 6 -c 1 -o gpurun_out/r2j_stepq python bench.py --no-cpu-baseline --no-e2e --steps 6 --warmup 3 > gpurun_out/r2j_ncu.log 2>&1; echo "ncu rc=$?"
