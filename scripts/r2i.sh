#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_golden.py -m gpu -x -q -k "generic or c2 or golden or device or halo" 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-extra --steps 200 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['ms_per_step']*1e3,'us/step', d['parity_check'], d['roofline']['kernel_ms']*1e3, d['roofline']['step_frac'], d['gpu_launches'])"
PROBE_FLAGS=1 timeout 600 python scripts/probe_step.py c2 2>&1 | tail -4
