python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
SBD_WIN4_MIN=1 timeout 200 python scripts/probe_step.py c2 2>&1 | grep -E "wall|per-kernel" | cut -c1-160 > gpurun_out/probe_c2_win4.log
cat gpurun_out/probe_c2_win4.log
