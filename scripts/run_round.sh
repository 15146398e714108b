python -m pytest tests -m gpu -x -q -k "wide or mid_radius or time_loop or seabreeze_diag_steps or golden" 2>&1 | tail -3 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python scripts/probe_py.py c4 3 > gpurun_out/probe_c4.log 2>&1; grep -E "us/step|checksum" gpurun_out/probe_c4.log
timeout 300 python scripts/probe_step.py c3 2>&1 | grep -E "wall|per-kernel" | cut -c1-160 > gpurun_out/probe_c3_strip.log
cat gpurun_out/probe_c3_strip.log
