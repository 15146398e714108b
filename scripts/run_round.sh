#!/bin/bash
# One GPU-box pass over everything the round is judged on (run under gpurun from the repo root):
# parity tests, smoke, the bench line, the reference arm, the ncu launch list of the bench command.
# Outputs land in gpurun_out/; scripts/summarize_launches.py turns the launch list into profiles/.
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r1.csv python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r1.log 2>&1
echo "ncu launch list rc=$?"
