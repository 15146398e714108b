#!/bin/bash
# One GPU-box pass over everything the round is judged on at N = 1 (run under gpurun from the repo root):
# parity tests, smoke, the bench line, the reference arm, the ncu launch list of the bench command and the
# --set full captures of the dominant kernels.  Outputs land in gpurun_out/; copy what is to be judged into profiles/.
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
T=r2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/${T}_smoke.log
python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_ref.err; echo "reference arm rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_list.log 2>&1
echo "ncu launch list rc=$?"
ncu --set full --import-source on --clock-control none -k regex:"k_scan_quad_regs|k_trigger" -s 8 -c 2 -f -o gpurun_out/${T}_step_c2 \
    python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_c2.log 2>&1
ncu -i gpurun_out/${T}_step_c2.ncu-rep --page details > gpurun_out/${T}_ncu_details_step_c2.txt 2>/dev/null
ncu --set full --import-source on --clock-control none -k regex:k_win8 -s 4 -c 1 -f -o gpurun_out/${T}_win8_c3 \
    python scripts/probe_win.py c3 3 > gpurun_out/${T}_ncu_c3.log 2>&1
ncu -i gpurun_out/${T}_win8_c3.ncu-rep --page details > gpurun_out/${T}_ncu_details_k_win8_c3.txt 2>/dev/null
ncu --set full --import-source on --clock-control none -k regex:k_win4 -s 4 -c 1 -f -o gpurun_out/${T}_win4_c4 \
    python scripts/probe_win.py c4 2 > gpurun_out/${T}_ncu_c4.log 2>&1
ncu -i gpurun_out/${T}_win4_c4.ncu-rep --page details > gpurun_out/${T}_ncu_details_k_win4_c4.txt 2>/dev/null
python scripts/probe_win.py c3 3,2 > gpurun_out/${T}_probe_c3.txt 2>&1
python scripts/probe_win.py c4 2,3 > gpurun_out/${T}_probe_c4.txt 2>&1
(PROBE_ROWS=32 python scripts/probe_tiles.py c3r8 1; PROBE_ROWS=32 python scripts/probe_tiles.py c3r8 2; PROBE_ROWS=16 python scripts/probe_tiles.py c3r8 1) > gpurun_out/${T}_probe_tiles.txt 2>&1
echo done
