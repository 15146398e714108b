set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r1.log 2>&1; echo "ncu list rc=$?"
ncu --set full --import-source on --clock-control none -k regex:"k_scan_quad_regs|k_trigger" -s 8 -c 2 -o gpurun_out/prof_r1_final -f python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r1b.log 2>&1; echo "ncu full rc=$?"
cut -c1-400 gpurun_out/bench_final.json; cut -c1-300 gpurun_out/bench_ref.json; tail -3 gpurun_out/smoke.log
