python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_final.json')); print(d['ms_per_step'], d['value'], d['steps'], d['e2e']['ms_per_step'], d['cpu_baseline']['ms_per_step'], d['roofline']['frac'], d['clocks'])"
