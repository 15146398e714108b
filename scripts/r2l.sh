#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
SECONDS=0
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err; echo "bench n1 rc=$? ${SECONDS}s"
tail -5 gpurun_out/r2l_bench_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2l_bench_n1.json').read().strip().splitlines()[-1])
    print('c2', d['ms_per_step']*1e3, 'us', d['parity_check'], 'e2e', d['e2e']['ms_per_step'], 'dropin', d['e2e_dropin'] and (d['e2e_dropin']['ms_per_step'], d['e2e_dropin']['first_calls_ms']), 'cpu', d['cpu_baseline'] and d['cpu_baseline']['ms_per_step'])
    for k,e in d.get('other_configs',{}).items():
        print(k, e['ms_per_step'], 'ms', e['parity_check'], e['roofline']['frac'], e.get('strong_scaling_base') and e['strong_scaling_base']['ms_per_step'])
except Exception as ex:
    print('parse failed', ex)
PY
