#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2m_bench_n8.json 2> gpurun_out/r2m_bench_n8.err; echo "bench n8 rc=$? ${SECONDS}s"
tail -5 gpurun_out/r2m_bench_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/r2m_bench_n8.json').read().strip().splitlines()[-1])
    print(d['config']['workload'], d['ms_per_step'], 'ms', d['parity_check'], 'base', d['strong_scaling_base']['ms_per_step'], 'e2e', d['e2e'] and d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], d['roofline']['kernel'])
    for k,e in d.get('other_configs',{}).items():
        print(k, e['ms_per_step'], 'ms', e['parity_check'], e['roofline']['frac'], e['e2e'] and e['e2e']['ms_per_step'])
except Exception as ex:
    print('parse failed', ex)
PY
