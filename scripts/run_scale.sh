#!/bin/bash
# bench.py on N GPUs of one box, launched the way the driver does; run as  gpurun --gpus N -- 'bash scripts/run_scale.sh N'
N=${1:-2}
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench n$N rc=$? ${SECONDS}s"
tail -5 gpurun_out/r2_bench_n$N.err
python - $N <<'PY'
import json, sys
try:
    d = json.loads(open('gpurun_out/r2_bench_n%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
    print(d['config']['workload'], d['ms_per_step'], 'ms', d['parity_check'], 'base', d['strong_scaling_base']['ms_per_step'], 'e2e', d['e2e'] and d['e2e']['ms_per_step'], 'roof', d['roofline']['frac'], d['roofline']['kernel'])
    for k,e in d.get('other_configs',{}).items():
        print(k, e['ms_per_step'], 'ms', e['parity_check'], e['roofline']['frac'], e['e2e'] and e['e2e']['ms_per_step'])
except Exception as ex:
    print('parse failed', ex)
PY
