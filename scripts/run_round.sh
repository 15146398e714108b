P="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $P scripts/check_multigpu.py > gpurun_out/mg2.log 2>&1; echo "rc=$?" >> gpurun_out/mg2.log
timeout 300 $P bench.py --gpus 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_n2_flags.json 2> gpurun_out/bench_n2.err
BENCH_PEER_SYNC=barrier timeout 300 $P bench.py --gpus 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_n2_barrier.json 2>> gpurun_out/bench_n2.err
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_n1_b.json 2>> gpurun_out/bench_n2.err
grep -E "MULTIGPU|DIFFERENT|TIMED|rc=|Error|error" gpurun_out/mg2.log | head; for f in bench_n2_flags bench_n2_barrier bench_n1_b; do python -c "
import json,sys
d=json.load(open('gpurun_out/$f.json')); print('$f', d['ms_per_step'], d['value'])"; done; tail -5 gpurun_out/bench_n2.err
