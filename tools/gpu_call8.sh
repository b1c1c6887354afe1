#!/bin/bash
# GPU call 8 (2 GPUs): NCCL test + N=2 bench with the final reduction paths; 1-GPU project_bwd occupancy A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parallel.py -m gpu -x -q > gpurun_out/c8_pytest_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/c8_pytest_nccl.log
run() { name=$1; n=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 40 --warmup 15 --no-e2e "$@" > gpurun_out/c8_$name.json 2> gpurun_out/c8_$name.err
  tail -1 gpurun_out/c8_$name.json | python -c "
import json,sys
try:
    b=json.loads(sys.stdin.read()); print('$name', round(b['ms_per_step'],4), b['ms_per_step_spread']['median'], b.get('grad_check',{}).get('rel_err_max_over_ranks'))
except Exception as e: print('$name ERR', e)"
}
run n2_bwd 2
run n2_deferred 2 --reduce deferred
run n2_bwd_deg0 2 --sh-degree 0
for m in 6 5; do
  B200GSR_PBWD_MINB=$m timeout 300 python bench.py --steps 60 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/c8_bench_minb$m.json 2> gpurun_out/c8_bench_minb$m.err
  python -c "
import json
b=json.loads(open('gpurun_out/c8_bench_minb$m.json').read().strip().splitlines()[-1]); print('minb$m', round(b['ms_per_step'],4), b['stages_ms']['project_bwd'])"
done
tail -3 gpurun_out/c8_pytest_nccl.log
