#!/bin/bash
# GPU call 11 (8 GPUs): the weak-scaling end point with the final code, degree 3 and degree 0
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/c11_ngpus.txt
run() { name=$1; n=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $n --steps 40 --warmup 15 "$@" > gpurun_out/c11_$name.json 2> gpurun_out/c11_$name.err
  tail -1 gpurun_out/c11_$name.json | python -c "
import json,sys
try:
    b=json.loads(sys.stdin.read()); print('$name', round(b['ms_per_step'],4), b['ms_per_step_spread']['median'], round(b['value'],1), b.get('grad_check',{}).get('rel_err_max_over_ranks'), (b.get('e2e') or {}).get('ms_per_step'))
except Exception as e: print('$name ERR', e)"
}
run n8_bwd 8
run n8_bwd_deg0 8 --sh-degree 0 --no-e2e
run n8_deferred 8 --reduce deferred --no-e2e
run n4_bwd 4 --no-e2e
