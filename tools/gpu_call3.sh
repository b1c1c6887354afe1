#!/bin/bash
# GPU call 3 (4 GPUs): NCCL correctness test, scaling probes
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c3_gpus.txt
timeout 600 python -m pytest tests/test_parallel.py -m gpu -x -q > gpurun_out/c3_pytest_nccl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c3_pytest_nccl.log
run() { # name nproc extra-args
  name=$1; n=$2; shift 2
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n --steps 40 --warmup 15 --no-e2e "$@" > gpurun_out/c3_$name.json 2> gpurun_out/c3_$name.err
  tail -1 gpurun_out/c3_$name.json | python -c "
import json,sys
try:
    b=json.loads(sys.stdin.read()); print('$name', round(b['ms_per_step'],4), b['ms_per_step_spread']['median'], b.get('grad_check',{}).get('rel_err_max_over_ranks'))
except Exception as e: print('$name ERR', e)"
}
timeout 300 python bench.py --steps 40 --warmup 15 --no-e2e --no-cpu-baseline > gpurun_out/c3_n1.json 2> gpurun_out/c3_n1.err
run n2_bwd 2
run n4_bwd 4
run n4_deferred 4 --reduce deferred
run n4_bwd_deg0 4 --sh-degree 0
run n4_deferred_deg0 4 --reduce deferred --sh-degree 0
timeout 300 python bench.py --steps 40 --warmup 15 --no-e2e --no-cpu-baseline --sh-degree 0 > gpurun_out/c3_n1_deg0.json 2> gpurun_out/c3_n1_deg0.err
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING run n4_bwd_dbg 4 ; grep -i -E "nvls|algo|channel" gpurun_out/c3_n4_bwd_dbg.err | head -20 > gpurun_out/c3_nccl_info.txt
tail -3 gpurun_out/c3_pytest_nccl.log
python -c "
import json
for n in ('n1','n1_deg0'):
    b=json.loads(open('gpurun_out/c3_%s.json'%n).read().strip().splitlines()[-1]); print(n, b['ms_per_step'], b['ms_per_step_spread']['median'])"
