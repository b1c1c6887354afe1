#!/bin/bash
# GPU call 9 (2 GPUs): where do the multi-ms outlier steps of the N>1 runs come from?
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 60 --warmup 15 --no-e2e > gpurun_out/c9_$name.json 2> gpurun_out/c9_$name.err
  tail -1 gpurun_out/c9_$name.json | python -c "
import json,sys
b=json.loads(sys.stdin.read()); print('$name', round(b['ms_per_step'],4), b['ms_per_step_spread'])"
}
run base1
run base2
TORCH_NCCL_AVOID_RECORD_STREAMS=1 run avoid1
TORCH_NCCL_AVOID_RECORD_STREAMS=1 run avoid2
BENCH_NO_CLOCKS=1 run noclk1
BENCH_NO_CLOCKS=1 run noclk2
# 1-GPU: project_sh with hoisted loads
timeout 300 python bench.py --steps 60 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
python -c "
import json
b=json.loads(open('gpurun_out/c9_bench.json').read().strip().splitlines()[-1]); print('n1', round(b['ms_per_step'],4), b['ms_per_step_spread'], {k:round(v,4) for k,v in b['stages_ms'].items()})"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -x -q 2>&1 | tail -2
