#!/bin/bash
# round 2, call 29 (8 GPUs): N=8 weak scaling with the factored SH exchange
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 50 --warmup 20 --no-e2e > gpurun_out/c29_bench_n8_factored.json 2> gpurun_out/c29_bench_n8_factored.err
python -c "
import json
b=json.loads(open('gpurun_out/c29_bench_n8_factored.json').read().strip().splitlines()[-1]); print('n8 factored', b['ms_per_step'], b['value'], b.get('grad_check'), b.get('reduce_mode'))" || tail -5 gpurun_out/c29_bench_n8_factored.err
