#!/bin/bash
# round 2, call 26 (2 GPUs): NCCL tests and the N=2 bench on the final tree
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parallel.py tests/test_gpu_parallel.py -q -x > gpurun_out/c26_pytest_parallel.log 2>&1; echo "rc=$?" >> gpurun_out/c26_pytest_parallel.log
tail -3 gpurun_out/c26_pytest_parallel.log
timeout 600 python -m pytest tests -m gpu -q -x -k "nccl or parallel or sharding" > gpurun_out/c26_pytest_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/c26_pytest_nccl.log
tail -3 gpurun_out/c26_pytest_nccl.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 20 > gpurun_out/c26_bench_n2.json 2> gpurun_out/c26_bench_n2.err
python -c "
import json
b=json.loads(open('gpurun_out/c26_bench_n2.json').read().strip().splitlines()[-1]); print('n2', b['ms_per_step'], b['value'], b.get('grad_check'), b['e2e']['ms_per_step'])" || tail -5 gpurun_out/c26_bench_n2.err
