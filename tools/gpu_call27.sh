#!/bin/bash
# round 2, call 27 (2 GPUs): factored SH exchange - kernel test, NCCL test (factored / dense / chunked), N=2 bench A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sh_exchange.py tests/test_parallel.py -q -x > gpurun_out/c27_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c27_pytest.log
tail -12 gpurun_out/c27_pytest.log
for x in factored dense; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 20 --sh-exchange $x --no-e2e > gpurun_out/c27_bench_n2_$x.json 2> gpurun_out/c27_bench_n2_$x.err
  python -c "
import json
b=json.loads(open('gpurun_out/c27_bench_n2_$x.json').read().strip().splitlines()[-1]); print('n2 $x', b['ms_per_step'], b['value'], b.get('grad_check'), b.get('reduce_mode'))" || tail -5 gpurun_out/c27_bench_n2_$x.err
done
for d in 0; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 50 --warmup 20 --sh-degree $d --no-e2e > gpurun_out/c27_bench_n2_deg$d.json 2> gpurun_out/c27_bench_n2_deg$d.err
  python -c "
import json
b=json.loads(open('gpurun_out/c27_bench_n2_deg$d.json').read().strip().splitlines()[-1]); print('n2 deg$d factored', b['ms_per_step'], b.get('grad_check'))" || tail -5 gpurun_out/c27_bench_n2_deg$d.err
done
