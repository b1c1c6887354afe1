#!/bin/bash
# round 2, call 24: forward with batched cull rounds
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py tests/test_gpu_multiview.py -m gpu -x -q > gpurun_out/c24_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c24_pytest.log
tail -3 gpurun_out/c24_pytest.log
for wl in cfg3_1M_1024 cfg2_100k_512 cfg3b_1M_1024_screenfill; do
    timeout 300 python bench.py --workload $wl --steps 60 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/c24_bench_${wl}.json 2> gpurun_out/c24_bench_${wl}.err
    python -c "
import json
b=json.loads(open('gpurun_out/c24_bench_${wl}.json').read().strip().splitlines()[-1]); print('$wl', round(b['ms_per_step'],4), 'fwd', round(b['stages_ms']['composite_fwd'],4), 'bwd', round(b['stages_ms']['composite_bwd'],4), json.dumps(b['roofline']['load_balance']['fwd']))" || tail -3 gpurun_out/c24_bench_${wl}.err
done
