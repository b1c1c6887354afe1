#!/bin/bash
# round 2, call 22: segmented backward with a worst-case checkpoint pool (experiment)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -x -q > gpurun_out/c22_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c22_pytest.log
tail -4 gpurun_out/c22_pytest.log
for wl in cfg3_1M_1024 cfg2_100k_512; do
    timeout 300 python bench.py --workload $wl --steps 60 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/c22_bench_${wl}.json 2> gpurun_out/c22_bench_${wl}.err
    python -c "
import json
b=json.loads(open('gpurun_out/c22_bench_${wl}.json').read().strip().splitlines()[-1]); print('$wl', round(b['ms_per_step'],4), {k: round(v,4) for k,v in b['stages_ms'].items()}, json.dumps(b['roofline']['load_balance']))" || tail -3 gpurun_out/c22_bench_${wl}.err
done
