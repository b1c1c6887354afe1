#!/bin/bash
# round 2, call 18: load-balance instrumentation of the persistent composite kernels
mkdir -p gpurun_out
for wl in cfg3_1M_1024 cfg3b_1M_1024_screenfill cfg2_100k_512; do
  timeout 300 python bench.py --workload $wl --steps 30 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/c18_bench_${wl}.json 2> gpurun_out/c18_bench_${wl}.err
  python -c "
import json
b=json.loads(open('gpurun_out/c18_bench_${wl}.json').read().strip().splitlines()[-1]); print('$wl', round(b['ms_per_step'],4), json.dumps(b['roofline']['load_balance']))" || tail -5 gpurun_out/c18_bench_${wl}.err
done
