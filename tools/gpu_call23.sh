#!/bin/bash
# round 2, call 23: finer forward work items (2 / 4 CTAs per tile) with two entries in flight
mkdir -p gpurun_out
B200GSR_FWD_VARIANT=42 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -x -q -k "cfg1 or cfg2 or non_square or long_tile or equal_depth or single or empty or work_lists" > gpurun_out/c23_pytest42.log 2>&1; echo "rc=$?" >> gpurun_out/c23_pytest42.log
tail -3 gpurun_out/c23_pytest42.log
for v in 0 42 44; do
  for wl in cfg3_1M_1024 cfg2_100k_512; do
    B200GSR_FWD_VARIANT=$v timeout 300 python bench.py --workload $wl --steps 60 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/c23_bench_${wl}_f$v.json 2> gpurun_out/c23_bench_${wl}_f$v.err
    python -c "
import json
b=json.loads(open('gpurun_out/c23_bench_${wl}_f$v.json').read().strip().splitlines()[-1]); print('fwd$v $wl', round(b['ms_per_step'],4), 'fwd', round(b['stages_ms']['composite_fwd'],4), 'bwd', round(b['stages_ms']['composite_bwd'],4))" || tail -3 gpurun_out/c23_bench_${wl}_f$v.err
  done
done
