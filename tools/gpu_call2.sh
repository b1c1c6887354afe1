#!/bin/bash
# GPU call 2: full tests, backward-variant correctness + A/B timing, parity stats, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
for v in 10 12; do
  B200GSR_BWD_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -x -q -k "backward or cfg1 or sh_degrees or precomputed or zero_scales or screen_filling or odd_point or retain" > gpurun_out/c2_pytest_v$v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest_v$v.log
done
for v in 0 10 11 12 14 125; do
  B200GSR_BWD_VARIANT=$v timeout 300 python bench.py --steps 60 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/c2_bench_v$v.json 2> gpurun_out/c2_bench_v$v.err
done
timeout 1200 python tools/parity_stats.py --out gpurun_out/r02_parity_stats.json > gpurun_out/c2_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/c2_parity.log
timeout 600 python bench.py --steps 50 --warmup 20 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; echo "bench rc=$?" >> gpurun_out/c2_bench.err
tail -3 gpurun_out/c2_pytest.log; tail -2 gpurun_out/c2_pytest_v10.log; tail -2 gpurun_out/c2_pytest_v12.log; tail -2 gpurun_out/c2_parity.log
for v in 0 10 11 12 14 125; do python -c "
import json,sys
b=json.loads(open('gpurun_out/c2_bench_v$v.json').read().strip().splitlines()[-1])
print('v$v', round(b['ms_per_step'],4), b['ms_per_step_spread']['median'], b['stages_ms']['composite_bwd'])
"; done
