#!/bin/bash
# round 2, call 19: two list entries in flight per warp (forward variant 5, backward variants 30/33/32): parity + A/B
mkdir -p gpurun_out
B200GSR_FWD_VARIANT=5 B200GSR_BWD_VARIANT=33 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py tests/test_gpu_multiview.py -m gpu -x -q > gpurun_out/c19_pytest_f5_b33.log 2>&1; echo "rc=$?" >> gpurun_out/c19_pytest_f5_b33.log
tail -3 gpurun_out/c19_pytest_f5_b33.log
B200GSR_BWD_VARIANT=30 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -x -q > gpurun_out/c19_pytest_b30.log 2>&1; echo "rc=$?" >> gpurun_out/c19_pytest_b30.log
tail -2 gpurun_out/c19_pytest_b30.log
for cfg in "0 10" "5 10" "0 30" "0 33" "0 32" "5 33"; do
  set -- $cfg
  for wl in cfg3_1M_1024 cfg2_100k_512; do
    B200GSR_FWD_VARIANT=$1 B200GSR_BWD_VARIANT=$2 timeout 300 python bench.py --workload $wl --steps 60 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/c19_bench_${wl}_f$1_b$2.json 2> gpurun_out/c19_bench_${wl}_f$1_b$2.err
    python -c "
import json
b=json.loads(open('gpurun_out/c19_bench_${wl}_f$1_b$2.json').read().strip().splitlines()[-1]); print('fwd$1 bwd$2 $wl', round(b['ms_per_step'],4), 'fwd', round(b['stages_ms']['composite_fwd'],4), 'bwd', round(b['stages_ms']['composite_bwd'],4))"
  done
done
