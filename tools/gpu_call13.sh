#!/bin/bash
mkdir -p gpurun_out
for g in torch fused_rng views; do
  timeout 300 python benchmarks/scene_step.py --steps 15 --warmup 5 --glue $g > gpurun_out/c13_scene_$g.json 2> gpurun_out/c13_scene_$g.err
done
B200GSR_FWD_VARIANT=4 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py tests/test_gpu_multiview.py -m gpu -x -q -k "cfg1 or cfg2 or non_square or long_tile or equal_depth or single_gaussian or empty or views" > gpurun_out/c13_pytest_fwd4.log 2>&1; echo "rc=$?" >> gpurun_out/c13_pytest_fwd4.log
for v in 0 4; do
  B200GSR_FWD_VARIANT=$v timeout 200 python bench.py --steps 60 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/c13_bench_fwd$v.json 2> gpurun_out/c13_bench_fwd$v.err
  python -c "
import json
b=json.loads(open('gpurun_out/c13_bench_fwd$v.json').read().strip().splitlines()[-1]); print('fwd$v', round(b['ms_per_step'],4), b['stages_ms']['composite_fwd'])"
done
tail -2 gpurun_out/c13_pytest_fwd4.log; cat gpurun_out/c13_scene_*.json
