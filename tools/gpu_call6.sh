#!/bin/bash
# GPU call 6: multi-view (f1) tests + full suite, cfg5 in all modes, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multiview.py -m gpu -x -q > gpurun_out/c6_pytest_mv.log 2>&1; echo "rc=$?" >> gpurun_out/c6_pytest_mv.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c6_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/c6_pytest_all.log
for g in torch fused_rng views; do
  timeout 300 python benchmarks/scene_step.py --steps 15 --warmup 5 --glue $g > gpurun_out/c6_scene_$g.json 2> gpurun_out/c6_scene_$g.err
done
timeout 300 python bench.py --steps 50 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
tail -5 gpurun_out/c6_pytest_mv.log; tail -3 gpurun_out/c6_pytest_all.log; cat gpurun_out/c6_scene_*.json; tail -2 gpurun_out/c6_scene_views.err
python -c "
import json
b=json.loads(open('gpurun_out/c6_bench.json').read().strip().splitlines()[-1]); print(b['ms_per_step'], b['ms_per_step_spread'], b['stages_ms'])"
