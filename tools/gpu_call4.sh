#!/bin/bash
# GPU call 4: new-op tests (scene assembly, disparity), forward staging A/B (LDGSTS vs UBLKCP keys vs +gather4), cfg5
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scene.py tests/test_gpu_postprocess.py -m gpu -x -q > gpurun_out/c4_pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/c4_pytest_new.log
for v in 1 2; do
  B200GSR_FWD_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -x -q -k "cfg1 or cfg2 or non_square or long_tile or equal_depth or single_gaussian or empty" > gpurun_out/c4_pytest_fwd$v.log 2>&1; echo "rc=$?" >> gpurun_out/c4_pytest_fwd$v.log
done
for v in 0 1 2; do
  B200GSR_FWD_VARIANT=$v timeout 200 python bench.py --steps 60 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/c4_bench_fwd$v.json 2> gpurun_out/c4_bench_fwd$v.err
done
for g in torch fused fused_rng; do
  timeout 300 python benchmarks/scene_step.py --steps 10 --warmup 3 --glue $g > gpurun_out/c4_scene_$g.json 2> gpurun_out/c4_scene_$g.err
done
# instruction counts of the three forward variants (one ncu pass each, composite_fwd kernels only)
for v in 0 1 2; do
  B200GSR_FWD_VARIANT=$v timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:composite_fwd -s 3 -c 2 --csv --log-file gpurun_out/c4_ncu_fwd$v.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2> gpurun_out/c4_ncu_fwd$v.err
done
tail -3 gpurun_out/c4_pytest_new.log; tail -2 gpurun_out/c4_pytest_fwd1.log; tail -2 gpurun_out/c4_pytest_fwd2.log
for v in 0 1 2; do python -c "
import json
b=json.loads(open('gpurun_out/c4_bench_fwd$v.json').read().strip().splitlines()[-1]); print('fwd$v', round(b['ms_per_step'],4), b['stages_ms']['composite_fwd'])"; done
cat gpurun_out/c4_scene_*.json
