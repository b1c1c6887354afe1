#!/bin/bash
# GPU call 10: full suite on the final kernels, cfg5 in all modes, parity stats refresh, final bench + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c10_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/c10_pytest_all.log
B200GSR_BWD_VARIANT=20 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -x -q -k "backward or cfg1 or sh_degrees or precomputed or zero_scales or screen_filling or odd_point or retain" > gpurun_out/c10_pytest_v20.log 2>&1; echo "rc=$?" >> gpurun_out/c10_pytest_v20.log
for v in 10 20 23; do
  B200GSR_BWD_VARIANT=$v timeout 300 python bench.py --steps 60 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/c10_bench_v$v.json 2> gpurun_out/c10_bench_v$v.err
  python -c "
import json
b=json.loads(open('gpurun_out/c10_bench_v$v.json').read().strip().splitlines()[-1]); print('v$v', round(b['ms_per_step'],4), b['stages_ms']['composite_bwd'])"
done
tail -2 gpurun_out/c10_pytest_v20.log
for g in torch fused_rng views; do
  timeout 300 python benchmarks/scene_step.py --steps 15 --warmup 5 --glue $g > gpurun_out/c10_scene_$g.json 2> gpurun_out/c10_scene_$g.err
done
timeout 1200 python tools/parity_stats.py --out gpurun_out/r02_parity_stats.json > gpurun_out/c10_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/c10_parity.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 40 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 4 --warmup 20 --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/c10_ncu_list.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"composite_bwd2|composite_fwd_kernel|project_bwd|project_sh|multisplit|sort_big|sort_small|scan_order" -s 60 -c 10 -o gpurun_out/prof_r02b python bench.py --steps 3 --warmup 10 --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/c10_ncu_full.err
timeout 600 python bench.py --steps 50 --warmup 20 > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c10_bench_ref.json 2> gpurun_out/c10_bench_ref.err
tail -3 gpurun_out/c10_pytest_all.log; cat gpurun_out/c10_scene_*.json; tail -2 gpurun_out/c10_parity.log
python -c "
import json
b=json.loads(open('gpurun_out/c10_bench.json').read().strip().splitlines()[-1]); print(b['ms_per_step'], b['ms_per_step_spread'], b['stages_ms'], b['e2e']['ms_per_step'], b['cpu_baseline']['value'])
r=json.loads(open('gpurun_out/c10_bench_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['ms_per_step'], r['cpu_baseline'].get('spread'))"
