#!/bin/bash
# round 2, call 15: balanced multisplit grid + row-bounded bucket sort: parity, A/B of the scatter batch, ncu of the binning kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py tests/test_gpu_multiview.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/c15_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c15_pytest.log
tail -3 gpurun_out/c15_pytest.log
for sb in 4 1; do
for wl in cfg3_1M_1024 cfg3b_1M_1024_screenfill cfg2_100k_512; do
  B200GSR_SCATTER_BATCH=$sb timeout 300 python bench.py --workload $wl --steps 60 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/c15_bench_${wl}_b$sb.json 2> gpurun_out/c15_bench_${wl}_b$sb.err
  python -c "
import json
b=json.loads(open('gpurun_out/c15_bench_${wl}_b$sb.json').read().strip().splitlines()[-1]); print('batch$sb $wl', round(b['ms_per_step'],4), {k: round(v,4) for k,v in b['stages_ms'].items()})"
done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"multisplit|sort_big|sort_small" -s 40 -c 4 -o gpurun_out/prof_r02c python bench.py --steps 3 --warmup 10 --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/c15_ncu_full.err
ls -la gpurun_out/prof_r02c.ncu-rep
