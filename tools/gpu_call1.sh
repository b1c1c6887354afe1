#!/bin/bash
# GPU call 1 (round 2): tests, parity stats (both variants), bench on the new and the r1 workload
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
nproc >> gpurun_out/c1_gpu.txt; free -g >> gpurun_out/c1_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 900 python tools/parity_stats.py --out gpurun_out/r02_parity_stats.json > gpurun_out/c1_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/c1_parity.log
timeout 600 python bench.py --steps 50 --warmup 20 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; echo "bench rc=$?" >> gpurun_out/c1_bench.err
timeout 300 python bench.py --steps 50 --warmup 20 --workload cfg3_r1scales --no-cpu-baseline --no-e2e > gpurun_out/c1_bench_r1scales.json 2> gpurun_out/c1_bench_r1scales.err
B200GSR_PAIR_MODE=sync timeout 300 python bench.py --steps 50 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/c1_bench_sync.json 2> gpurun_out/c1_bench_sync.err
tail -3 gpurun_out/c1_pytest.log; tail -2 gpurun_out/c1_parity.log; cat gpurun_out/c1_bench.json | cut -c1-600
