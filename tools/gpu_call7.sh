#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c7_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/c7_pytest_all.log
for m in 8 6; do
  B200GSR_PBWD_MINB=$m timeout 300 python bench.py --steps 60 --warmup 20 --no-e2e --no-cpu-baseline > gpurun_out/c7_bench_minb$m.json 2> gpurun_out/c7_bench_minb$m.err
done
tail -3 gpurun_out/c7_pytest_all.log
for m in 8 6; do python -c "
import json
b=json.loads(open('gpurun_out/c7_bench_minb$m.json').read().strip().splitlines()[-1]); print('minb$m', round(b['ms_per_step'],4), b['ms_per_step_spread']['median'], {k:round(v,4) for k,v in b['stages_ms'].items()})"; done
