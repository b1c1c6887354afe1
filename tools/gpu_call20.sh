#!/bin/bash
# round 2, call 20: backward work lists by consumed length (longest first) + forward ILP sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c20_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c20_pytest.log
tail -4 gpurun_out/c20_pytest.log
for cfg in "0 10" "51 10" "53 10" "54 10" "0 33" "0 32"; do
  set -- $cfg
  for wl in cfg3_1M_1024 cfg2_100k_512; do
    B200GSR_FWD_VARIANT=$1 B200GSR_BWD_VARIANT=$2 timeout 300 python bench.py --workload $wl --steps 60 --warmup 20 --no-cpu-baseline --no-e2e > gpurun_out/c20_bench_${wl}_f$1_b$2.json 2> gpurun_out/c20_bench_${wl}_f$1_b$2.err
    python -c "
import json
b=json.loads(open('gpurun_out/c20_bench_${wl}_f$1_b$2.json').read().strip().splitlines()[-1]); print('fwd$1 bwd$2 $wl', round(b['ms_per_step'],4), 'fwd', round(b['stages_ms']['composite_fwd'],4), 'bwd', round(b['stages_ms']['composite_bwd'],4), json.dumps(b['roofline']['load_balance']['bwd']))" || tail -3 gpurun_out/c20_bench_${wl}_f$1_b$2.err
  done
done
