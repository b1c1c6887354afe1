#!/bin/bash
# round 2, call 25: validation of the final tree - full suite, smoke, sanitizer, benches of every config, cfg5, ncu evidence
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c25_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/c25_pytest_all.log
tail -3 gpurun_out/c25_pytest_all.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/c25_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c25_smoke.log; tail -2 gpurun_out/c25_smoke.log
cat > /tmp/san.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from tests import util_scene as U
from tests import parity_tools as PT
sc, cam, deg = U.make_inputs(10000, 256, 256)
g = torch.Generator().manual_seed(0)
gc, gd = torch.randn(3, 256, 256, generator=g), torch.randn(2, 256, 256, generator=g)
r = PT.cuda_forward_backward(sc, cam, deg, gc, gd, score=False)
torch.cuda.synchronize()
print("sanitized run ok", float(r["color"].sum()))
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/san.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r02_sanitizer_racecheck.log
tail -3 gpurun_out/r02_sanitizer_memcheck.log; tail -3 gpurun_out/r02_sanitizer_racecheck.log
timeout 600 python bench.py --steps 50 --warmup 20 > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c25_bench_ref.json 2> gpurun_out/c25_bench_ref.err
for w in cfg2_100k_512 cfg2b_81920_512 cfg3b_1M_1024_screenfill cfg1_10k_256 cfg3_r1scales; do
  timeout 300 python bench.py --steps 50 --warmup 20 --no-e2e --no-cpu-baseline --workload $w > gpurun_out/c25_bench_$w.json 2> gpurun_out/c25_bench_$w.err
done
for g in torch fused_rng views; do
  timeout 300 python benchmarks/scene_step.py --steps 15 --warmup 5 --glue $g > gpurun_out/c25_scene_$g.json 2> gpurun_out/c25_scene_$g.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 40 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 4 --warmup 20 --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/c25_ncu_list.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"composite_bwd2|composite_fwd_kernel|project_bwd|project_sh|multisplit|sort_big|sort_small|scan_order" -s 60 -c 10 -o gpurun_out/prof_r02f python bench.py --steps 3 --warmup 10 --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/c25_ncu_full.err
python -c "
import json
b=json.loads(open('gpurun_out/c25_bench.json').read().strip().splitlines()[-1]); print(b['ms_per_step'], b.get('ms_per_step_spread'), b['stages_ms'], b['e2e']['ms_per_step'], b['cpu_baseline']['value'], b['roofline']['frac'])
r=json.loads(open('gpurun_out/c25_bench_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['ms_per_step'])"
for w in cfg2_100k_512 cfg2b_81920_512 cfg3b_1M_1024_screenfill cfg1_10k_256 cfg3_r1scales; do python -c "
import json
b=json.loads(open('gpurun_out/c25_bench_$w.json').read().strip().splitlines()[-1]); print('$w', round(b['ms_per_step'],4), {k: round(v,4) for k,v in b['stages_ms'].items()})"; done
cat gpurun_out/c25_scene_*.json
