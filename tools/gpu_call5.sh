#!/bin/bash
# GPU call 5: f2/f4/disparity tests, full suite, sanitizer logs, extra bench artefacts, ncu of the shipping kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scene.py tests/test_gpu_postprocess.py tests/test_gpu_densify.py -m gpu -q > gpurun_out/c5_pytest_new.log 2>&1; echo "rc=$?" >> gpurun_out/c5_pytest_new.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c5_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/c5_pytest_all.log
# compute-sanitizer: cfg1 fwd+bwd (smoke-sized) and the kNN, memcheck + racecheck
cat > /tmp/san.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
from tests import util_scene as U
from tests import parity_tools as PT
from simple_knn._C import distCUDA2
sc, cam, deg = U.make_inputs(10000, 256, 256)
g = torch.Generator().manual_seed(0)
gc, gd = torch.randn(3, 256, 256, generator=g), torch.randn(2, 256, 256, generator=g)
r = PT.cuda_forward_backward(sc, cam, deg, gc, gd, score=False)
d = distCUDA2(sc["means3D"].cuda())
torch.cuda.synchronize()
print("sanitized run ok", float(r["color"].sum()), float(d.sum()))
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san.py > gpurun_out/c5_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/c5_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/san.py > gpurun_out/c5_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/c5_racecheck.log
# bench artefacts for the other configs
for w in cfg2_100k_512 cfg2b_81920_512 cfg3b_1M_1024_screenfill cfg1_10k_256; do
  timeout 300 python bench.py --steps 50 --warmup 20 --no-e2e --no-cpu-baseline --workload $w > gpurun_out/c5_bench_$w.json 2> gpurun_out/c5_bench_$w.err
done
# kNN timings
python - > gpurun_out/c5_knn.json 2> gpurun_out/c5_knn.err <<'PY'
import torch, json, sys
sys.path.insert(0, '.')
from simple_knn._C import distCUDA2
out = {}
for n in (1_000_000, 2_300_000):
    p = torch.rand(n, 3, device="cuda")
    for _ in range(3): distCUDA2(p)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): distCUDA2(p)
    e1.record(); torch.cuda.synchronize()
    out[str(n)] = e0.elapsed_time(e1) / 20
print(json.dumps({"distCUDA2_ms": out}))
PY
# ncu: launch list of one bench step + full capture of the dominant kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 4 --warmup 20 --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/c5_ncu_list.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"composite_bwd2|composite_fwd_kernel|project_bwd|project_sh" -s 40 -c 8 -o gpurun_out/prof_r02 python bench.py --steps 3 --warmup 10 --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/c5_ncu_full.err
timeout 600 python bench.py --steps 50 --warmup 20 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
tail -3 gpurun_out/c5_pytest_new.log; tail -3 gpurun_out/c5_pytest_all.log; tail -3 gpurun_out/c5_memcheck.log; tail -3 gpurun_out/c5_racecheck.log; cat gpurun_out/c5_knn.json; ls -la gpurun_out/prof_r02*
