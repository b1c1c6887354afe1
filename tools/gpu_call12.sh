#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/c12_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c12_smoke.log
for g in torch fused_rng views; do
  timeout 300 python benchmarks/scene_step.py --steps 15 --warmup 5 --glue $g > gpurun_out/c12_scene_$g.json 2> gpurun_out/c12_scene_$g.err
done
tail -2 gpurun_out/c12_smoke.log; cat gpurun_out/c12_scene_*.json
