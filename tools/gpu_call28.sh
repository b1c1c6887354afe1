#!/bin/bash
# round 2, call 28 (2 GPUs): factored SH exchange tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sh_exchange.py tests/test_parallel.py -q > gpurun_out/c28_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c28_pytest.log
tail -12 gpurun_out/c28_pytest.log
