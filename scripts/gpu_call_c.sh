#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_prove.py tests/test_gpu_primitives.py tests/test_gpu_configs.py tests/test_gpu_lanes.py -m gpu -x -q > gpurun_out/t_prove.log 2>&1; tail -3 gpurun_out/t_prove.log
BPP_PROVE_TRACE=1 timeout 200 python scripts/prove_lanes_probe.py 1024 1 2>&1 | tail -4
timeout 200 python scripts/prove_lanes_probe.py 4096 4 2>&1 | tail -2
timeout 200 python scripts/prove_lanes_probe.py 4096 8 2>&1 | tail -2
timeout 200 python scripts/prove_lanes_probe.py 16384 16 2>&1 | tail -2
