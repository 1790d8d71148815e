#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_prove.py tests/test_gpu_primitives.py -m gpu -x -q > gpurun_out/t_prove.log 2>&1; tail -2 gpurun_out/t_prove.log
for cfg in "BPP_FB_NOPREFETCH=1" "BPP_FB_NOPREFETCH=0" "BPP_FB_NOPREFETCH=1 BPP_FB_C=8" "BPP_FB_NOPREFETCH=0 BPP_FB_C=8" "BPP_FB_NOPREFETCH=0 BPP_FB_C=10"; do
  echo "== $cfg"
  env $cfg BPP_PROVE_TRACE=1 timeout 100 python scripts/prove_lanes_probe.py 1024 1 2>&1 | tail -2
  env $cfg timeout 100 python scripts/prove_lanes_probe.py 8192 8 2>&1 | tail -1
done
