#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_lanes.py -m gpu -x -q > gpurun_out/t_msm.log 2>&1; tail -3 gpurun_out/t_msm.log
for parts in 1 0; do
  echo "== BPP_MSM_REDUCE_PARTS=$parts (0 = automatic)"
  if [ $parts = 1 ]; then export BPP_MSM_REDUCE_PARTS=1; else unset BPP_MSM_REDUCE_PARTS; fi
  timeout 200 python scripts/msm_sweep.py 16 18 20 22 2>&1 | tail -24
done
