#!/bin/bash
# one GPU call: full GPU suite, bucket-kernel probe, adaptive-wait A/B with 4 and all host cores
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core" 
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -3 gpurun_out/t.log
timeout 200 python scripts/msm_bucket_probe.py 13 14 15 16 17 18 > gpurun_out/msm_probe.log 2>&1; tail -5 gpurun_out/msm_probe.log
for aw in 0 1; do
  for cpus in 0-3 0-15; do
    echo "== adaptive_wait=$aw cpus=$cpus" >> gpurun_out/wait_probe.log
    BPP_ADAPTIVE_WAIT=$aw PROBE_ONE_MODE=1 timeout 200 taskset -c $cpus python scripts/pipeline_probe.py 48 32 >> gpurun_out/wait_probe.log 2>&1
  done
done
cat gpurun_out/wait_probe.log
