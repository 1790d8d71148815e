#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/nap_probe.log
for nap in 60 120 250; do
  echo "== BPP_NAP_US=$nap cpus=0-3" >> gpurun_out/nap_probe.log
  BPP_NAP_US=$nap PROBE_ONE_MODE=1 timeout 200 taskset -c 0-3 python scripts/pipeline_probe.py 64 32 >> gpurun_out/nap_probe.log 2>&1
done
echo "== BPP_NAP_US=120 all cpus" >> gpurun_out/nap_probe.log
BPP_NAP_US=120 PROBE_ONE_MODE=1 timeout 200 python scripts/pipeline_probe.py 64 32 >> gpurun_out/nap_probe.log 2>&1
cat gpurun_out/nap_probe.log
