#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests/test_gpu_queue.py -m gpu -x -q 2>&1 | tail -3
for pg in 0 1 0; do
python bench.py --steps 20 --warmup 3 --extras 0 --pageable-inputs $pg > gpurun_out/r2_b19.json 2> gpurun_out/r2_b19.err; echo "pageable $pg rc=$?"; tail -2 gpurun_out/r2_b19.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b19.json')); e=d['e2e']
print('value %.3e e2e %.3e'%(d['value'],e['value']), e['lane_time_share'], e['queue']['passes'], e['host_ms_per_pass_of_16_jobs'])
P
done
# 4 host cores only (what a rank of the 8-GPU box has)
taskset -c 0-3 python bench.py --steps 20 --warmup 3 --extras 0 --queue-lanes 6 > gpurun_out/r2_b19c.json 2> gpurun_out/r2_b19c.err; echo "4 cores rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b19c.json')); e=d['e2e']
print('4 cores: value %.3e e2e %.3e'%(d['value'],e['value']), e['lane_time_share'], d['engine']['queue_lanes_per_gpu'], d['engine']['queue_host_threads_per_lane'])
P
