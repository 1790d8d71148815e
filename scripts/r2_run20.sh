#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests/test_gpu_lanes.py tests/test_gpu_verify.py tests/test_gpu_queue.py -m gpu -x -q 2>&1 | tail -3
for cfg in "1 6 12" "1 8 12" "1 8 16" "0 6 12"; do
set -- $cfg
python bench.py --steps 20 --warmup 3 --extras 0 --device-weights $1 --lanes $2 --queue-lanes $3 > gpurun_out/r2_b20.json 2> gpurun_out/r2_b20.err; echo "dw $1 lanes $2 qlanes $3 rc=$?"; tail -2 gpurun_out/r2_b20.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b20.json')); e=d['e2e']
print('value %.3e e2e %.3e'%(d['value'],e['value']), e['lane_time_share'], d['roofline']['one_pass_alone']['ms'])
P
done
taskset -c 0-3 python bench.py --steps 20 --warmup 3 --extras 0 --device-weights 1 --lanes 8 --queue-lanes 8 --host-threads-per-lane 1 > gpurun_out/r2_b20c.json 2> gpurun_out/r2_b20c.err; echo "4 cores dw rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b20c.json')); e=d['e2e']
print('4 cores dw: value %.3e e2e %.3e'%(d['value'],e['value']), e['lane_time_share'])
P
