#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
run() {
"$@" > gpurun_out/r2_b23.json 2> gpurun_out/r2_b23.err; echo "$* rc=$?"; tail -2 gpurun_out/r2_b23.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b23.json')); e=d['e2e']
print('   value %.3e e2e %.3e'%(d['value'],e['value']), e['lane_time_share'], e['timed_regions_s'], d['engine']['verifier_weights'][:6], d['engine']['queue_lanes_per_gpu'])
P
}
run taskset -c 0-3 python bench.py --steps 20 --warmup 3 --extras 0
run env BPP_ADAPTIVE_WAIT=0 taskset -c 0-3 python bench.py --steps 20 --warmup 3 --extras 0
run taskset -c 0-1 python bench.py --steps 20 --warmup 3 --extras 0
run env BPP_ADAPTIVE_WAIT=0 taskset -c 0-1 python bench.py --steps 20 --warmup 3 --extras 0
run python bench.py --steps 20 --warmup 3 --extras 0
run env BPP_ADAPTIVE_WAIT=0 python bench.py --steps 20 --warmup 3 --extras 0
python -m pytest tests/test_gpu_queue.py tests/test_gpu_verify.py -m gpu -x -q 2>&1 | tail -2
