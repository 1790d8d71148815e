#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests/test_gpu_prove.py -m gpu -x -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_b32_n2.json 2> gpurun_out/r2_b32_n2.err; echo "n2 rc=$?"; tail -2 gpurun_out/r2_b32_n2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b32_n2.json')); e=d['e2e']
print('value %.4e e2e %.4e'%(d['value'],e['value']), e['lane_time_share'], e['timed_regions_s'], d['host_cores'])
print(d['engine']['per_call_check']['value'], d['engine']['verifier_weights'][:12], d['engine']['lanes_per_gpu'], d['engine']['queue_lanes_per_gpu'])
print(d['one_shot_4096']); print(d['extras'].get('msm_sharded'))
P
