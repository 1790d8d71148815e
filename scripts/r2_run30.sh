#!/bin/bash
# N=8 torchrun bench (the driver's command shape), final defaults
cd ${GRAFT_REPO_ROOT:-.}
nvidia-smi -L | wc -l; nproc
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_b30_n8.json 2> gpurun_out/r2_b30_n8.err; echo "n8 rc=$?"; tail -3 gpurun_out/r2_b30_n8.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b30_n8.json')); e=d['e2e']
print('value %.4e e2e %.4e'%(d['value'],e['value']), e['lane_time_share'], e['timed_regions_s'])
print(d['engine']['per_call_check'], d['engine']['verifier_weights'][:12], d['engine']['lanes_per_gpu'], d['engine']['queue_lanes_per_gpu'])
print(d['one_shot_4096']); print(d['cpu_baseline'])
P
