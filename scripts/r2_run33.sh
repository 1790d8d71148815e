#!/bin/bash
# N=8 torchrun bench with per-rank core slices
cd ${GRAFT_REPO_ROOT:-.}
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --extras 0 > gpurun_out/r2_b33_n8.json 2> gpurun_out/r2_b33_n8.err; echo "n8 rc=$?"; tail -3 gpurun_out/r2_b33_n8.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b33_n8.json')); e=d['e2e']
print('value %.4e e2e %.4e'%(d['value'],e['value']), e['lane_time_share'], e['timed_regions_s'])
print(d['engine']['host_cores_of_this_rank'], d['cpu_baseline']['value'], d['one_shot_4096']['ms_end_to_end'])
P
