#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
for cfg in "6 0" "9 1" "12 1" "9 2" "12 2" "6 0" "12 1"; do
set -- $cfg
python bench.py --steps 20 --warmup 3 --extras 0 --queue-lanes $1 --host-threads-per-lane $2 > gpurun_out/r2_b14.json 2> gpurun_out/r2_b14.err; echo "qlanes $1 htl $2 rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b14.json')); e=d['e2e']
print('value %.3e e2e %.3e'%(d['value'],e['value']), e['lane_time_share'], e['queue']['passes'])
P
done
