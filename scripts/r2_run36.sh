#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
for cfg in "16 2" "12 2" "16 1" "20 1"; do
set -- $cfg
python bench.py --steps 20 --warmup 3 --extras 0 --queue-lanes $1 --host-threads-per-lane $2 > gpurun_out/r2_b36.json 2> gpurun_out/r2_b36.err; echo "qlanes $1 htl $2 rc=$?"; tail -2 gpurun_out/r2_b36.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b36.json')); e=d['e2e']
print('   value %.3e e2e %.3e'%(d['value'],e['value']), e['lane_time_share'], e['timed_regions_s'])
P
done
