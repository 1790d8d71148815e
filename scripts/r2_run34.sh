#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
for nap in 60 200 500 200; do
BPP_NAP_US=$nap taskset -c 0-3 python bench.py --steps 20 --warmup 3 --extras 0 > gpurun_out/r2_b34.json 2> gpurun_out/r2_b34.err; echo "nap $nap rc=$?"; tail -2 gpurun_out/r2_b34.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b34.json')); e=d['e2e']
print('   value %.3e e2e %.3e'%(d['value'],e['value']), e['lane_time_share'], e['timed_regions_s'], d['engine']['verifier_weights'][:6], d['host_cores'])
P
done
