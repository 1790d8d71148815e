#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
run() {
python bench.py --steps 20 --warmup 3 --extras 0 "$@" > gpurun_out/r2_b21.json 2> gpurun_out/r2_b21.err; echo "$* rc=$?"; tail -2 gpurun_out/r2_b21.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b21.json')); e=d['e2e']
print('   value %.3e e2e %.3e'%(d['value'],e['value']), e['lane_time_share'])
P
}
for i in 1 2 3; do
run --device-weights 1 --lanes 8 --queue-lanes 8 --host-threads-per-lane 1
run --device-weights 1 --lanes 8 --queue-lanes 12 --host-threads-per-lane 1
run --device-weights 0 --lanes 6 --queue-lanes 12
done
