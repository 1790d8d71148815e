#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests/test_gpu_primitives.py tests/test_gpu_lanes.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 --extras 0 > gpurun_out/r2_b24.json 2> gpurun_out/r2_b24.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_b24.err
python scripts/r2_summary.py gpurun_out/r2_b24.json 2>&1 | grep "^value\|^pass\|^job"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b24.json')); e=d['e2e']
print('   value %.3e e2e %.3e'%(d['value'],e['value']), e['lane_time_share'], e['timed_regions_s'])
P
bash scripts/sanitize.sh host 2>&1 | tail -15
