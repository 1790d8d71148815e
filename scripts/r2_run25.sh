#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests/test_gpu_verify.py tests/test_gpu_queue.py tests/test_gpu_lanes.py -m gpu -x -q 2>&1 | tail -8
for mc in 1 0 1; do
python bench.py --steps 20 --warmup 3 --extras 0 --merged-check $mc > gpurun_out/r2_b25_m$mc.json 2> gpurun_out/r2_b25.err; echo "merged $mc bench rc=$?"; tail -2 gpurun_out/r2_b25.err
python scripts/r2_summary.py gpurun_out/r2_b25_m$mc.json 2>&1 | grep "^value\|^pass\|^job\|^one_shot" | cut -c1-420
python - <<P
import json
d=json.load(open('gpurun_out/r2_b25_m$mc.json')); e=d['e2e']
print('   value %.3e e2e %.3e'%(d['value'],e['value']), e['lane_time_share'], e['timed_regions_s'])
P
done
python bench.py --steps 20 --warmup 3 --extras 0 --merged-check 1 --device-weights 1 > gpurun_out/r2_b25_m1dw.json 2> gpurun_out/r2_b25.err; echo "merged+dw rc=$?"
python - <<P
import json
d=json.load(open('gpurun_out/r2_b25_m1dw.json')); e=d['e2e']
print('   value %.3e e2e %.3e'%(d['value'],e['value']), e['lane_time_share'], e['timed_regions_s'])
P
