#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python - <<'P'
import sys; sys.path.insert(0,'tests')
import bpp
eng=bpp.engine()
for w in (0,1,2,12):
    print('microbench', w, '%.4e'%eng.microbench(w,2000)[0])
P
for cfg in "6 0 1" "6 0 2" "8 2 2" "12 2 2" "6 0 2" "8 2 2" "12 2 2" "12 2 4"; do
set -- $cfg
python bench.py --steps 20 --warmup 3 --extras 0 --queue-lanes $1 --host-threads-per-lane $2 --submitters $3 > gpurun_out/r2_b15.json 2> gpurun_out/r2_b15.err; echo "qlanes $1 htl $2 sub $3 rc=$?"; tail -2 gpurun_out/r2_b15.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_b15.json')); e=d['e2e']
print('value %.3e e2e %.3e'%(d['value'],e['value']), e['lane_time_share'], e['queue']['passes'], d['wall_s_timed_region'])
P
done
