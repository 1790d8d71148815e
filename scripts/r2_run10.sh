#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
for H in 2 5 8; do
python bench.py --steps 20 --warmup 3 --extras 0 --host-threads-per-lane $H > gpurun_out/r2_b10_h$H.json 2> gpurun_out/r2_b10_h$H.err; echo "htl $H bench rc=$?"; tail -2 gpurun_out/r2_b10_h$H.err
python scripts/r2_summary.py gpurun_out/r2_b10_h$H.json 2>&1 | grep "^value\|^one_shot"
done
