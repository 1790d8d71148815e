#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
for th in 200000 100000 200000 100000; do
BPP_MSM_OCC3_MIN=$th python bench.py --steps 20 --warmup 3 --extras 0 > gpurun_out/r2_b28.json 2> gpurun_out/r2_b28.err; echo "occ3_min $th rc=$?"; tail -2 gpurun_out/r2_b28.err
python scripts/r2_summary.py gpurun_out/r2_b28.json 2>&1 | grep "^value\|^pass" | cut -c1-330
done
