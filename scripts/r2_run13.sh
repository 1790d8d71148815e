#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python bench.py --steps 20 --warmup 3 --extras 0 > gpurun_out/r2_b13.json 2> gpurun_out/r2_b13.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_b13.err
python scripts/r2_summary.py gpurun_out/r2_b13.json 2>&1 | grep "^value\|^e2e"
python -m pytest tests/test_gpu_prove.py -m gpu -x -q 2>&1 | tail -2
for occ in 0 20 24 28; do
echo "== BPP_FB_OCC=$occ"
BPP_FB_OCC=$occ python scripts/prove_lanes_probe.py 8192 8 16 2>&1 | tail -2
BPP_FB_OCC=$occ python scripts/prove_lanes_probe.py 1024 1 2>&1 | tail -1
done
