#!/bin/bash
# heavy buckets + occupancy experiment
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests/test_gpu_primitives.py tests/test_gpu_lanes.py -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 20 --warmup 3 --lanes 4 --prove-batch 2048 --prove-lanes 2 > gpurun_out/r2_b4.json 2> gpurun_out/r2_b4.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_b4.err
python scripts/r2_summary.py gpurun_out/r2_b4.json 2>&1 | grep -v "^e2e\|^clocks\|^cpu"
BPP_TUNE_OCC=1 python bench.py --steps 20 --warmup 3 --lanes 4 --extras 0 > gpurun_out/r2_b4o.json 2> gpurun_out/r2_b4o.err; echo "bench (BPP_TUNE_OCC=1) rc=$?"; tail -3 gpurun_out/r2_b4o.err
python scripts/r2_summary.py gpurun_out/r2_b4o.json 2>&1 | grep -v "^e2e\|^clocks\|^cpu\|msm\|dist\|sharded\|extras"
