#!/bin/bash
# safegcd check: verify tests, then one-job / pass latencies with thread-form and warp-form scalar prep
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests/test_gpu_verify.py tests/test_gpu_queue.py tests/test_gpu_lanes.py -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 20 --warmup 3 --lanes 4 --extras 0 > gpurun_out/r2_b3.json 2> gpurun_out/r2_b3.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_b3.err
python scripts/r2_summary.py gpurun_out/r2_b3.json 2>&1 | grep -v "msm\|dist\|sharded\|extras"
BPP_VPREP_WARP=4096 python bench.py --steps 20 --warmup 3 --lanes 4 --extras 0 > gpurun_out/r2_b3w.json 2> gpurun_out/r2_b3w.err; echo "bench (warp prep <= 4096) rc=$?"; tail -3 gpurun_out/r2_b3w.err
python scripts/r2_summary.py gpurun_out/r2_b3w.json 2>&1 | grep -v "msm\|dist\|sharded\|extras"
