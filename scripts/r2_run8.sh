#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests/test_gpu_lanes.py tests/test_gpu_queue.py tests/test_gpu_primitives.py -m gpu -x -q 2>&1 | tail -4
for cfg in "4 16" "6 16" "8 16" "4 32" "3 32"; do
set -- $cfg
python bench.py --steps 20 --warmup 3 --lanes $1 --pass-jobs $2 --extras 0 > gpurun_out/r2_b8_l$1_k$2.json 2> gpurun_out/r2_b8_l$1_k$2.err; echo "lanes $1 jobs/pass $2 bench rc=$?"; tail -2 gpurun_out/r2_b8_l$1_k$2.err
python scripts/r2_summary.py gpurun_out/r2_b8_l$1_k$2.json 2>&1 | grep "^value\|^job\|^pass\|^one_shot"
done
