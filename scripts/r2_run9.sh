#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests/test_gpu_lanes.py tests/test_gpu_queue.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 --extras 0 > gpurun_out/r2_b9.json 2> gpurun_out/r2_b9.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_b9.err
python scripts/r2_summary.py gpurun_out/r2_b9.json 2>&1 | grep "^value\|^job\|^pass\|^one_shot\|^e2e"
bash scripts/sanitize.sh ${1:-memcheck}
