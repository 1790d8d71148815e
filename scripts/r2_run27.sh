#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests/test_gpu_verify.py tests/test_gpu_queue.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
for mc in 1 0; do
python bench.py --steps 20 --warmup 3 --extras 0 --merged-check $mc > gpurun_out/r2_b27_m$mc.json 2> gpurun_out/r2_b27.err; echo "merged $mc bench rc=$?"; tail -2 gpurun_out/r2_b27.err
python scripts/r2_summary.py gpurun_out/r2_b27_m$mc.json 2>&1 | grep "^value\|^pass\|^job" | cut -c1-420
done
