#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --extras 0 > gpurun_out/r2_b35.json 2> gpurun_out/r2_b35.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_b35.err
python scripts/r2_summary.py gpurun_out/r2_b35.json 2>&1 | grep "^value"
