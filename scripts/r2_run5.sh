#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests/test_gpu_primitives.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 --lanes 4 --prove-batch 8192 --prove-lanes 8 > gpurun_out/r2_b5.json 2> gpurun_out/r2_b5.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_b5.err
python scripts/r2_summary.py gpurun_out/r2_b5.json 2>&1 | grep -v "^e2e\|^clocks\|^cpu"
