#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b16_n1.json 2> gpurun_out/r2_b16_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_b16_n1.err
python scripts/r2_summary.py gpurun_out/r2_b16_n1.json 2>&1 | cut -c1-600
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_b16_ref.json 2> gpurun_out/r2_b16_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r2_b16_ref.json
