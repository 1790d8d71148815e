#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
for th in 32 64 16 32; do
BPP_VPREP_VEC_THREADS=$th python bench.py --steps 20 --warmup 3 --extras 0 > gpurun_out/r2_b29.json 2> gpurun_out/r2_b29.err; echo "vec threads $th rc=$?"; tail -2 gpurun_out/r2_b29.err
python scripts/r2_summary.py gpurun_out/r2_b29.json 2>&1 | grep "^value\|^pass\|^job" | cut -c1-250
done
python -m pytest tests/test_gpu_verify.py -m gpu -x -q -k "device_replay_sm" 2>&1 | tail -2
