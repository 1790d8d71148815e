#!/bin/bash
# round-2 GPU check: full GPU suite, then bench at a few lane counts
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for L in 3 4 6; do
python bench.py --steps 20 --warmup 3 --lanes $L --extras 0 > gpurun_out/r2_b2_l$L.json 2> gpurun_out/r2_b2_l$L.err; echo "lanes $L bench rc=$?"; tail -3 gpurun_out/r2_b2_l$L.err
python scripts/r2_summary.py gpurun_out/r2_b2_l$L.json 2>&1 | grep -v "msm\|dist\|sharded\|extras"
done
