#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b26_n1.json 2> gpurun_out/r2_b26_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_b26_n1.err
python scripts/r2_summary.py gpurun_out/r2_b26_n1.json 2>&1 | cut -c1-500
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_b26_ref.json 2> gpurun_out/r2_b26_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_b26_ref.json
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
