#!/bin/bash
# final validation of the round: smoke, full GPU suite, the driver's bench command, the reference arm
cd ${GRAFT_REPO_ROOT:-.}
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b31_n1.json 2> gpurun_out/r2_b31_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_b31_n1.err
python scripts/r2_summary.py gpurun_out/r2_b31_n1.json 2>&1 | cut -c1-420
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_b31_ref.json 2> gpurun_out/r2_b31_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2_b31_ref.json
