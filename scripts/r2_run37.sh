#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b37_n1.json 2> gpurun_out/r2_b37_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_b37_n1.err
python scripts/r2_summary.py gpurun_out/r2_b37_n1.json 2>&1 | grep "^value\|^one_shot\|^extras\|^cpu" | cut -c1-300
