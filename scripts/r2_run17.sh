#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests/test_gpu_primitives.py tests/test_gpu_verify.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --extras 0 > gpurun_out/r2_b17.json 2> gpurun_out/r2_b17.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_b17.err
python scripts/r2_summary.py gpurun_out/r2_b17.json 2>&1 | grep "^value\|^pass\|^job"
python - <<'P'
import sys; sys.path.insert(0,'tests')
import bpp
eng=bpp.engine()
for w in (2,4,5,6,7,8,11,12):
    print('microbench', w, '%.4e'%eng.microbench(w,2000)[0])
P
