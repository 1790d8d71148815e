#!/bin/bash
# device weights (one graph per pass) against host weights, lanes sweep
cd ${GRAFT_REPO_ROOT:-.}
for cfg in "0 6" "1 6" "1 4" "1 8" "1 3"; do
set -- $cfg
python bench.py --steps 20 --warmup 3 --extras 0 --device-weights $1 --lanes $2 > gpurun_out/r2_b12_w$1_l$2.json 2> gpurun_out/r2_b12_w$1_l$2.err; echo "dw $1 lanes $2 bench rc=$?"; tail -2 gpurun_out/r2_b12_w$1_l$2.err
python scripts/r2_summary.py gpurun_out/r2_b12_w$1_l$2.json 2>&1 | grep "^value\|^one_shot"
done
python -m pytest tests/test_gpu_queue.py tests/test_gpu_verify.py -m gpu -x -q 2>&1 | tail -3
