#!/bin/bash
# round-2 GPU check: queue / boundary / verify tests, then one bench run with a short summary
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests/test_gpu_queue.py tests/test_gpu_verify.py -m gpu -x -q 2>&1 | tail -25
if [ ${PIPESTATUS[0]} -ne 0 ]; then
  echo "=== tests failed with the warp-form scalar prep: retrying with BPP_VPREP_THREAD=1"
  export BPP_VPREP_THREAD=1
  python -m pytest tests/test_gpu_queue.py tests/test_gpu_verify.py -m gpu -x -q 2>&1 | tail -15
fi
python bench.py --steps 20 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.err; echo bench rc=$?; tail -5 gpurun_out/r2_b1.err
python scripts/r2_summary.py gpurun_out/r2_b1.json
