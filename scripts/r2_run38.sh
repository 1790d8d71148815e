#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python -m pytest tests/test_gpu_queue.py tests/test_gpu_verify.py -m gpu -x -q -k "queue or merged_check or device_replay_sm" 2>&1 | tail -2
