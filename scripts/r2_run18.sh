#!/bin/bash
# N=8 torchrun bench (the driver's command shape)
cd ${GRAFT_REPO_ROOT:-.}
nvidia-smi -L | wc -l; nproc
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_b18_n8.json 2> gpurun_out/r2_b18_n8.err; echo "n8 rc=$?"; tail -3 gpurun_out/r2_b18_n8.err
python scripts/r2_summary.py gpurun_out/r2_b18_n8.json 2>&1 | cut -c1-700 | head -30
