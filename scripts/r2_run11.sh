#!/bin/bash
# N=2 torchrun bench + world-size-2 GPU tests if any
cd ${GRAFT_REPO_ROOT:-.}
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_b11_n2.json 2> gpurun_out/r2_b11_n2.err; echo "n2 rc=$?"; tail -3 gpurun_out/r2_b11_n2.err
python scripts/r2_summary.py gpurun_out/r2_b11_n2.json 2>&1 | head -40
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_b11_n1.json 2> gpurun_out/r2_b11_n1.err; echo "n1 rc=$?"
python scripts/r2_summary.py gpurun_out/r2_b11_n1.json 2>&1 | grep "^value\|^one_shot"
