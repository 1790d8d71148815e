#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests/test_gpu_primitives.py tests/test_gpu_lanes.py tests/test_gpu_queue.py tests/test_gpu_verify.py tests/test_gpu_prove.py -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 --lanes 4 > gpurun_out/r2_b7.json 2> gpurun_out/r2_b7.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_b7.err
python scripts/r2_summary.py gpurun_out/r2_b7.json 2>&1 | grep -v "^clocks\|^cpu"
python - <<'P'
import torch, time
# host<->device link: pinned 64 MiB copies
a = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); b = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
for name, src, dst in (("h2d", a, b), ("d2h", b, a)):
    dst.copy_(src); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    print(name, "GB/s", 10 * (64 << 20) / (time.perf_counter() - t0) / 1e9)
P
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" 
