#!/usr/bin/env python
"""Raw Ristretto255 MSM sweep sharded over the GPUs of one box (BASELINE.json configs[4]): rank r owns points/scalars
[r*N/G, (r+1)*N/G) device-resident (bpp_msm_plan), reduces them to ONE 32-byte partial point, the partials are gathered (NCCL
all_gather of 32 bytes per rank) and summed by rank 0 with a G-term MSM.  Time = max over ranks, CUDA events + the gather.
  python scripts/msm_multi.py [log2 sizes ...]                     (1 GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29533 scripts/msm_multi.py 20 22 24
Prints one JSON line per size on rank 0; every size is checked against the 1-GPU result of the same inputs (rank 0 recomputes
sizes <= 2^20 alone)."""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bpp  # noqa: E402

par = __import__("importlib").import_module("bulletproofs-plus_b200.parallel")
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dist = None
if world > 1:
    os.environ.pop("NCCL_DEBUG", None)
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = bpp.pkg.Engine(local)
sizes = [int(x) for x in sys.argv[1:]] or [16, 20, 22]
base = eng.from_uniform(hashlib.shake_256(b"sweep").digest(64 * (1 << 14)))      # 16 k distinct points, repeated (duplicates are legal input)


def inputs(lg, lo, hi):
    n = hi - lo
    # point i = base[i mod 16384]; scalar i = SHAKE256("sc" || lg) block i, top nibble cleared (< 2^252 < l)
    reps = (hi + (1 << 14) - 1) // (1 << 14) + 1
    pts = (base * reps)[32 * lo: 32 * hi] if hi <= (1 << 14) * reps else None
    allsc = hashlib.shake_256(b"sc%d" % lg).digest(32 * (1 << lg))
    sc = bytearray(allsc[32 * lo: 32 * hi])
    for i in range(31, len(sc), 32):
        sc[i] &= 0x0F
    return bytes(sc), pts, n


for lg in sizes:
    N = 1 << lg
    lo, hi = par.shard_range(N, world, rank)
    sc, pts, n = inputs(lg, lo, hi)
    plan = bpp.pkg.MsmPlan(eng, pts)
    plan.set_scalars(sc)
    buf = torch.zeros(32, dtype=torch.uint8, device="cuda")
    gathered = torch.zeros(32 * world, dtype=torch.uint8, device="cuda")

    def once():
        partial = plan.run(True)
        if world == 1:
            return partial
        buf.copy_(torch.frombuffer(bytearray(partial), dtype=torch.uint8))
        dist.all_gather_into_tensor(gathered, buf)
        parts = bytes(gathered.cpu().numpy())
        return par.sum_partials(eng, parts)

    res = once()
    reps = 5 if lg <= 20 else 3
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        r2 = once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    assert r2 == res
    # kernel-only time of this rank's shard (CUDA events on the engine stream)
    eng.timer_start()
    for _ in range(reps):
        plan.run(False)
    dev_ms = eng.timer_stop() / reps
    t = torch.tensor([dt, dev_ms * 1e-3], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = None
    if rank == 0 and world > 1 and lg <= 20:       # the sharded result equals the single-GPU result
        sc_all, pts_all, _ = inputs(lg, 0, N)
        p1 = bpp.pkg.MsmPlan(eng, pts_all)
        p1.set_scalars(sc_all)
        ok = p1.run(True) == res
        p1.close()
        assert ok
    if rank == 0:
        print(json.dumps({"metric": "raw Ristretto255 MSM", "log2_points": lg, "n_gpus": world, "mpoints_per_s": N / float(t[0]) / 1e6,
                          "ms": float(t[0]) * 1e3, "kernel_ms_max_over_ranks": float(t[1]) * 1e3,
                          "mpoints_per_s_kernels_only": N / float(t[1]) / 1e6, "window_bits": plan.window_bits,
                          "equals_single_gpu_result": ok}), flush=True)
    plan.close()
if dist:
    dist.barrier()
    dist.destroy_process_group()
