"""N > 1 host logic on CPU: two gloo ranks shard the reference calls / MSM entries, gather, and agree on the merged result.
The device engine is replaced by a stand-in (no GPU here); what is under test is partitioning, ordering and the gather."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    import torch.distributed as dist

    import bpp

    par = __import__("importlib").import_module("bulletproofs-plus_b200.parallel")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        calls = [("call%d" % i, i) for i in range(7)]
        seen = []

        def fake_verify(params, local_calls, action):
            seen.extend(c[1] for c in local_calls)
            return [c[1] % 3 for c in local_calls], [["mask%d" % c[1]] for c in local_calls]

        status, masks = par.verify_chunks_distributed(None, calls, 2, verify_fn=fake_verify)

        class FakeEngine:   # "points" are integers mod 2^61-1 in 32 bytes; msm = sum s*P
            M = 2**61 - 1

            def msm(self, scalars, points):
                n = len(scalars) // 32
                acc = 0
                for i in range(n):
                    acc += int.from_bytes(scalars[32 * i:32 * i + 32], "little") * int.from_bytes(points[32 * i:32 * i + 32], "little")
                return (acc % self.M).to_bytes(32, "little")

        sc = b"".join((3 * i + 1).to_bytes(32, "little") for i in range(11))
        pt = b"".join((7 * i + 5).to_bytes(32, "little") for i in range(11))
        total = par.msm_distributed(FakeEngine(), sc, pt)
        q.put((rank, seen, status, masks, int.from_bytes(total, "little")))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything_once():
    import bpp

    par = __import__("importlib").import_module("bulletproofs-plus_b200.parallel")
    for n in (0, 1, 7, 16, 4096, 65538):
        for world in (1, 2, 3, 4, 8):
            spans = [par.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_two_gloo_ranks_merge_results():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, seen0, st0, mk0, tot0), (r1, seen1, st1, mk1, tot1) = out
    assert sorted(seen0 + seen1) == list(range(7)) and seen0 == [0, 1, 2, 3] and seen1 == [4, 5, 6]
    assert st0 == st1 == [i % 3 for i in range(7)]
    assert mk0 == mk1 == [["mask%d" % i] for i in range(7)]
    expect = sum((3 * i + 1) * (7 * i + 5) for i in range(11)) % (2**61 - 1)
    assert tot0 == tot1 == expect
