"""Multi-GPU split of the accelerated path: one process per GPU (torch.distributed), no data-path collective.

Batch verification shards by reference call: every chunk (<= 256 proofs, one RangeProof::verify_batch call of the
reference, /root/reference/src/range_proof.rs:712-752) is independent -- its weights come from its own weight transcript
(:811-853) -- so rank r verifies chunks [lo_r, hi_r) on its own GPU and only the per-chunk statuses / masks (a few bytes)
are gathered.  Raw MSMs shard the (scalar, point) vectors; each GPU reduces its slice to ONE 32-byte point and the <= 8
partials are summed.  torch.distributed is plumbing only (rendezvous, gather of results, barriers).
"""
import torch.distributed as dist


def shard_range(n, world, rank):
    """contiguous balanced split of range(n): sizes differ by at most one, earlier ranks take the larger parts"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group):
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def verify_chunks_distributed(params, calls, action, group=None, verify_fn=None):
    """K independent verify_batch calls split over the ranks of `group`; every rank returns the full
    (status per call, masks per call) in call order.  `params` are this rank's device-resident RangeParameters."""
    if verify_fn is None:
        from .api import verify_chunks as verify_fn
    world, rank = _world(group)
    lo, hi = shard_range(len(calls), world, rank)
    status, masks = verify_fn(params, calls[lo:hi], action) if hi > lo else ([], [])
    if world == 1:
        return status, masks
    parts = [None] * world
    dist.all_gather_object(parts, (lo, status, masks), group=group)
    parts.sort(key=lambda t: t[0])
    all_status, all_masks = [], []
    for _, s, m in parts:
        all_status += s
        all_masks += m
    return all_status, all_masks


def msm_distributed(engine, scalars, points, group=None):
    """sum_i s_i * P_i with the (scalar, point) vectors sharded over the ranks: each GPU produces one partial point, the
    partials are gathered (32 bytes per rank) and added up with a <= world-point MSM with unit scalars on every rank."""
    world, rank = _world(group)
    n = len(scalars) // 32
    lo, hi = shard_range(n, world, rank)
    partial = engine.msm(scalars[32 * lo: 32 * hi], points[32 * lo: 32 * hi])
    if world == 1:
        return partial
    parts = [None] * world
    dist.all_gather_object(parts, partial, group=group)
    return sum_partials(engine, b"".join(parts))


def sum_partials(engine, parts):
    """<= 8 partial results (32-byte encodings) -> their sum: added up on the host (bpp_points_sum_host); engines without that
    entry point (the CPU stand-in of the gloo tests) fall back to a unit-scalar MSM"""
    from . import points_sum_host

    if hasattr(engine, "h"):
        return points_sum_host(parts)
    return engine.msm((1).to_bytes(32, "little") * (len(parts) // 32), parts)
