"""Stream sharding across GPUs: contiguous stream ranges per rank, templates replicated,
no collective on the data path. torch.distributed is used only to agree on timing and to gather the
per-rank results (rank 0 concatenates detection lists in stream order)."""
from __future__ import annotations


def shard_range(n_streams: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the streams rank `rank` owns: contiguous, sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_streams, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_step_stats(dist, device, elapsed_ms: float, units: float):
    """max over ranks of the elapsed time, sum over ranks of the processed units (windows)."""
    import torch
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    u = torch.tensor([units], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])


def gather_detections(dist, local: list, lo: int):
    """local: [(stream_in_shard, chunk, detection dict)] -> on rank 0 the global list with absolute
    stream indices, ordered by (stream, chunk); other ranks get None."""
    item = [(lo + s, c, d) for s, c, d in local]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return sorted(item, key=lambda x: (x[0], x[1]))
    out = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(item, out, dst=0)
    if dist.get_rank() != 0:
        return None
    merged = [x for part in out for x in part]
    return sorted(merged, key=lambda x: (x[0], x[1]))
