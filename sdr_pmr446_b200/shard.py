"""Stream sharding across the GPUs of one box.

Streams are independent, so the data path needs no collective: rank r owns a contiguous block of streams and runs
the identical pipeline on it.  torch.distributed (NCCL on the GPU box, gloo in CPU tests) is used only to gather
per-rank statistics -- samples processed, elapsed time, checksums -- and to take the max-over-ranks time.
"""
import numpy as np


def shard_range(total_streams: int, world: int, rank: int):
    """Contiguous block [start, start + count) of rank `rank`; blocks differ by at most one stream."""
    base, extra = divmod(total_streams, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def stream_checksum(pcm: np.ndarray) -> int:
    """Order-sensitive 64-bit checksum of one stream's s16 audio [channels][samples]."""
    v = np.ascontiguousarray(pcm).astype(np.int64).ravel()
    w = (np.arange(v.size, dtype=np.int64) % 8191) + 1
    return int((v * w).sum() & 0x7FFFFFFFFFFFFFFF)


def gather_stats(local: dict, device=None):
    """All-gathers a small dict of numbers; returns the list of per-rank dicts (on every rank)."""
    import torch
    import torch.distributed as dist
    keys = sorted(local)
    t = torch.tensor([float(local[k]) for k in keys], dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [dict(zip(keys, t.tolist()))]
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [dict(zip(keys, o.tolist())) for o in out]


def max_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
