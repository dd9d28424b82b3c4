"""Channel sharding across the GPUs of one box.

The reference has no multi-device code: one object = one channel = one device (reference cl_fft.cpp:49,
cl_conv.cpp:154, cl_dconv.cpp:53). Channels and transform batches never exchange data, so the B200
deployment is one process per GPU, each owning a contiguous range of channels and all of their
device state; there is NO data-path collective. torch.distributed is used only
to agree on the timing of a step (barrier + max over ranks).
"""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) of `total` units for `rank` of `world` (sizes differ by at most 1)."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value: float) -> float:
    """Slowest rank's value (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
