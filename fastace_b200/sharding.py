"""Economy sharding across the GPUs of one box.

Economies never interact (every bit of state hangs off one ``Economy`` object in the
reference, /root/reference/src/base/base.h:83-132), so the path shards by economy index
with NO collective in the step: rank r of W owns a contiguous block of economies.  The
only cross-rank traffic is the timing / metric reduction done by the caller.
"""


def shard_range(total_econ, rank, world):
    """Contiguous, balanced block [lo, hi) of economy indices owned by `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(int(total_econ), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owner_of(econ, total_econ, world):
    """Rank that owns global economy index `econ`."""
    base, extra = divmod(int(total_econ), int(world))
    cut = extra * (base + 1)
    if econ < cut:
        return econ // (base + 1)
    return extra + (econ - cut) // base


def reduce_max(value, dist=None, device=None):
    """max over ranks of a python float (device-timed milliseconds); identity when not distributed."""
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
