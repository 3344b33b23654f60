"""Sharding of the independent curve fits over the GPUs of one box: contiguous ranges of the
fit index, no data-path collective, one gather of the fitted parameters at the end (NCCL over NVLink on the
GPUs; the same code runs over gloo on CPUs for the host-logic tests)."""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the fits rank `rank` owns: contiguous, sizes differ by at most one, ranks in index order."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_rows(local, total: int, rank: int, world: int):
    """All-gathers per-rank blocks [b_r, ...] (the rows a rank owns by shard_range) into the global [total, ...] tensor;
    every rank gets the full result, like the reference's single address space."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local
    sizes = [shard_range(total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def gather_rows_to_root(local, total: int, rank: int, world: int, dst: int = 0):
    """The same exchange with one consumer: rank `dst` receives the global [total, ...] tensor, the others None.  This is
    what a caller that continues on one host (the reference's single process) needs, and moves world-times less data
    than the all-gather."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local
    sizes = [shard_range(total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, out, dst=dst)
    if rank != dst:
        return None
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def gather_fits(local_p, total: int, rank: int, world: int):
    """The fits' only exchange: the fitted parameter blocks [b_r, m] of all ranks -> [total, m]."""
    return gather_rows(local_p, total, rank, world)


def gather_gate(local_idx, total: int, rank: int, world: int):
    """Mahalanobis gate sharded by measurement: every rank holds the landmark state (x, P: the packed
    landmarks are 1.9 MB at 20k landmarks) and gates its contiguous range of the measurements; the association indices
    [b_r] int32 are all-gathered (40 KB for 10k measurements)."""
    return gather_rows(local_idx, total, rank, world)
