"""Multi-GPU layout of the hot path: one process per GPU (torch.distributed), series sharded by id.

The reference's own parallelism is data-parallel over lists of series (utilities/distribution.py:118-148,
471-486); the series never interact, so the only exchange step is assembling the final
[n_ids x n_features] matrix.  Rank r owns the contiguous row range shard_bounds(n, world, r); every rank
extracts its rows with its own tsfx context and ONE all-gather of the (padded) row blocks rebuilds the full
matrix on every rank (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_rows(n_rows, world):
    """rows per rank (every rank gets the same padded count so one all_gather_into_tensor suffices)"""
    return (n_rows + world - 1) // world


def shard_bounds(n_rows, world, rank):
    per = shard_rows(n_rows, world)
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def gather_rows(local, n_rows, group=None):
    """local: [hi - lo, F] tensor of this rank's rows (device or CPU).  Returns the full [n_rows, F] tensor
    on every rank.  Ranks whose shard is short (the tail) are padded with NaN rows that are dropped again."""
    world = dist.get_world_size(group)
    per = shard_rows(n_rows, world)
    F = local.shape[1]
    if local.shape[0] != per:
        pad = torch.full((per - local.shape[0], F), float("nan"), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    full = torch.empty((world * per, F), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full, local.contiguous(), group=group)
    return full[:n_rows]


def extract_dense_sharded(values, fc_parameters, device=None, group=None):
    """values: [n_series, L] float32 host array present on every rank (or only the local shard is read).
    Every rank extracts its contiguous shard on its own GPU; returns (columns, full matrix as a torch tensor
    on the rank's device)."""
    import numpy as np

    from .extraction import _device_plan, get_context
    from .plan import Plan

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = values.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    plan = Plan(fc_parameters)
    ctx = get_context(device)
    dp = _device_plan(ctx, plan)
    local = dp.extract_dense(np.ascontiguousarray(values[lo:hi], dtype=np.float32)) if hi > lo else \
        np.empty((0, plan.n_cols))
    dev = torch.device("cuda", ctx.device)
    full = gather_rows(torch.from_numpy(local).to(dev), n, group)
    return plan.suffixes, full


def shard_windows(parent, n_parents, world, rank):
    """Rolled windows (tsfresh_b200.rolling / tsfx_roll_windows) shard by PARENT series so that a window never
    straddles ranks: contiguous parent ranges with as equal window counts as whole parents allow.
    `parent[w]` is the (ascending) parent index of window w.  Returns (lo, hi): this rank owns windows [lo, hi)."""
    import numpy as np
    parent = np.asarray(parent)
    n = len(parent)
    if n == 0:
        return 0, 0
    # first window of every parent that has windows, and the ideal split points in window units
    starts = np.flatnonzero(np.concatenate([[True], parent[1:] != parent[:-1]]))
    cuts = [0]
    for r in range(1, world):
        target = (n * r) // world
        k = int(np.searchsorted(starts, target, side="left"))     # next parent boundary at or after the target
        cuts.append(int(starts[k]) if k < len(starts) else n)
    cuts.append(n)
    for r in range(1, len(cuts)):                                 # monotone (small inputs: several ranks may be empty)
        cuts[r] = max(cuts[r], cuts[r - 1])
    return cuts[rank], cuts[rank + 1]
