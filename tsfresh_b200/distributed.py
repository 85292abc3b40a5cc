"""Multi-GPU layout of the hot path: one process per GPU (torch.distributed), series sharded by id.

The reference's own parallelism is data-parallel over lists of series (utilities/distribution.py:118-148,
471-486); the series never interact, so the only exchange step is assembling the final
[n_ids x n_features] matrix.  Rank r owns the contiguous row range shard_bounds(n, world, r) and extracts its rows
with its own tsfx context.

Placement of the result (`GatheredMatrix`): the full matrix lives in symmetric memory (every rank maps every other
rank's copy, torch.distributed._symmetric_memory).  The library's assemble pass writes each finished row block into
the local copy AND into the peers' copies (tsfx_set_peer_outputs: one multicast store through the NVSwitch, or copy
engines / P2P stores over NVLink), block by block while the kernels of the next block run -- there is no separate
collective kernel and no second pass over the matrix.  Without symmetric memory (CPU tests with gloo, or a box
without P2P) the same row blocks are exchanged with all_gather_into_tensor on a side stream.
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
    dev = torch.device("cuda", ctx.device)
    per = shard_rows(n, world)
    if dist.get_backend(group) == "nccl" and hi > lo:
        # device path: this rank's rows are placed on every rank while the later row blocks are still being computed
        gm = GatheredMatrix(per, plan.n_cols, dev, group)
        local = torch.from_numpy(np.ascontiguousarray(values[lo:hi], dtype=np.float32)).to(dev)
        if hi - lo < per:
            gm.local[hi - lo:].fill_(float("nan"))
        torch.cuda.synchronize(dev)                 # the library launches on its own stream
        gm.attach(ctx)
        try:
            extract_dense_sharded_device(dp, local, gm)
            gm.finish(ctx)
        finally:
            gm.detach(ctx)
        torch.cuda.synchronize(dev)
        dist.barrier(group)
        return plan.suffixes, gm.full[:n]
    local = dp.extract_dense(np.ascontiguousarray(values[lo:hi], dtype=np.float32)) if hi > lo else \
        np.empty((0, plan.n_cols))
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


class GatheredMatrix:
    """The [world * rows_per_rank, n_cols] float64 feature matrix, replicated on every rank's GPU.

        gm = GatheredMatrix(rows_per_rank, n_cols, device)        # collective: every rank constructs it
        gm.attach(ctx)                                            # rows written by ctx now also land on the peers
        ... dp.extract_*_device(..., out_ptr=gm.local_ptr(row)) ...
        gm.finish(ctx)                                            # all ranks' rows are in gm.full after this
    """

    def __init__(self, rows_per_rank, n_cols, device, group=None, mode="auto", n_blocks=7):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rows, self.n_cols, self.device = int(rows_per_rank), int(n_cols), device
        self.kind, self.hdl, self.peer_ptrs, self.mc_ptr = "local", None, None, 0
        self.mode = mode
        self.n_blocks = n_blocks
        self._pending = []                 # (lo, hi) row blocks not yet exchanged (nccl fallback)
        self._comm = None
        shape = (self.world * self.rows, self.n_cols)
        if self.world > 1 and device.type == "cuda" and mode != "nccl":
            try:
                import torch.distributed._symmetric_memory as symm
                self.full = symm.empty(shape, dtype=torch.float64, device=device)
                self.hdl = symm.rendezvous(self.full, group if group is not None else dist.group.WORLD)
                self.peer_ptrs = [int(p) for p in self.hdl.buffer_ptrs]
                try:
                    self.mc_ptr = int(self.hdl.multicast_ptr) if self.hdl.has_multicast_support(device.type, device.index) else 0
                except Exception:
                    self.mc_ptr = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
                self.kind = "peer"
            except Exception as e:            # no symmetric memory on this box: NCCL exchange of the row blocks
                self.why_not_peer = repr(e)
                self.full = torch.empty(shape, dtype=torch.float64, device=device)
                self.kind = "nccl"
        else:
            self.full = torch.empty(shape, dtype=torch.float64, device=device)
            self.kind = "nccl" if self.world > 1 else "local"
        if self.kind == "nccl" and device.type == "cuda":
            self._comm = torch.cuda.Stream(device=device)

    def block_bounds(self, span):
        """Row blocks of one pass over `span` rows.  The placement of block b overlaps the kernels of block b + 1, so only
        the LAST block's transfer is exposed: the blocks shrink geometrically towards the end (1/4, 1/4, 1/4, 1/8, 1/16,
        1/32, 1/32 of the rows for the default of 7) -- big blocks keep the kernels' grids full, the exposed tail is 1/32
        of the matrix instead of 1/8."""
        if self.world <= 1 or self.n_blocks <= 1 or span < 64 * self.n_blocks:
            return [0, span]
        nb = self.n_blocks
        fr = [0.25, 0.5, 0.75]
        rest, x = nb - 3, 0.75
        for k in range(rest - 1):
            x += 0.25 / (2 ** (k + 1))
            fr.append(x)
        cuts = sorted(set([0] + [int(span * f) for f in fr] + [span]))
        return cuts

    @property
    def local(self):
        return self.full[self.rank * self.rows:(self.rank + 1) * self.rows]

    def local_ptr(self, row=0):
        return self.local.data_ptr() + int(row) * self.n_cols * 8

    def placement(self):
        from . import _lib
        if self.kind != "peer":
            return self.kind
        want = {"auto": _lib.PEER_AUTO, "copy": _lib.PEER_COPY, "store": _lib.PEER_STORE, "multicast": _lib.PEER_MULTICAST}[self.mode]
        if want == _lib.PEER_AUTO:
            want = _lib.PEER_COPY
        return {_lib.PEER_COPY: "copy engines over NVLink", _lib.PEER_STORE: "P2P stores from the assemble kernel",
                _lib.PEER_MULTICAST: "multicast stores from the assemble kernel (NVSwitch)"}[want]

    def attach(self, ctx):
        from . import _lib
        if self.kind == "peer":
            mode = {"auto": _lib.PEER_AUTO, "copy": _lib.PEER_COPY, "store": _lib.PEER_STORE,
                    "multicast": _lib.PEER_MULTICAST}[self.mode]
            ctx.set_peer_outputs(self.peer_ptrs, self.rank, self.mc_ptr, mode)

    def detach(self, ctx):
        if self.kind == "peer":
            ctx.set_peer_outputs([], 0)

    def block_done(self, lo, hi, stream):
        """rows [lo, hi) of the local shard are final on `stream` (nccl fallback: exchange them now, overlapped)"""
        if self.kind != "nccl":
            return
        if self.device.type != "cuda":                 # gloo on the host (CPU tests): a plain blocking exchange
            stage = torch.empty((self.world, hi - lo, self.n_cols), dtype=torch.float64)
            dist.all_gather_into_tensor(stage.view(-1), self.local[lo:hi].reshape(-1).contiguous(), group=self.group)
            self.full.view(self.world, self.rows, self.n_cols)[:, lo:hi].copy_(stage)
            return
        ev = torch.cuda.Event()
        ev.record(stream)
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ev)
            stage = torch.empty((self.world, hi - lo, self.n_cols), dtype=torch.float64, device=self.device)
            dist.all_gather_into_tensor(stage.view(-1), self.local[lo:hi].reshape(-1), group=self.group)
            self.full.view(self.world, self.rows, self.n_cols)[:, lo:hi].copy_(stage)

    def finish(self, ctx, stream=None, ctx_on_current_stream=False):
        """Every rank's rows are visible in self.full on every rank once this returns (and the stream has run).
        ctx_on_current_stream: the context was created on torch's current stream, so the barrier below is ordered
        after the placement by stream order; otherwise the context's stream is synchronised first."""
        if self.kind == "peer":
            ctx.peer_flush()
            if not ctx_on_current_stream:
                ctx.sync()
            self.hdl.barrier(channel=0)
        elif self.kind == "nccl" and self._comm is not None:
            (stream or torch.cuda.current_stream(self.device)).wait_stream(self._comm)


def extract_dense_sharded_device(dp, values, gm, stream=None, ctx_on_current_stream=False):
    """values: this rank's [rows, L] float32 CUDA tensor; dp: DevicePlan on a context created on `stream`.
    Runs the pass in gm.n_blocks row blocks (so that placement / exchange of block b overlaps the kernels of block
    b + 1) and leaves this rank's rows everywhere (gm)."""
    S, L = values.shape
    stream = stream or torch.cuda.current_stream(values.device)
    span = gm.rows if gm.world > 1 else S      # every rank cuts the SAME row blocks (a short last shard is padded by the caller)
    cuts = gm.block_bounds(span)
    nb = len(cuts) - 1
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        if lo < S:
            dp.extract_dense_device(values[lo:min(hi, S)].data_ptr(), min(hi, S) - lo, L, gm.local_ptr(lo))
        if gm.kind == "nccl" and not ctx_on_current_stream:
            dp.ctx.sync()                  # the library ran on its own stream: order the exchange after it
        gm.block_done(lo, hi, stream)
    return nb


def extract_csr_sharded_device(dp, values, begin, length, gm, stream=None, ctx_on_current_stream=False, max_len=0):
    """Same for a CSR shard (e.g. rolled windows sharded by parent, shard_windows): begin / length are this rank's
    int64 / int32 CUDA tensors over the shared `values` buffer."""
    S = int(begin.shape[0])
    stream = stream or torch.cuda.current_stream(values.device)
    span = gm.rows if gm.world > 1 else S
    cuts = gm.block_bounds(span)
    nb = len(cuts) - 1
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        if lo < S:
            h = min(hi, S)
            dp.extract_csr_device(values.data_ptr(), values.numel(), begin[lo:h].data_ptr(), length[lo:h].data_ptr(), h - lo,
                                  gm.local_ptr(lo), max_len=max_len)
        if gm.kind == "nccl" and not ctx_on_current_stream:
            dp.ctx.sync()                  # the library ran on its own stream: order the exchange after it
        gm.block_done(lo, hi, stream)
    return nb
