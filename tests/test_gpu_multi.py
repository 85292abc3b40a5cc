"""Multi-GPU path of the PRODUCT (tsfresh_b200.distributed) over NCCL / symmetric memory, world size 2: every rank
extracts its shard and finds every rank's rows in its own copy of the matrix.  Needs two GPUs (skipped otherwise;
run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    try:
        from tsfresh_b200 import EfficientFCParameters, _lib
        from tsfresh_b200.distributed import GatheredMatrix, extract_dense_sharded, extract_dense_sharded_device
        from tsfresh_b200.plan import Plan
        rng = np.random.default_rng(11)
        values = rng.standard_normal((3001, 64)).astype(np.float32)          # odd count: the last shard is padded
        plan = Plan(EfficientFCParameters())
        # (1) device path with an explicit placement mode, context on torch's stream
        stream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(stream)
        ctx = _lib.Context(rank, stream=stream.cuda_stream)
        dp = _lib.DevicePlan(ctx, plan)
        per = (len(values) + world - 1) // world
        lo, hi = rank * per, min(len(values), (rank + 1) * per)
        gm = GatheredMatrix(per, plan.n_cols, dev, mode=mode, n_blocks=4)
        local = torch.from_numpy(values[lo:hi]).to(dev)
        if hi - lo < per:
            gm.local[hi - lo:].fill_(float("nan"))
        torch.cuda.synchronize()
        dist.barrier()
        gm.attach(ctx)
        extract_dense_sharded_device(dp, local, gm, stream=stream, ctx_on_current_stream=True)
        gm.finish(ctx, stream=stream, ctx_on_current_stream=True)
        gm.detach(ctx)
        torch.cuda.synchronize()
        dist.barrier()
        full = gm.full[:len(values)].cpu().numpy()
        single = dp.extract_dense(values)
        ok1 = bool(np.array_equal(full, single, equal_nan=True))
        placement = gm.placement()
        # (2) the product entry point on host arrays (own context / stream)
        cols, full2 = extract_dense_sharded(values, EfficientFCParameters(), device=rank)
        ok2 = bool(np.array_equal(full2.cpu().numpy(), single, equal_nan=True)) and list(cols) == list(plan.suffixes)
        ret[rank] = (ok1, ok2, gm.kind, placement)
        dp.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["auto", "copy", "store", "nccl"])
def test_world_size_2_every_rank_holds_the_full_matrix(mode):
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), mode, ret), nprocs=2, join=True)
    assert len(ret) == 2
    for r in (0, 1):
        ok1, ok2, kind, placement = ret[r]
        assert ok1, (r, kind, placement)
        assert ok2, (r, kind, placement)


def test_n_jobs_drives_several_gpus_from_one_process():
    """extract_features(n_jobs=2): the frame is cut at id boundaries, each GPU fills its rows of one pinned matrix"""
    torch = pytest.importorskip("torch")
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import pandas as pd
    from tsfresh_b200 import MinimalFCParameters, extract_features
    rng = np.random.default_rng(3)
    S, L = 24_000, 64                                   # 1.5 M rows: above the multi-GPU threshold
    df = pd.DataFrame({"id": np.repeat(np.arange(S, dtype=np.int64) * 3, L), "time": np.tile(np.arange(L, dtype=np.int64), S),
                       "value": rng.standard_normal(S * L).astype(np.float32)})
    one = extract_features(df, column_id="id", column_sort="time", default_fc_parameters=MinimalFCParameters(), n_jobs=1)
    two = extract_features(df, column_id="id", column_sort="time", default_fc_parameters=MinimalFCParameters(), n_jobs=2)
    assert list(one.index) == list(two.index) and list(one.columns) == list(two.columns)
    assert np.array_equal(one.to_numpy(), two.to_numpy(), equal_nan=True)
    # rows out of order: falls back to one GPU (device sort), same frame
    shuffled = df.sample(frac=1.0, random_state=0)
    three = extract_features(shuffled, column_id="id", column_sort="time", default_fc_parameters=MinimalFCParameters(), n_jobs=2)
    assert np.array_equal(one.to_numpy(), three.to_numpy(), equal_nan=True) and list(one.index) == list(three.index)
