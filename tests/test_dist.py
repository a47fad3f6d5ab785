"""world_size-2 gloo test (CPU) of the N>1 path's host logic: row sharding of witness + key, one all-gather of
partial commitments, combine.  The per-rank MSM is done by the CPU oracle here (test infrastructure standing in
for the CUDA kernel, which needs a GPU); the sharding / gather / combine plan is the product code under test."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import oracle
    from oracle import pyref as R
    from sirius_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    curve, ncols, n = R.CURVE_BN256, 3, 64
    W = oracle.random_field(R.FIELD_FR, 5, ncols * n)          # identical on every rank (same seed)
    T = oracle.random_field(R.FIELD_FR, 6, n)
    bases = oracle.running_bases(curve, ncols * n)
    W_loc = sharding.shard_column_major(W, ncols, n, rank, world)
    ck_loc = sharding.shard_column_major(bases, ncols, n, rank, world)
    T_loc = sharding.shard_column_major(T, 1, n, rank, world)
    out = {}
    for name, s_loc in (("W", W_loc), ("T", T_loc)):
        part = oracle.msm(curve, s_loc, ck_loc[: s_loc.shape[0]])  # prefix property of the local key
        t = torch.from_numpy(part.view(np.int64).copy())
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        acc = np.zeros(8, dtype=np.uint64)
        for g in gathered:
            acc = oracle.point_add(curve, acc, g.numpy().view(np.uint64))
        out[name] = acc
    if rank == 0:
        q.put((out["W"], oracle.msm(curve, W, bases), out["T"], oracle.msm(curve, T, bases)))
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_commit_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got_w, exp_w, got_t, exp_t = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got_w, exp_w)
    assert np.array_equal(got_t, exp_t)


def test_key_segments_cover_everything():
    from sirius_b200 import sharding

    ncols, n, world = 5, 32, 4
    seen = []
    for r in range(world):
        for first, count in sharding.key_segments(ncols, n, r, world):
            seen += list(range(first, first + count))
    assert sorted(seen) == list(range(ncols * n))
    with pytest.raises(ValueError):
        sharding.row_slice(0, 3, 32)
    with pytest.raises(ValueError):
        sharding.check_rotations_row_local([0, -1])
