"""world_size-2 gloo tests (CPU) of the host-side multi-rank logic: chain sharding and the per-block all-reduce of
estimator accumulators used when one process per GPU runs the engine."""
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pimc_jl_b200.pimc import shard, allreduce_sum, chain_mean_over_ranks
    off, cnt = shard(4097, rank, world)
    # every rank owns a contiguous range of global chain ids; pretend block means come from those chains
    ids = np.arange(off, off + cnt, dtype=np.float64)
    block = np.stack([ids.mean() * np.ones(5), (ids ** 2).mean() * np.ones(5)])  # "E" and "Ev" block means of this rank
    mean = chain_mean_over_ranks(block, cnt)
    hist = np.zeros((4, 4))
    hist[rank, :] = cnt
    tot = allreduce_sum(hist)
    q.put((rank, off, cnt, mean.tolist(), tot.tolist()))
    dist.destroy_process_group()


def test_shard_and_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in ps]
    (r0, off0, c0, m0, t0), (r1, off1, c1, m1, t1) = res
    assert (off0, c0, off1, c1) == (0, 2049, 2049, 2048)
    allids = np.arange(4097, dtype=np.float64)
    for m in (m0, m1):  # the chain-weighted mean over ranks equals the mean over all global chains
        assert np.allclose(m[0], allids.mean()) and np.allclose(m[1], (allids ** 2).mean())
    assert t0 == t1 and t0[0][0] == 2049 and t0[1][0] == 2048


def test_shard_covers_all_chains():
    from pimc_jl_b200.pimc import shard
    for total in (1, 7, 4096, 4099):
        for world in (1, 2, 4, 8):
            spans = [shard(total, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == total
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
