"""world_size-2 gloo test (CPU) of the N>1 host logic: shard partition, per-rank sub-batches, gathers.
The step backend here is the oracle library (test infrastructure) because there is no GPU in this container;
the sharding code is backend-agnostic and takes the library as a parameter."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from parity_util import B, orc_lib
from ode_b200 import scenes
from ode_b200.shard import shard_range, ShardedBatch


def test_shard_range_partitions():
    for W in (1, 2, 7, 4096, 65536, 16385):
        for G in (1, 2, 3, 4, 8):
            parts = [shard_range(W, r, G) for r in range(G)]
            assert parts[0][0] == 0 and parts[-1][1] == W
            for (b0, e0), (b1, e1) in zip(parts, parts[1:]):
                assert e0 == b1
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world_size, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        sc = scenes.box_stack(nworlds=5, nboxes=4)
        sb = ShardedBatch(orc_lib("single"), sc, rank, world_size)
        sb.step(0.02, 30)
        stats = sb.gather_stats(dist)
        obs = sb.gather_observations(dist)
        if rank == 0:
            q.put((stats, obs))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_run_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    stats, obs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sc = scenes.box_stack(nworlds=5, nboxes=4)
    whole = B.Batch(orc_lib("single"), sc)
    whole.step(0.02, 30)
    st = whole.get_state()
    ref_obs = np.concatenate([st["pos"], st["quat"], st["lvel"], st["avel"]], axis=-1)
    assert obs.shape == ref_obs.shape and np.array_equal(obs, ref_obs)
    ref_stats = np.stack([whole.get_stats(w) for w in range(5)])
    assert np.array_equal(stats, ref_stats)
