"""Multi-rank host logic on CPU: world_size-2 gloo process group (no GPU).  The device sampler is
replaced by a stub so that only the sharding / all-gather / global-MAP plumbing is exercised."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hb_mcmc_b200 import pt as ptmod


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_shard_ensembles_partition():
    for n_ens in (2, 7, 8, 256, 257):
        for world in (1, 2, 3, 8):
            if n_ens < world:
                with pytest.raises(ValueError):
                    ptmod.shard_ensembles(n_ens, world, 0)
                continue
            blocks = [ptmod.shard_ensembles(n_ens, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n_ens
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


class StubSampler:
    """cold logL = -(global ensemble id), cold x = id in every slot"""

    def __init__(self, first, count):
        self.first, self.count, self.steps = first, count, 0

    def step(self, n):
        self.steps += n

    def cold(self):
        ids = np.arange(self.first, self.first + self.count, dtype=np.float64)
        return np.repeat(ids[:, None], 21, axis=1), -np.abs(ids - 4.0)  # best (0.0) at ensemble 4


def _worker(rank, world, port, n_ens, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sp = ptmod.ShardedPT.__new__(ptmod.ShardedPT)
        sp.dist, sp.rank, sp.world, sp.n_ens_total = dist, rank, world, n_ens
        sp.mode, sp.comm = "ensembles", None
        sp.first, sp.count = ptmod.shard_ensembles(n_ens, world, rank)
        sp.sampler = StubSampler(sp.first, sp.count)
        sp.step(3)
        allL = sp.gather_cold_logL(device="cpu")
        xc, _ = sp.sampler.cold()
        best, val, x_best = ptmod.global_map(allL, xc, sp.first, sp.count)
        out.put((rank, allL.tolist(), best, val, None if x_best is None else float(x_best[0]), sp.sampler.steps))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_ens", [7, 8])
def test_all_gather_cold_logL_world2(n_ens):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_ens, out)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    want = (-np.abs(np.arange(n_ens) - 4.0)).tolist()
    for rank, allL, best, val, xb, steps in res:
        assert allL == want and best == 4 and val == 0.0 and steps == 3
    owners = [xb for _, _, _, _, xb, _ in res]
    assert owners.count(None) == 1 and 4.0 in owners  # exactly one rank owns the MAP ensemble


@pytest.mark.parametrize("W,world", [(64, 8), (50, 8), (7, 8), (2048, 3), (1, 2), (64, 64)])
def test_walker_shards_tile_the_sampler(W, world):
    """hb_pt_set_eval_shard's partition (mirrored by shard_walkers): contiguous chunks of ceil(W / world) that tile
    [0, W) in rank order, so that the in-place all-gather of `chunk` doubles per rank lands every walker at its index."""
    covered = []
    for r in range(world):
        first, count, chunk = ptmod.shard_walkers(W, world, r)
        assert chunk == -(-W // world) and 0 <= count <= chunk
        assert first == min(r * chunk, W)
        covered += list(range(first, first + count))
    assert covered == list(range(W))
    assert world * chunk >= W and world * chunk < W + world
