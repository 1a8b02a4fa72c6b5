"""The sampler split over several GPUs walks the chains of the one-GPU run bit for bit -- checked here on ONE GPU with
several samplers standing in for the ranks (the host-driven exchange hb_pt_exchange_local takes the place of the NCCL
all-gather; tests/test_gpu_multi.py runs the real thing on 2 GPUs).

  rung split      every rank holds the whole ladder, evaluates the likelihood of its shard of the walkers only; the
                  logL vector is exchanged and the swaps (mcmc_wrapper2.c:554-563, ptmcmc :768-817) are decided
                  identically everywhere
  ensemble split  whole ladders per rank, Philox streams keyed on the GLOBAL ensemble id"""
import numpy as np
import pytest

from hb_mcmc_b200 import workload as wl
from hb_mcmc_b200.pt import PTSampler, exchange_local, shard_walkers

pytestmark = pytest.mark.gpu


@pytest.fixture()
def data(ctx):
    N = 3000
    t, flux, err = wl.make_dataset(N, wl.TRUTH_A, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    ctx.set_mags([1000, 1, 1, 1, 1], [1e15] * 4, 1, 0)
    return float(wl.TRUTH_A[2])


def _same(a, b):
    for u, v in zip(a.state(), b.state()):
        assert np.array_equal(u, v, equal_nan=True)
    assert np.array_equal(a.map()[1], b.map()[1])
    ca, cb = a.counters(), b.counters()
    assert all(np.array_equal(ca[k], cb[k]) for k in ca)


@pytest.mark.parametrize("world,n_temps,n_ens", [(2, 16, 1), (3, 50, 1), (8, 64, 1), (4, 5, 3), (8, 6, 1)])
def test_rung_split_equals_one_gpu(ctx, data, world, n_temps, n_ens):
    ref = PTSampler(ctx, n_temps, n_ens, data, seed=7, npast=12)
    ref.init_random()
    ranks = [PTSampler(ctx, n_temps, n_ens, data, seed=7, npast=12) for _ in range(world)]
    W = n_temps * n_ens
    for r, s in enumerate(ranks):
        s.init_random()
        s.set_eval_shard(r, world)
        assert s.eval_shard() == shard_walkers(W, world, r)
        with pytest.raises(Exception):
            s.step(1)  # sharded, but neither a communicator nor the host-driven phases
    steps = 30  # past npast: differential-evolution proposals from the history rings are in
    ref.step(steps)
    for _ in range(steps):
        for s in ranks:
            s.step_begin()
        exchange_local(ranks)
        for s in ranks:
            s.step_end()
    for s in ranks:
        _same(s, ref)
        assert s.iteration == steps
    for s in ranks + [ref]:
        s.close()


def test_ensemble_split_equals_one_gpu(ctx, data):
    n_temps, n_ens = 8, 6
    ref = PTSampler(ctx, n_temps, n_ens, data, seed=3, npast=10)
    ref.init_random()
    ref.step(25)
    x, ll, idx = ref.state()
    x = x.reshape(n_ens, n_temps, 21)
    ll = ll.reshape(n_ens, n_temps)
    for first, count in ((0, 2), (2, 3), (5, 1)):
        part = PTSampler(ctx, n_temps, count, data, seed=3, npast=10, ens_offset=first)
        part.init_random()
        part.step(25)
        px, pl, pidx = part.state()
        assert np.array_equal(px.reshape(count, n_temps, 21), x[first:first + count])
        assert np.array_equal(pl.reshape(count, n_temps), ll[first:first + count], equal_nan=True)
        assert np.array_equal(pidx, idx[first:first + count])
        part.close()
    ref.close()
