"""Two ranks on two GPUs (NCCL): sharded likelihood == single-GPU likelihood bit for bit, and the
sharded PT driver all-gathers the cold-rung logL.  Skipped on boxes with fewer than 2 GPUs."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["HB_ROOT"])
import hb_mcmc_b200 as hb
from hb_mcmc_b200 import workload as wl
from hb_mcmc_b200.pt import ShardedPT, shard_ensembles
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
ctx = hb.Context(rank)
N, n = 3000, 512
t, flux, err = wl.make_dataset(N, wl.TRUTH_A, ctx.calc_light_curve)
ctx.set_data(t, flux, err)
P = wl.draw_chains(n, wl.TRUTH_A, ctx.roche_overflow, seed=1)
full = ctx.loglikelihood(P)                       # every rank: the whole batch
lo, hi = rank * n // world, (rank + 1) * n // world
mine = torch.from_numpy(ctx.loglikelihood(P[lo:hi])).cuda()   # its shard only
parts = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(parts, mine)
gathered = torch.cat(parts).cpu().numpy()
ok_shard = bool(np.array_equal(gathered, full, equal_nan=True))
sp = ShardedPT(ctx, 8, 6, float(wl.TRUTH_A[2]), seed=5, npast=20)
sp.sampler.init_random()
sp.step(25)
g = sp.gather_cold_logL_device().cpu().numpy()
_, local = sp.sampler.cold()
first, count = shard_ensembles(6, world, rank)
ok_pt = bool(np.array_equal(g[rank, :count], local)) and bool(np.isfinite(g[:, :count]).all())
host = sp.gather_cold_logL()
ok_host = bool(np.array_equal(host[first:first + count], local)) and host.shape == (6,)
if rank == 0:
    print(json.dumps({"ok_shard": ok_shard, "ok_pt": ok_pt, "ok_host": ok_host}))
dist.barrier(); dist.destroy_process_group()
'''


def test_two_gpu_sharding_and_allgather(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, HB_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    import json
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res == {"ok_shard": True, "ok_pt": True, "ok_host": True}, res
