"""Two ranks on two GPUs (NCCL): sharded likelihood == single-GPU likelihood bit for bit; the sampler split by
ensembles and split by rungs (logL all-gathered by NCCL from inside libhb_b200's captured step) walks the one-GPU
chains bit for bit.  Skipped on boxes with fewer than 2 GPUs."""
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
from hb_mcmc_b200.pt import PTSampler
logp = float(wl.TRUTH_A[2])
# (1) whole ensembles per rank: this rank's ladders are the one-GPU run's ladders, bit for bit; host gather of the cold rung
sp = ShardedPT(ctx, 8, 6, logp, seed=5, npast=10)
sp.sampler.init_random()
sp.step(25)
one = PTSampler(ctx, 8, 6, logp, seed=5, npast=10)
one.init_random()
one.step(25)
first, count = shard_ensembles(6, world, rank)
x1, l1, i1 = one.state()
xs, ls, is_ = sp.sampler.state()
ok_pt = sp.mode == "ensembles" and bool(np.array_equal(xs, x1.reshape(6, 8, 21)[first:first + count].reshape(-1, 21))) \
    and bool(np.array_equal(ls, l1.reshape(6, 8)[first:first + count].ravel(), equal_nan=True)) and bool(np.array_equal(is_, i1[first:first + count]))
host = sp.gather_cold_logL()
ok_host = bool(np.array_equal(host, one.cold()[1])) and host.shape == (6,)
sp.close(); one.close()
# (2) ONE ladder over the ranks: likelihood evaluation sharded, logL all-gathered by NCCL inside the library's captured
# step, swaps decided redundantly -- every rank ends with the one-GPU chain, bit for bit
sr = ShardedPT(ctx, 16, 1, logp, seed=9, npast=10)
sr.sampler.init_random()
sr.step(3)      # single steps
sr.step(37)     # CUDA-graph replays with the all-gather captured
one = PTSampler(ctx, 16, 1, logp, seed=9, npast=10)
one.init_random()
one.step(40)
ok_rungs = sr.mode == "rungs" and all(bool(np.array_equal(a, b, equal_nan=True)) for a, b in zip(sr.sampler.state(), one.state()))
ok_rungs = ok_rungs and sr.sampler.eval_shard()[1] == 8
flag = torch.tensor([int(ok_pt), int(ok_host), int(ok_rungs)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
ok_pt, ok_host, ok_rungs = (bool(v) for v in flag.tolist())
sr.close(); one.close()
if rank == 0:
    print(json.dumps({"ok_shard": ok_shard, "ok_pt": ok_pt, "ok_host": ok_host, "ok_rungs": ok_rungs}))
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
    assert res == {"ok_shard": True, "ok_pt": True, "ok_host": True, "ok_rungs": True}, res
