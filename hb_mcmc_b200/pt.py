"""Device-resident parallel-tempering sampler (host-side mirror of the hb_pt_* C ABI) and its
multi-GPU driver.

Single GPU: :class:`PTSampler` -- the step/swap loop of mcmc_wrapper2.c:378-563 with every rung's
state on the device (one likelihood evaluation per rung per step).

Multi GPU: :class:`ShardedPT` -- one process per GPU.  With at least as many ensembles (whole temperature
ladders) as ranks, ensembles are split over the ranks: swaps stay GPU-local, nothing is exchanged per step.  With
fewer (the reference's own case is ONE ladder, mcmc_wrapper2.h:11), every rank holds the whole sampler and only the
likelihood evaluation -- all of the cost -- is split over the ranks; the per-step log-likelihood vector is
all-gathered by NCCL from inside the library's captured step (hb_comm_*), and the swaps are decided identically on
every rank.  Either way the random streams are keyed on global ids: the chains do not depend on the GPU count.
``torch.distributed`` only carries the 128-byte NCCL id and the occasional log-line gather.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .lib import NPARS, Context, HBError, _f64, _p, load_library

COUNTER_NAMES = ("acc_slot0", "de_trials_slot0", "de_acc_slot0", "accepted", "proposed", "swaps_accepted",
                 "swaps_proposed", "iterations")


class PTSampler:
    def __init__(self, ctx: Context, n_temps: int, n_ens: int, log_lc_period: float, seed: int = 1,
                 dtemp: float = 1.4, npast: int = 500, quirks: bool = True, ens_offset: int = 0):
        self.ctx = ctx
        self._L = ctx._L
        self.n_temps, self.n_ens = int(n_temps), int(n_ens)
        self.n_walkers = self.n_temps * self.n_ens
        self.seed, self.dtemp, self.npast, self.quirks = int(seed), float(dtemp), int(npast), bool(quirks)
        self.log_lc_period = float(log_lc_period)
        self.ens_offset = int(ens_offset)  # global id of the first ensemble (random streams are keyed on global ids)
        h = C.c_void_p()
        rc = self._L.hb_pt_create_sharded(ctx.handle, C.byref(h), self.n_temps, self.n_ens, self.ens_offset,
                                          self.log_lc_period, C.c_ulonglong(self.seed), self.dtemp, self.npast,
                                          int(self.quirks))
        if rc != 0:
            raise HBError(self._L.hb_last_error(ctx.handle).decode() or f"hb_pt_create failed ({rc})")
        self._h = h

    def _ck(self, rc):
        if rc != 0:
            raise HBError(self._L.hb_last_error(self.ctx.handle).decode() or f"libhb_b200 error {rc}")

    def close(self):
        if getattr(self, "_h", None):
            self._L.hb_pt_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def temps(self) -> np.ndarray:
        t = np.empty(self.n_temps)
        t[0] = 1.0
        for i in range(1, self.n_temps):
            t[i] = t[i - 1] * self.dtemp  # mcmc_wrapper2.c:332-338
        return t

    @property
    def iteration(self) -> int:
        return int(self._L.hb_pt_iteration(self._h))

    def init_random(self):
        self._ck(self._L.hb_pt_init_random(self._h))

    def set_state(self, x):
        x = _f64(x).reshape(self.n_walkers, NPARS)
        self._ck(self._L.hb_pt_set_state(self._h, _p(x)))

    def step(self, n_iters: int = 1):
        self._ck(self._L.hb_pt_step(self._h, int(n_iters)))

    def set_one_launch(self, enable: bool):
        """Short light curves: run whole step loops in one launch (default) or as stream-ordered kernels."""
        self._ck(self._L.hb_pt_set_one_launch(self._h, int(enable)))

    # -- likelihood evaluation split over several samplers that hold the same state (see hb_b200.h) --------------
    def set_eval_shard(self, rank: int, world: int):
        self._ck(self._L.hb_pt_set_eval_shard(self._h, int(rank), int(world)))

    def eval_shard(self):
        """(first, count, chunk): the walkers whose likelihood this sampler evaluates, and the all-gather chunk."""
        a, b, c = C.c_long(), C.c_long(), C.c_long()
        self._ck(self._L.hb_pt_get_eval_shard(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def set_comm(self, comm: "Comm | None"):
        self._ck(self._L.hb_pt_set_comm(self._h, comm.handle if comm is not None else None))
        self._comm = comm  # keep it alive

    def step_begin(self):
        self._ck(self._L.hb_pt_step_begin(self._h))

    def step_end(self):
        self._ck(self._L.hb_pt_step_end(self._h))

    def state(self):
        """(x[W,21], logL[W] by chain slot, index[E,T] rung -> slot)."""
        x = np.empty((self.n_walkers, NPARS))
        ll = np.empty(self.n_walkers)
        idx = np.empty(self.n_walkers, dtype=np.int32)
        self._ck(self._L.hb_pt_get_state(self._h, _p(x), _p(ll), idx.ctypes.data_as(C.POINTER(C.c_int))))
        return x, ll, idx.reshape(self.n_ens, self.n_temps)

    def proposal(self):
        y = np.empty((self.n_walkers, NPARS))
        ll = np.empty(self.n_walkers)
        lp = np.empty(self.n_walkers)
        self._ck(self._L.hb_pt_get_proposal(self._h, _p(y), _p(ll), _p(lp)))
        return y, ll, lp

    def cold(self):
        x = np.empty((self.n_ens, NPARS))
        ll = np.empty(self.n_ens)
        self._ck(self._L.hb_pt_get_cold(self._h, _p(x), _p(ll)))
        return x, ll

    def logL_by_rung(self) -> np.ndarray:
        out = np.empty((self.n_ens, self.n_temps))
        self._ck(self._L.hb_pt_get_logL_by_rung(self._h, _p(out)))
        return out

    def map(self):
        x = np.empty((self.n_ens, NPARS))
        ll = np.empty(self.n_ens)
        self._ck(self._L.hb_pt_get_map(self._h, _p(x), _p(ll)))
        return x, ll

    def counters(self) -> dict:
        out = np.zeros((self.n_ens, 8), dtype=np.uint64)
        self._ck(self._L.hb_pt_get_counters(self._h, out.ctypes.data_as(C.POINTER(C.c_ulonglong))))
        return {k: out[:, i].copy() for i, k in enumerate(COUNTER_NAMES)}

    def cold_logL_into(self, device_ptr: int) -> None:
        """Cold-rung logL[n_ens] into a device buffer (async on the context's stream)."""
        self._ck(self._L.hb_pt_cold_logL_dev(self._h, C.c_void_p(device_ptr)))

    def device_logL_ptr(self) -> int:
        return int(self._L.hb_pt_device_logL(self._h) or 0)


def exchange_local(samplers) -> None:
    """The samplers of ONE process (same or different GPUs) swap their shards of the proposal logL (peer copies)."""
    L = samplers[0]._L
    arr = (C.c_void_p * len(samplers))(*[s._h for s in samplers])
    rc = L.hb_pt_exchange_local(arr, len(samplers))
    if rc != 0:
        raise HBError(L.hb_last_error(samplers[0].ctx.handle).decode() or f"hb_pt_exchange_local failed ({rc})")


class Comm:
    """NCCL communicator owned by libhb_b200 (hb_comm_*): the all-gather of the per-step logL vector is issued from C on
    the context's stream.  `exchange_id(id_bytes_or_None) -> id_bytes` carries rank 0's 128-byte id to every rank."""

    def __init__(self, device: int, rank: int, world: int, exchange_id):
        self._L = load_library()
        ident = C.create_string_buffer(128)
        if rank == 0:
            if self._L.hb_comm_unique_id(ident) != 0:
                raise HBError(self._L.hb_comm_last_error().decode())
        raw = exchange_id(bytes(ident.raw) if rank == 0 else None)
        h = C.c_void_p()
        if self._L.hb_comm_create(C.byref(h), int(device), raw, int(rank), int(world)) != 0:
            raise HBError(self._L.hb_comm_last_error().decode())
        self.handle, self.rank, self.world = h, int(rank), int(world)

    def allgather_inplace(self, device_ptr: int, count_per_rank: int, cuda_stream: int) -> None:
        if self._L.hb_comm_allgather_f64(self.handle, C.c_void_p(device_ptr), int(count_per_rank), C.c_void_p(cuda_stream)) != 0:
            raise HBError(self._L.hb_comm_last_error().decode())

    def close(self):
        if getattr(self, "handle", None):
            self._L.hb_comm_destroy(self.handle)
            self.handle = None


def shard_ensembles(n_ens: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block of ensembles owned by `rank`: (first, count).  Whole ladders stay on one
    GPU, so swaps need no communication."""
    if n_ens < world:
        raise ValueError(f"{n_ens} ensembles cannot be split over {world} ranks (whole ladders per GPU)")
    base, rem = divmod(n_ens, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def shard_walkers(n_walkers: int, world: int, rank: int) -> tuple[int, int, int]:
    """(first, count, chunk) of the walkers whose likelihood `rank` evaluates when the evaluation of one sampler is
    split (hb_pt_set_eval_shard): chunks of ceil(W / world), the last ranks may get fewer or none."""
    chunk = -(-n_walkers // world)
    first = min(rank * chunk, n_walkers)
    return first, min(chunk, n_walkers - first), chunk


def global_map(cold_logL_all: np.ndarray, cold_x_local: np.ndarray, first: int, count: int):
    """Which ensemble holds the best cold-rung logL, and whether this rank owns it."""
    best = int(np.nanargmax(cold_logL_all))
    owner_local = best - first if first <= best < first + count else None
    return best, float(cold_logL_all[best]), (cold_x_local[owner_local] if owner_local is not None else None)


class ShardedPT:
    """The sampler over the ranks of a torch.distributed process group (one process per GPU).

    mode "ensembles" (n_ens_total >= ranks): whole ladders per rank, no per-step exchange.
    mode "rungs"     (fewer ladders than ranks): every rank holds all ladders, evaluates its shard of the walkers'
                     likelihoods, and the library all-gathers the logL vector over NCCL inside its captured step."""

    def __init__(self, ctx: Context, n_temps: int, n_ens_total: int, log_lc_period: float, seed: int = 1, **kw):
        import torch.distributed as dist
        self.dist = dist
        self.ctx = ctx
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.n_ens_total = int(n_ens_total)
        self.comm = None
        if self.n_ens_total >= self.world:
            self.mode = "ensembles"
            self.first, self.count = shard_ensembles(self.n_ens_total, self.world, self.rank)
            # ONE seed: the stream ids carry the global ensemble, so rank r walks the chains the one-GPU run walks
            self.sampler = PTSampler(ctx, n_temps, self.count, log_lc_period, seed=seed, ens_offset=self.first, **kw)
        else:
            self.mode = "rungs"
            self.first, self.count = 0, self.n_ens_total
            self.sampler = PTSampler(ctx, n_temps, self.n_ens_total, log_lc_period, seed=seed, **kw)
            if self.world > 1:
                self.sampler.set_eval_shard(self.rank, self.world)
                self.comm = Comm(ctx.device, self.rank, self.world, self._exchange_id)
                self.sampler.set_comm(self.comm)

    def _exchange_id(self, raw):
        box = [raw]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]

    def step(self, n_iters: int = 1):
        self.sampler.step(n_iters)

    def close(self):
        self.sampler.close()
        if self.comm is not None:
            self.comm.close()

    def gather_cold_logL(self, device=None) -> np.ndarray:
        """Cold-rung logL of every ensemble of the job (for log lines / the global MAP; not needed by the sampler).
        Mode "rungs": every rank already holds it.  Mode "ensembles": an all-gather through torch.distributed."""
        import torch
        _, ll = self.sampler.cold()  # (synchronises the library's stream)
        if self.world == 1 or self.mode == "rungs":
            return ll
        counts = [shard_ensembles(self.n_ens_total, self.world, r)[1] for r in range(self.world)]
        mx = max(counts)
        dev = device if device is not None else ("cuda" if self.dist.get_backend() == "nccl" else "cpu")
        send = torch.full((mx,), float("nan"), dtype=torch.float64, device=dev)
        send[: self.count] = torch.from_numpy(ll).to(dev)
        recv = [torch.empty(mx, dtype=torch.float64, device=dev) for _ in range(self.world)]
        self.dist.all_gather(recv, send)
        return np.concatenate([recv[r][: counts[r]].cpu().numpy() for r in range(self.world)])
