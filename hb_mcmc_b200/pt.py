"""Device-resident parallel-tempering sampler (host-side mirror of the hb_pt_* C ABI) and its
multi-GPU driver.

Single GPU: :class:`PTSampler` -- the step/swap loop of mcmc_wrapper2.c:378-563 with every rung's
state on the device (one likelihood evaluation per rung per step).

Multi GPU: :class:`ShardedPT` -- one process per GPU (``torch.distributed``), ensembles (whole
temperature ladders) are split over the ranks so replica-exchange swaps stay GPU-local and chain
state never moves; the only exchange is an all-gather of the per-step cold-rung log-likelihood
vector (8 B per ensemble) used for the global MAP / log lines (NCCL over NVLink on GPUs, gloo in
the CPU tests of the sharding logic).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .lib import NPARS, Context, HBError, _f64, _p

COUNTER_NAMES = ("acc_slot0", "de_trials_slot0", "de_acc_slot0", "accepted", "proposed", "swaps_accepted",
                 "swaps_proposed", "iterations")


class PTSampler:
    def __init__(self, ctx: Context, n_temps: int, n_ens: int, log_lc_period: float, seed: int = 1,
                 dtemp: float = 1.4, npast: int = 500, quirks: bool = True):
        self.ctx = ctx
        self._L = ctx._L
        self.n_temps, self.n_ens = int(n_temps), int(n_ens)
        self.n_walkers = self.n_temps * self.n_ens
        self.seed, self.dtemp, self.npast, self.quirks = int(seed), float(dtemp), int(npast), bool(quirks)
        self.log_lc_period = float(log_lc_period)
        h = C.c_void_p()
        rc = self._L.hb_pt_create(ctx.handle, C.byref(h), self.n_temps, self.n_ens, self.log_lc_period,
                                  C.c_ulonglong(self.seed), self.dtemp, self.npast, int(self.quirks))
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


def shard_ensembles(n_ens: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block of ensembles owned by `rank`: (first, count).  Whole ladders stay on one
    GPU, so swaps need no communication."""
    if n_ens < world:
        raise ValueError(f"{n_ens} ensembles cannot be split over {world} ranks (whole ladders per GPU)")
    base, rem = divmod(n_ens, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def global_map(cold_logL_all: np.ndarray, cold_x_local: np.ndarray, first: int, count: int):
    """Which ensemble holds the best cold-rung logL, and whether this rank owns it."""
    best = int(np.nanargmax(cold_logL_all))
    owner_local = best - first if first <= best < first + count else None
    return best, float(cold_logL_all[best]), (cold_x_local[owner_local] if owner_local is not None else None)


class ShardedPT:
    """Ensembles split over the ranks of a torch.distributed process group."""

    def __init__(self, ctx: Context, n_temps: int, n_ens_total: int, log_lc_period: float, seed: int = 1, **kw):
        import torch.distributed as dist
        self.dist = dist
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.first, self.count = shard_ensembles(n_ens_total, self.world, self.rank)
        self.n_ens_total = n_ens_total
        # distinct Philox streams per rank: the stream id carries the GLOBAL ensemble through the seed
        self.sampler = PTSampler(ctx, n_temps, self.count, log_lc_period, seed=seed + 0x9E3779B97F4A7C15 * self.rank, **kw)

    def step(self, n_iters: int = 1):
        self.sampler.step(n_iters)

    def gather_cold_logL_device(self):
        """Device-side variant: the gather kernel writes straight into the NCCL send buffer; returns
        the gathered [world, max_count] tensor (padding is NaN) without a host round trip."""
        import torch
        counts = [shard_ensembles(self.n_ens_total, self.world, r)[1] for r in range(self.world)]
        mx = max(counts)
        if getattr(self, "_send", None) is None:
            self._send = torch.full((mx,), float("nan"), dtype=torch.float64, device="cuda")
            self._recv = torch.empty((self.world, mx), dtype=torch.float64, device="cuda")
        self.sampler.cold_logL_into(self._send.data_ptr())
        if self.world == 1:
            self._recv.copy_(self._send[None])
        else:
            self.dist.all_gather_into_tensor(self._recv, self._send)
        return self._recv

    def gather_cold_logL(self, device=None) -> np.ndarray:
        """All-gather of the cold-rung logL vector (the only per-step exchange)."""
        import torch
        _, ll = self.sampler.cold()
        if self.world == 1:
            return ll
        counts = [shard_ensembles(self.n_ens_total, self.world, r)[1] for r in range(self.world)]
        mx = max(counts)
        dev = device if device is not None else ("cuda" if self.dist.get_backend() == "nccl" else "cpu")
        send = torch.full((mx,), float("nan"), dtype=torch.float64, device=dev)
        send[: self.count] = torch.from_numpy(ll).to(dev)
        recv = [torch.empty(mx, dtype=torch.float64, device=dev) for _ in range(self.world)]
        self.dist.all_gather(recv, send)
        return np.concatenate([recv[r][: counts[r]].cpu().numpy() for r in range(self.world)])
