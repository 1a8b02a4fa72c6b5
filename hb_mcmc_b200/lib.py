"""ctypes binding of libhb_b200.so (the C ABI declared in include/hb_b200.h).

The library is the product; this module only marshals numpy buffers into it.  It fails
loudly (:class:`HBError`) when the library is missing or no sm_100 device is usable -- there
is no CPU fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

NPARS = 21
BIG_NUM = 1.0e15

_dp = C.POINTER(C.c_double)


class HBError(RuntimeError):
    pass


# every symbol include/hb_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = (
    "hb_create", "hb_destroy", "hb_last_error", "hb_global_error", "hb_device_info", "hb_set_stream", "hb_sync",
    "hb_set_data", "hb_set_mags", "hb_loglikelihood_batch", "hb_loglikelihood_batch_dev", "hb_light_curve_batch",
    "hb_calc_light_curve", "hb_chain_info_batch", "hb_traj", "hb_order_statistic", "hb_remove_median", "hb_scalar", "hb_gaia_batch", "hb_fp64_peak",
    "hb_time_kernels", "hb_last_eval_kernel_ms", "hb_launch_count", "hb_set_bracket_sigma", "hb_set_sincos_range",
    "hb_set_max_parts", "hb_evaluated_chains",
    "hb_pt_create", "hb_pt_destroy", "hb_pt_init_random", "hb_pt_set_state", "hb_pt_step", "hb_pt_iteration",
    "hb_pt_get_state", "hb_pt_get_proposal", "hb_pt_get_cold", "hb_pt_get_logL_by_rung", "hb_pt_get_map",
    "hb_pt_get_counters", "hb_pt_device_logL", "hb_pt_cold_logL_dev",
    "hb_pt_create_sharded", "hb_pt_set_eval_shard", "hb_pt_get_eval_shard", "hb_pt_set_comm", "hb_pt_step_begin",
    "hb_pt_exchange_local", "hb_pt_step_end", "hb_pt_set_one_launch",
    "hb_comm_unique_id", "hb_comm_create", "hb_comm_create_all", "hb_comm_destroy", "hb_comm_rank", "hb_comm_world",
    "hb_comm_nccl_version", "hb_comm_last_error", "hb_comm_allgather_f64",
    "hb_gaia_pt_create", "hb_gaia_pt_destroy", "hb_gaia_pt_set_data", "hb_gaia_pt_set_sigma", "hb_gaia_pt_init_random",
    "hb_gaia_pt_set_state", "hb_gaia_pt_records", "hb_gaia_pt_run", "hb_gaia_pt_iteration", "hb_gaia_pt_get_state",
    "hb_gaia_pt_get_proposal", "hb_gaia_pt_get_history", "hb_gaia_pt_get_map", "hb_gaia_pt_get_counters",
)

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen libhb_b200.so (building it first if the sources are newer) and set prototypes."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("HB_B200_LIB") or _build.LIB  # (HB_B200_LIB: a -D variant of the library, for tools/)
    if not os.path.exists(path):
        try:
            _build.build_lib()
        except Exception as exc:  # pragma: no cover - build environment problem
            raise HBError(f"libhb_b200.so is missing and could not be built: {exc}") from exc
    try:
        L = C.CDLL(path)
    except OSError as exc:
        raise HBError(f"cannot load {path}: {exc}") from exc
    vp, i, l, d = C.c_void_p, C.c_int, C.c_long, C.c_double
    L.hb_create.argtypes = [C.POINTER(vp), i]
    L.hb_destroy.argtypes = [vp]
    L.hb_destroy.restype = None
    L.hb_last_error.argtypes = [vp]
    L.hb_last_error.restype = C.c_char_p
    L.hb_global_error.restype = C.c_char_p
    L.hb_device_info.argtypes = [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(l)]
    L.hb_set_stream.argtypes = [vp, vp]
    L.hb_sync.argtypes = [vp]
    L.hb_set_data.argtypes = [vp, _dp, _dp, _dp, l]
    L.hb_set_mags.argtypes = [vp, _dp, _dp, i, i]
    L.hb_loglikelihood_batch.argtypes = [vp, _dp, l, _dp]
    L.hb_loglikelihood_batch_dev.argtypes = [vp, vp, l, vp]
    L.hb_light_curve_batch.argtypes = [vp, _dp, l, _dp]
    L.hb_calc_light_curve.argtypes = [vp, _dp, l, _dp, _dp]
    L.hb_chain_info_batch.argtypes = [vp, _dp, l, d, _dp]
    L.hb_traj.argtypes = [vp, _dp, l, _dp] + [_dp] * 5
    L.hb_scalar.argtypes = [vp, i, _dp, i, _dp]
    L.hb_order_statistic.argtypes = [vp, _dp, l, l, _dp]
    L.hb_remove_median.argtypes = [vp, _dp, l]
    L.hb_gaia_batch.argtypes = [vp, _dp, l, d, _dp, _dp, _dp, _dp]
    L.hb_fp64_peak.argtypes = [vp, d, _dp]
    L.hb_set_bracket_sigma.argtypes = [vp, d]
    L.hb_set_sincos_range.argtypes = [vp, d]
    L.hb_set_max_parts.argtypes = [vp, i]
    L.hb_evaluated_chains.argtypes = [vp, C.POINTER(C.c_ulonglong), i]
    L.hb_time_kernels.argtypes = [vp, i]
    L.hb_last_eval_kernel_ms.argtypes = [vp, _dp]
    L.hb_launch_count.argtypes = [vp]
    L.hb_launch_count.restype = l
    ull = C.c_ulonglong
    L.hb_pt_create.argtypes = [vp, C.POINTER(vp), i, i, d, ull, d, i, i]
    L.hb_pt_destroy.argtypes = [vp]
    L.hb_pt_destroy.restype = None
    L.hb_pt_init_random.argtypes = [vp]
    L.hb_pt_set_state.argtypes = [vp, _dp]
    L.hb_pt_step.argtypes = [vp, l]
    L.hb_pt_iteration.argtypes = [vp]
    L.hb_pt_iteration.restype = l
    L.hb_pt_get_state.argtypes = [vp, _dp, _dp, C.POINTER(i)]
    L.hb_pt_get_proposal.argtypes = [vp, _dp, _dp, _dp]
    L.hb_pt_get_cold.argtypes = [vp, _dp, _dp]
    L.hb_pt_get_logL_by_rung.argtypes = [vp, _dp]
    L.hb_pt_get_map.argtypes = [vp, _dp, _dp]
    L.hb_pt_get_counters.argtypes = [vp, C.POINTER(ull)]
    L.hb_pt_device_logL.argtypes = [vp]
    L.hb_pt_device_logL.restype = vp
    L.hb_pt_cold_logL_dev.argtypes = [vp, vp]
    L.hb_pt_create_sharded.argtypes = [vp, C.POINTER(vp), i, i, i, d, ull, d, i, i]
    L.hb_pt_set_eval_shard.argtypes = [vp, i, i]
    L.hb_pt_get_eval_shard.argtypes = [vp, C.POINTER(l), C.POINTER(l), C.POINTER(l)]
    L.hb_pt_set_comm.argtypes = [vp, vp]
    L.hb_pt_step_begin.argtypes = [vp]
    L.hb_pt_exchange_local.argtypes = [C.POINTER(vp), i]
    L.hb_pt_step_end.argtypes = [vp]
    L.hb_pt_set_one_launch.argtypes = [vp, i]
    L.hb_comm_unique_id.argtypes = [C.c_char_p]
    L.hb_comm_create.argtypes = [C.POINTER(vp), i, C.c_char_p, i, i]
    L.hb_comm_create_all.argtypes = [C.POINTER(vp), C.POINTER(i), i]
    L.hb_comm_destroy.argtypes = [vp]
    L.hb_comm_destroy.restype = None
    L.hb_comm_rank.argtypes = [vp]
    L.hb_comm_world.argtypes = [vp]
    L.hb_comm_nccl_version.argtypes = [C.POINTER(i)]
    L.hb_comm_last_error.restype = C.c_char_p
    L.hb_comm_allgather_f64.argtypes = [vp, vp, l, vp]
    ip = C.POINTER(i)
    L.hb_gaia_pt_create.argtypes = [vp, C.POINTER(vp), i, i, ull, d, i]
    L.hb_gaia_pt_destroy.argtypes = [vp]
    L.hb_gaia_pt_destroy.restype = None
    L.hb_gaia_pt_set_data.argtypes = [vp, _dp, _dp, _dp]
    L.hb_gaia_pt_set_sigma.argtypes = [vp, _dp]
    L.hb_gaia_pt_init_random.argtypes = [vp]
    L.hb_gaia_pt_set_state.argtypes = [vp, _dp]
    L.hb_gaia_pt_records.argtypes = [vp, l, i]
    L.hb_gaia_pt_records.restype = l
    L.hb_gaia_pt_run.argtypes = [vp, l, i, _dp, _dp]
    L.hb_gaia_pt_iteration.argtypes = [vp]
    L.hb_gaia_pt_iteration.restype = l
    L.hb_gaia_pt_get_state.argtypes = [vp, _dp, _dp, ip]
    L.hb_gaia_pt_get_proposal.argtypes = [vp, _dp, _dp, _dp, ip]
    L.hb_gaia_pt_get_history.argtypes = [vp, _dp]
    L.hb_gaia_pt_get_map.argtypes = [vp, _dp, _dp]
    L.hb_gaia_pt_get_counters.argtypes = [vp, C.POINTER(ull)]
    if path == _build.LIB:
        _lib = L
    return L


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a: np.ndarray):
    return a.ctypes.data_as(_dp)


class Context:
    """One device context: the uploaded light curve plus the batched likelihood on it."""

    def __init__(self, device: int = 0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.hb_create(C.byref(h), int(device))
        if rc != 0:
            raise HBError(self._L.hb_global_error().decode() or f"hb_create failed ({rc})")
        self._h = h
        self.device = int(device)
        self.n_points = 0

    # -- plumbing ---------------------------------------------------------------------
    def _ck(self, rc: int) -> None:
        if rc != 0:
            raise HBError(self._L.hb_last_error(self._h).decode() or f"libhb_b200 error {rc}")

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.hb_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def device_info(self) -> dict:
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_long()
        self._ck(self._L.hb_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "global_mem_mb": mem.value}

    def set_stream(self, cuda_stream: int | None) -> None:
        self._ck(self._L.hb_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def sync(self) -> None:
        self._ck(self._L.hb_sync(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._L.hb_launch_count(self._h))

    # -- data -------------------------------------------------------------------------
    def set_data(self, t, flux, err) -> None:
        t, flux, err = _f64(t), _f64(flux), _f64(err)
        if not (t.shape == flux.shape == err.shape and t.ndim == 1):
            raise ValueError("t, flux, err must be 1-D arrays of equal length")
        self._ck(self._L.hb_set_data(self._h, _p(t), _p(flux), _p(err), t.size))
        self.n_points = t.size

    def set_mags(self, mag_data, magerr, use_gmag: int = 1, use_color: int = 0) -> None:
        md, me = _f64(mag_data), _f64(magerr)
        if md.size != 5 or me.size != 4:
            raise ValueError("mag_data needs 5 entries, magerr 4")
        self._ck(self._L.hb_set_mags(self._h, _p(md), _p(me), int(use_gmag), int(use_color)))

    # -- hot path -----------------------------------------------------------------------
    def loglikelihood(self, params) -> np.ndarray:
        """logL for params[n, 21] (host buffers; copies are inside the call)."""
        P = _f64(params).reshape(-1, NPARS)
        out = np.empty(P.shape[0])
        self._ck(self._L.hb_loglikelihood_batch(self._h, _p(P), P.shape[0], _p(out)))
        return out

    def loglikelihood_into(self, P: np.ndarray, out: np.ndarray) -> None:
        """Same without allocations: P must be C-contiguous float64 [n, 21]."""
        self._ck(self._L.hb_loglikelihood_batch(self._h, _p(P), P.shape[0], _p(out)))

    def loglikelihood_dev(self, d_params_ptr: int, n_chains: int, d_logL_ptr: int) -> None:
        """Device buffers (raw pointers, e.g. torch ``data_ptr()``), asynchronous on the stream."""
        self._ck(self._L.hb_loglikelihood_batch_dev(self._h, C.c_void_p(d_params_ptr), int(n_chains),
                                                    C.c_void_p(d_logL_ptr)))

    def light_curves(self, params) -> np.ndarray:
        P = _f64(params).reshape(-1, NPARS)
        out = np.empty((P.shape[0], self.n_points))
        self._ck(self._L.hb_light_curve_batch(self._h, _p(P), P.shape[0], _p(out)))
        return out

    def calc_light_curve(self, times, pars) -> np.ndarray:
        t, p = _f64(times), _f64(pars)
        if p.size != NPARS:
            raise ValueError("pars needs 21 entries")
        out = np.empty(t.size)
        self._ck(self._L.hb_calc_light_curve(self._h, _p(t), t.size, _p(p), _p(out)))
        return out

    # -- helpers ------------------------------------------------------------------------
    def chain_info(self, params, D: float = 1000.0) -> np.ndarray:
        """[n, 9] = R1 R2 Teff1 Teff2 G B-V V-G G-T RocheOverflow."""
        P = _f64(params).reshape(-1, NPARS)
        out = np.empty((P.shape[0], 9))
        self._ck(self._L.hb_chain_info_batch(self._h, _p(P), P.shape[0], float(D), _p(out)))
        return out

    def roche_overflow(self, params) -> np.ndarray:
        return self.chain_info(params)[:, 8].astype(np.int32)

    def traj(self, times, traj_pars) -> dict:
        t, tp = _f64(times), _f64(traj_pars)
        outs = [np.empty(t.size) for _ in range(5)]
        self._ck(self._L.hb_traj(self._h, _p(t), t.size, _p(tp), *[_p(o) for o in outs]))
        return dict(zip(("d", "Z1", "Z2", "r", "nu"), outs))

    def order_statistic(self, x, k: int) -> float:
        a = _f64(x)
        out = C.c_double()
        self._ck(self._L.hb_order_statistic(self._h, _p(a), a.size, int(k), C.byref(out)))
        return out.value

    def remove_median(self, arr) -> np.ndarray:
        a = _f64(arr).copy()
        self._ck(self._L.hb_remove_median(self._h, _p(a), a.size))
        return a

    def scalar(self, op: int, *args: float) -> float:
        a = _f64(args)
        out = np.empty(1)
        self._ck(self._L.hb_scalar(self._h, int(op), _p(a), a.size, _p(out)))
        return float(out[0])

    def gaia(self, p6, D: float, data=None, err=None):
        p = _f64(p6).reshape(-1, 6)
        mags = np.empty((p.shape[0], 4))
        if data is None:
            self._ck(self._L.hb_gaia_batch(self._h, _p(p), p.shape[0], float(D), None, None, _p(mags), None))
            return mags
        dd, ee = _f64(data), _f64(err)
        ll = np.empty(p.shape[0])
        self._ck(self._L.hb_gaia_batch(self._h, _p(p), p.shape[0], float(D), _p(dd), _p(ee), _p(mags), _p(ll)))
        return mags, ll

    def set_bracket_sigma(self, sigma: float) -> None:
        """Half-width of the pre-sample median bracket (cost knob; results do not depend on it)."""
        self._ck(self._L.hb_set_bracket_sigma(self._h, float(sigma)))

    def set_sincos_range(self, max_abs: float) -> None:
        """|E| above which the logL-only pass re-evaluates a chain with the libm fallback (test knob)."""
        self._ck(self._L.hb_set_sincos_range(self._h, float(max_abs)))

    def set_max_parts(self, max_parts: int) -> None:
        """Most CTAs one light curve may be spread over in small batches (latency knob; results do not depend on it)."""
        self._ck(self._L.hb_set_max_parts(self._h, int(max_parts)))

    def evaluated_chains(self, reset: bool = False) -> int:
        """Chains whose model was really evaluated since the last reset (Roche / e >= 1 early-outs are not counted)."""
        out = C.c_ulonglong()
        self._ck(self._L.hb_evaluated_chains(self._h, C.byref(out), int(reset)))
        return int(out.value)

    def time_kernels(self, enable: bool = True) -> None:
        self._ck(self._L.hb_time_kernels(self._h, int(enable)))

    def last_eval_kernel_ms(self) -> float:
        out = C.c_double()
        self._ck(self._L.hb_last_eval_kernel_ms(self._h, C.byref(out)))
        return out.value

    def fp64_peak_tflops(self, seconds: float = 0.3) -> float:
        out = C.c_double()
        self._ck(self._L.hb_fp64_peak(self._h, float(seconds), C.byref(out)))
        return out.value
