#!/usr/bin/env python
"""bench.py -- the hot-path benchmark of hb_mcmc_b200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: the batched light-curve model +
log-likelihood (likelihood3.c:809-873) for `n_chains` parameter vectors on the resident light
curve.  Workload at any N: BASELINE.json configs[1] per GPU (4096 chains x 20 000 TESS 2-min
points, truth A of test_likelihoods.c:33-36, prior draws with Roche-overflow draws rejected) --
chains are independent, so ranks shard them with no data-path collective ("scaling": "weak").

One JSON line on stdout (rank 0):
  value      model-point logL evals/s, whole job, inputs resident in HBM (device-buffer C ABI),
             CUDA-event time per step, L2 flushed between steps, max over ranks
  e2e        same metric through hb_loglikelihood_batch with HOST buffers (H2D of the parameter
             batch from pinned memory + D2H of logL inside every step)
  roofline   FP64 CUDA-core roofline of k_chain_eval on EXECUTED work: FP64 flop per model point as counted by
             the committed `ncu --set full` capture of this kernel (2 x DFMA + DMUL + DADD, profiles/) x points /
             kernel duration (CUDA events on the launching stream) against the DFMA peak measured on this box by
             hb_fp64_peak in the same run: `frac` <= 1.  `frac_vs_reference_formulation` is the same time set
             against the 520 flop / point of the REFERENCE's formulation -- it exceeds 1 because
             the kernel reaches the reference's numbers with a quarter of its operations
  extra      the other BASELINE configs in the same line: C1 latency, C3's per-GPU share, C4 (+Gaia), C5's share;
             at N > 1 C3 as named (64 x 256 walkers, strong scaling), one ladder of 64 rungs x 200 000 points with
             the rungs split over the GPUs, and `shard_bit_identical` (sharded logL == full logL on a 512-chain probe)
  cpu_baseline  the reference's own loglikelihood() (oracle/_ref, or the oracle port when the
             compiled reference is absent) on all host cores over a bounded sample
`--impl reference` times that CPU path alone on the same config and prints the same line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from hb_mcmc_b200 import workload as wl  # noqa: E402

METRIC = "model_point_logL_evals_per_sec"
UNIT = "points/s"
FLOP_PER_POINT = 520.0  # 156 plain + 6 sincos x 40 + 8 div x 14 + 1 sqrt x 14
NOMINAL_FP64_TFLOPS = 37.2  # 148 SM x 64 lanes x 2 x 1.965 GHz
CHAIN_CONST_BYTES = 47 * 8  # sizeof(ChainConst): what k_chain_eval reads per chain
# From the committed `ncu --set full` capture of one k_chain_eval launch on C2 (profiles/r2_chain_eval_ncu_summary.txt,
# tools/ncu_mix.py); only meaningful for the default workload:
NCU_TRAFFIC_C2_BYTES = 2.806016e6 + 18.314496e6  # dram__bytes_read.sum + dram__bytes_write.sum
NCU_EXECUTED = {
    "fp64_instr_per_point": 78.7,   # DFMA 48.4 + DMUL 16.6 + DADD 10.0 + DSETP 3.8 (warp instructions / 32 samples)
    "flop_per_point": 123.4,        # 2 x DFMA + DMUL + DADD
    "all_instr_per_point": 186.0,
    "fp64_pipe_active": 0.562,      # sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active
    "issue_active": 0.664,          # smsp__issue_active.avg.pct_of_peak_sustained_active
    "source": "profiles/r2_chain_eval_ncu_summary.txt",
}


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback of B200_PROFILING.md"



def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(wl.CONFIGS))
    ap.add_argument("--chains", type=int, default=0, help="override chains per GPU")
    ap.add_argument("--points", type=int, default=0, help="override points per light curve")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="skip the L2 flush between steps (diagnostic)")
    ap.add_argument("--pt-steps", type=int, default=240, help="PT-MCMC iterations timed for the steps/s figure (0 = skip)")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra legs (other BASELINE configs, multi-GPU probes)")
    ap.add_argument("--no-pt-reference", action="store_true",
                    help="skip the reference-size PT comparison (50 rungs x 375 real points, GPU vs the reference driver binary)")
    return ap.parse_args()


def workload_spec(args):
    cfg = dict(wl.CONFIGS[args.workload])
    if args.workload in ("C3", "C5"):
        # multi-GPU configs name a global chain count over 8 GPUs; per-GPU share is fixed (weak scaling)
        cfg["n_chains"] //= 8
    if args.chains:
        cfg["n_chains"] = args.chains
    if args.points:
        cfg["n_points"] = args.points
    cfg["truth_vec"] = wl.TRUTH_A if cfg["truth"] == "A" else wl.TRUTH_B
    return cfg


def gaia_setup(truth, G_truth):
    # mag_data = {D=100, G(truth)+0.02, 1,1,1}, magerr = {0.05, BIG x3}, USE_GMAG only
    return np.array([100.0, G_truth + 0.02, 1.0, 1.0, 1.0]), np.array([0.05, 1e15, 1e15, 1e15])


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.dev = device_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                 str(self.dev)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout: float = 8.0):
        """nvidia-smi needs up to a few seconds for its first line on a fresh box: do not start the timed region
        (which may last only a fraction of a second) before the sampler delivers."""
        t_end = time.perf_counter() + timeout
        while self.proc is not None and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.02)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, t0: float, t1: float) -> dict:
        sm, smax, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = [r for r in self.rows if t0 <= r[0] <= t1 + 0.05]
        if not rows:  # a timed region shorter than the sampling period: the samples around it
            rows = [r for r in self.rows if t0 - 0.5 <= r[0] <= t1 + 0.5]
        for ts, line in rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU reference arm
# --------------------------------------------------------------------------------------------
def cpu_backend():
    import oracle
    oracle.build(ref=True)
    if oracle.have_reference():
        return oracle.Reference(), "reference"
    return oracle.Oracle(), "port"


def cpu_inputs(cfg, backend):
    """Dataset + parameter draws for the CPU arm, made with the CPU backend itself."""
    t, flux, err = wl.make_dataset(cfg["n_points"], cfg["truth_vec"], backend.calc_light_curve)
    roche = lambda P: np.array([backend.roche_overflow(p) for p in P])  # noqa: E731
    return t, flux, err, roche


def run_cpu_sample(backend, t, flux, err, P, mags, threads):
    t0 = time.perf_counter()
    out = backend.loglikelihood_batch(t, flux, err, P, mag_data=mags[0], magerr=mags[1], nthreads=threads)
    dt = time.perf_counter() - t0
    return dt, out


def reference_arm(args, cfg, rank):
    if rank != 0:
        return
    backend, kind = cpu_backend()
    cores = os.cpu_count() or 1
    N = cfg["n_points"]
    t, flux, err, roche = cpu_inputs(cfg, backend)
    mags = (None, None)
    if cfg["gaia"]:
        mags = gaia_setup(cfg["truth_vec"], backend.calc_mags(cfg["truth_vec"], 100.0)[0])
    # bounded sample per step: sized from a one-chain-per-core probe so that K steps stay within ~2 min
    P_all = wl.draw_chains(max(4 * cores, 64), cfg["truth_vec"], roche, seed=1)
    probe_dt, _ = run_cpu_sample(backend, t, flux, err, P_all[:cores], mags, cores)
    per_round = max(probe_dt, 1e-3)
    step_target = min(0.5, 120.0 / max(args.steps + args.warmup, 1))
    rounds = max(1, int(step_target / per_round))
    n_step = rounds * cores
    P = wl.draw_chains(n_step, cfg["truth_vec"], roche, seed=1)
    for _ in range(args.warmup):
        run_cpu_sample(backend, t, flux, err, P, mags, cores)
    total = 0.0
    for _ in range(args.steps):
        dt, _ = run_cpu_sample(backend, t, flux, err, P, mags, cores)
        total += dt
    ms = total / args.steps * 1e3
    value = n_step * N / (ms * 1e-3)
    sample = f"{n_step} of {cfg['n_chains']} chains x {N} points per step, {cores} threads (dynamic schedule)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_block(args, cfg, extra={"cpu_step": sample}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_block(args, cfg, extra=None):
    c = {
        "workload": f"{args.workload}: {cfg['n_chains']} chains x {cfg['n_points']} points per GPU (TESS 2-min cadence), "
                    f"truth {cfg['truth']}, prior draws with Roche-overflow draws rejected"
                    + (", Gaia G-mag term" if cfg["gaia"] else ""),
        "n_chains_per_gpu": cfg["n_chains"], "n_points": cfg["n_points"], "n_pars": 21,
        "sharding": "chains split over ranks, no data-path collective",
        "l2": "flushed between timed steps (512 MiB write)" if not args.no_flush else "not flushed",
    }
    if extra:
        c.update(extra)
    return c


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def gpu_arm(args, cfg, rank, local_rank, world):
    import torch

    import hb_mcmc_b200 as hb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hb_mcmc_b200 path has no CPU fallback")
    # stdout carries exactly one JSON line: anything native libraries print (NCCL's version banner
    # goes to fd 1) is sent to stderr until the result line is written
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = hb.Context(local_rank)
    # a real (non-NULL) torch stream, made current, so torch.cuda.Event and the library's launches
    # share one stream (a NULL handle would put the library back on its private stream)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    N, n = cfg["n_points"], cfg["n_chains"]
    truth = cfg["truth_vec"]

    # identical synthetic data on every rank (generated by the GPU library itself), own chain shard
    t, flux, err = wl.make_dataset(N, truth, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    if cfg["gaia"]:
        md, me = gaia_setup(truth, ctx.chain_info(truth[None], 100.0)[0, 4])
        ctx.set_mags(md, me, 1, 0)
    P = wl.draw_chains(n, truth, ctx.roche_overflow, seed=1 + rank)

    d_params = torch.from_numpy(P).to("cuda")
    d_logL = torch.empty(n, dtype=torch.float64, device="cuda")
    flush = None if args.no_flush else torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    peak_tf = ctx.fp64_peak_tflops(0.5)

    def step_dev():
        ctx.loglikelihood_dev(d_params.data_ptr(), n, d_logL.data_ptr())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    for _ in range(max(args.warmup, 3)):
        step_dev()
    torch.cuda.synchronize()
    ctx.time_kernels(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
        time.sleep(0.15)
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    kernel_ms = []
    launches0 = ctx.launch_count
    barrier()
    t_begin = time.perf_counter()
    for k in range(args.steps):
        if flush is not None:
            flush.fill_(k & 0xFF)  # evicts the 126 MB L2; outside the per-step events
        ev0[k].record(stream)
        step_dev()
        ev1[k].record(stream)
        kernel_ms.append(ctx.last_eval_kernel_ms())  # waits for this step's kernel
    barrier()
    t_end = time.perf_counter()
    launches = ctx.launch_count - launches0
    ctx.time_kernels(False)
    step_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    total_ms = float(sum(step_ms))
    if dist is not None:
        tt = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    if rank == 0:
        time.sleep(0.1)
        sampler.stop()
    logL_dev = d_logL.cpu().numpy()

    # ---- end to end through the host-buffer C ABI ("e2e") ----
    # page-locked host buffers, as the contract asks.  The library moves them without a copy engine: k_prologue reads the
    # parameter array over the bus (one coalesced read per chain, behind the other warps' libm work) and k_chain_eval
    # writes logL into the result array; the same bytes cross the bus inside the timed region (HB_ZERO_COPY=0 in the
    # environment selects cudaMemcpyAsync both ways: 25 us per step slower)
    P_pin = torch.from_numpy(P).clone().pin_memory()
    out_pin = torch.empty(n, dtype=torch.float64).pin_memory()
    P, out = P_pin.numpy(), out_pin.numpy()
    for _ in range(2):
        ctx.loglikelihood_into(P, out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.loglikelihood_into(P, out)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    assert np.array_equal(out, logL_dev, equal_nan=True), "host-buffer and device-buffer paths disagree"

    pt_info = pt_leg(args, ctx, cfg, rank, world, dist, stream) if args.pt_steps > 0 else None
    extra = None
    if not args.no_extra:
        extra = {}
        if world > 1:
            extra["shard_bit_identical"] = shard_probe(ctx, rank, world, dist)
        if args.pt_steps > 0:
            extra["pt_one_ladder_64_rungs_x_200k"] = rung_split_leg(ctx, rank, world, dist, stream, max(args.pt_steps, 8))
        if world == 1:
            extra.update(extra_legs(ctx, stream))
        # back to the headline data set (the reference-size leg below sets its own)
        ctx.set_data(t, flux, err)

    if pt_info is not None and rank == 0 and world == 1 and not args.no_pt_reference and not args.no_cpu_baseline:
        pt_info["reference_size"] = pt_reference_size_leg(ctx)

    if rank == 0:
        ms_per_step = total_ms / args.steps
        pts_per_step = float(n) * N * world
        value = pts_per_step / (ms_per_step * 1e-3)
        e2e_value = pts_per_step / (e2e_s / args.steps)
        k_ms = float(np.mean(kernel_ms))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(args, cfg),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(P.nbytes), "d2h_bytes_per_step": int(out.nbytes),
                    "transfer": ("cudaMemcpyAsync both ways" if os.environ.get("HB_ZERO_COPY", "1")[:1] == "0" else
                                 "in place: the kernels read the page-locked parameters and write the page-locked results over the bus")},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(t_begin, t_end),
            "roofline": roofline_block(args, n, N, k_ms, ms_per_step, peak_tf),
            "nan_fraction": float(np.isnan(logL_dev).mean()),
        }
        if pt_info is not None:
            line["pt"] = pt_info
        if extra is not None:
            line["extra"] = extra
            if "shard_bit_identical" in extra:
                line["shard_bit_identical"] = extra["shard_bit_identical"]
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_leg(cfg, t, flux, err, P, logL_dev)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def hbm_block(n, N, k_ms):
    """Secondary roofline: the kernel's algorithmic HBM stream (per-chain constants in, logL out, the
    light curve once) against the measured copy bandwidth.  It is nowhere near the bound: the
    template never leaves the chip."""
    peak, src = measured_hbm_peak()
    nbytes = float(n) * (CHAIN_CONST_BYTES + 8) + 24.0 * N
    gbs = nbytes / (k_ms * 1e-3) * 1e-9
    return {"achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "bytes_per_launch": nbytes,
            "bytes_per_point": nbytes / (float(n) * N), "peak_source": src}


def roofline_block(args, n, N, k_ms, ms_per_step, peak_tf):
    """FP64 roofline of k_chain_eval on executed work (see the module docstring)."""
    default = args.workload == "C2" and not args.chains and not args.points
    ex = dict(NCU_EXECUTED)
    pts_per_s = float(n) * N / (k_ms * 1e-3)
    achieved = pts_per_s * ex["flop_per_point"] * 1e-12
    formulation = pts_per_s * FLOP_PER_POINT * 1e-12
    return {
        "bound": "fp64", "kernel": "k_chain_eval", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
        "frac": achieved / peak_tf if peak_tf > 0 else None,
        "flop_per_point": ex["flop_per_point"],
        "flop_per_point_source": ex["source"] + " (2 x DFMA + DMUL + DADD per model point, counted by ncu on C2"
                                 + ("" if default else "; this workload was not captured separately") + ")",
        "frac_vs_reference_formulation": formulation / peak_tf if peak_tf > 0 else None,
        "reference_formulation_flop_per_point": FLOP_PER_POINT,
        "traffic": NCU_TRAFFIC_C2_BYTES if default else None,
        "traffic_unit": "bytes per launch (ncu dram__bytes_read+write, " + ex["source"] + ")",
        "hbm": hbm_block(n, N, k_ms),
        "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_per_step,
        "executed": ex if default else None,
        "peak_source": "hb_fp64_peak DFMA probe on this GPU in this run (MEASURED_PEAKS.json has no FP64 entry); "
                       f"nominal {NOMINAL_FP64_TFLOPS} TFLOP/s",
        "frac_of_nominal": achieved / NOMINAL_FP64_TFLOPS,
    }


def time_batch(ctx, stream, n, N, truth, gaia, reps, seed=1, e_max=0.95):
    """One BASELINE configuration, device-resident: mean CUDA-event time per call of hb_loglikelihood_batch_dev."""
    import torch
    t, flux, err = wl.make_dataset(N, truth, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    if gaia:
        md, me = gaia_setup(truth, ctx.chain_info(truth[None], 100.0)[0, 4])
        ctx.set_mags(md, me, 1, 0)
    else:
        ctx.set_mags([1000, 1, 1, 1, 1], [1e15] * 4, 1, 0)
    P = wl.draw_chains(n, truth, ctx.roche_overflow, seed=seed, e_max=e_max)
    dP = torch.from_numpy(P).to("cuda")
    dL = torch.empty(n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        ctx.loglikelihood_dev(dP.data_ptr(), n, dL.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        ctx.loglikelihood_dev(dP.data_ptr(), n, dL.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out = dL.cpu().numpy()
    return {"n_chains": n, "n_points": N, "ms": ms, "points_per_sec": n * N / (ms * 1e-3), "nan_fraction": float(np.isnan(out).mean()),
            "reps": reps}


def extra_legs(ctx, stream):
    """The BASELINE configs the headline does not carry, on this GPU (C3 / C5: the per-GPU share of the 8-GPU job)."""
    out = {}
    out["C1_latency"] = dict(time_batch(ctx, stream, 1, 20000, wl.TRUTH_A, False, 200),
                             note="one chain x 20 000 points (test_likelihoods.c): the light curve is shared by several CTAs")
    out["C3_share"] = dict(time_batch(ctx, stream, 2048, 20000, wl.TRUTH_A, False, 20),
                           note="64 temperatures x 32 ensembles: the likelihood batch of one PT step")
    out["C4"] = dict(time_batch(ctx, stream, 8192, 50000, wl.TRUTH_A, True, 5), note="8192 chains x 50 000 points + Gaia G term")
    out["C5_share"] = dict(time_batch(ctx, stream, 2048, 200000, wl.TRUTH_B, False, 5),
                           note="2048 of 16 384 chains x 200 000 points, truth B, e <= 0.95")
    ctx.set_mags([1000, 1, 1, 1, 1], [1e15] * 4, 1, 0)
    return out


def shard_probe(ctx, rank, world, dist):
    """Sharded logL == full logL, bit for bit, on a 512-chain probe (every rank evaluates the whole probe and its
    own shard; the shards are all-gathered over NCCL)."""
    import torch
    N, n = 20000, 512
    t, flux, err = wl.make_dataset(N, wl.TRUTH_A, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    P = wl.draw_chains(n, wl.TRUTH_A, ctx.roche_overflow, seed=77)
    full = ctx.loglikelihood(P)
    per = -(-n // world)
    lo, hi = min(rank * per, n), min((rank + 1) * per, n)
    mine = np.full(per, np.nan)
    mine[: hi - lo] = ctx.loglikelihood(P[lo:hi]) if hi > lo else []
    send = torch.from_numpy(mine).to("cuda")
    recv = torch.empty(world * per, dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(recv, send)
    gathered = recv.cpu().numpy()[:n]
    ok = bool(np.array_equal(gathered, full, equal_nan=True))
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag.item())


def _timed_steps(sp, ctx, stream, dist, n_steps):
    """n_steps iterations through hb_pt_step (captured graphs), CUDA-event time, max over ranks; the evaluated-walker
    count of the same steps from the likelihood kernel's own counter."""
    import torch
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ctx.evaluated_chains(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launch_count
    e0.record(stream)
    sp.step(n_steps)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    evaluated = ctx.evaluated_chains(reset=True)
    if dist is not None:
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
        ev = torch.tensor([evaluated], dtype=torch.int64, device="cuda")
        dist.all_reduce(ev, op=dist.ReduceOp.SUM)
        evaluated = int(ev.item())
    return ms, evaluated, int(ctx.launch_count - l0)


def _pt_summary(sp, ms, evaluated, n_steps, walkers_total, n_points, launches):
    cnt = sp.sampler.counters()
    it = max(1, int(cnt["iterations"].sum()))
    return {
        "steps_per_sec": n_steps / (ms * 1e-3), "ms_per_step": ms / n_steps, "steps": n_steps,
        "walkers": walkers_total, "n_points": n_points, "likelihood_evals_per_walker_per_step": 1,
        # walkers whose model the likelihood kernel really evaluated (Roche-overflowing and e >= 1 proposals
        # return early, quirk Q13, and are NOT counted), from the kernel's own device counter
        "evaluated_walkers_per_step": evaluated / n_steps,
        "skipped_fraction": 1.0 - evaluated / (n_steps * walkers_total),
        "model_points_per_sec": evaluated * n_points / (ms * 1e-3),
        "acceptance": float(cnt["accepted"].sum() / max(1, cnt["proposed"].sum())),
        "de_share_cold_chain": float(cnt["de_trials_slot0"].sum() / it),
        "swap_acceptance": float(cnt["swaps_accepted"].sum() / max(1, cnt["swaps_proposed"].sum())),
        "gpu_launches": launches,
    }


def pt_leg(args, ctx, cfg, rank, world, dist, stream):
    """Second half of BASELINE.json's metric: full PT-MCMC iterations per second.  Every iteration = propose + ONE
    likelihood per walker + accept + n_temps swap proposals per ladder, replayed from captured CUDA graphs
    (hb_pt_step); timed in steady state: after `npast` = 50 iterations, so the differential-evolution proposals of
    mcmc_wrapper2.c:1091-1140 are in, from a start near the truth (a converged chain), Roche-rejected proposals
    counted as NOT evaluated.
      weak     C3's per-GPU share on every GPU: 64 temperatures x 32 ensembles x 20 000 points per GPU
      (N > 1) strong   C3 as named: 64 x 256 walkers in all, whole ladders per GPU, no per-step exchange
      (N > 1) rungs    ONE ladder of 64 rungs x 200 000 points: every GPU holds the ladder, evaluates its shard of the
                       rungs, and the logL vector is all-gathered by NCCL inside the captured step"""
    import torch
    from hb_mcmc_b200.pt import ShardedPT
    n_temps, ens_per_gpu, npast = 64, 32, 50
    steps = max(args.pt_steps, 8)
    truth = cfg["truth_vec"]
    logp = float(truth[2])

    def start_near_truth(sp):
        W = sp.sampler.n_walkers
        rng = np.random.default_rng(1234 + sp.first)
        x = truth + 1e-4 * rng.standard_normal((W, 21)) * np.abs(truth + 0.1)
        x[:, 2] = logp
        sp.sampler.set_state(x)

    def run(n_ens_total, label):
        sp = ShardedPT(ctx, n_temps, n_ens_total, logp, seed=11, npast=npast)
        start_near_truth(sp)
        sp.step(npast + 14)  # history rings filled: DE proposals active
        ms, evaluated, launches = _timed_steps(sp, ctx, stream, dist, steps)
        info = _pt_summary(sp, ms, evaluated, steps, n_temps * n_ens_total, cfg["n_points"], launches)
        info.update({"n_temps": n_temps, "ensembles_total": n_ens_total, "split": sp.mode if world > 1 else "none", "label": label})
        cold = sp.gather_cold_logL()
        info["cold_logL_finite"] = bool(np.isfinite(cold).all())
        sp.close()
        return info

    info = run(ens_per_gpu * world, "C3 share per GPU (weak scaling): 64 temperatures x 32 ensembles x 20 000 points per GPU")
    info["ensembles_per_gpu"] = ens_per_gpu
    info["walkers_per_gpu"] = n_temps * ens_per_gpu
    info["exchange"] = "none per step (whole ladders per GPU; random streams keyed on global ids)"
    if world > 1:
        info["strong_C3_as_named"] = run(256, "C3 as named: 64 temperatures x 256 ensembles x 20 000 points in all")
    return info


def rung_split_leg(ctx, rank, world, dist, stream, steps):
    """ONE ladder of 64 rungs on a 200 000-point light curve (truth B): at N > 1 every GPU holds the ladder, evaluates
    its shard of the rungs and the per-step logL vector is all-gathered by NCCL from inside libhb_b200's captured step."""
    from hb_mcmc_b200.pt import ShardedPT
    N, n_temps, npast = 200000, 64, 50
    t, flux, err = wl.make_dataset(N, wl.TRUTH_B, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    ctx.set_mags([1000, 1, 1, 1, 1], [1e15] * 4, 1, 0)
    sp = ShardedPT(ctx, n_temps, 1, float(wl.TRUTH_B[2]), seed=13, npast=npast)
    rng = np.random.default_rng(99)
    x = wl.TRUTH_B + 1e-4 * rng.standard_normal((n_temps, 21)) * np.abs(wl.TRUTH_B + 0.1)
    x[:, 2] = wl.TRUTH_B[2]
    sp.sampler.set_state(x)
    sp.step(npast + 14)
    ms, evaluated, launches = _timed_steps(sp, ctx, stream, dist, steps)
    info = _pt_summary(sp, ms, evaluated, steps, n_temps, N, launches)
    info.update({"n_temps": n_temps, "ensembles_total": 1, "split": sp.mode if world > 1 else "none",
                 "exchange": ("ncclAllGather of the logL vector (%d doubles per rank) inside the captured step" % sp.sampler.eval_shard()[2])
                 if world > 1 else "single rank"})
    sp.close()
    return info


def pt_reference_size_leg(ctx):
    """The reference's own use case: ONE ladder of 50 rungs on a real folded light curve (TIC 102289966,
    375 points, tests/golden/).  GPU: hb_pt_step.  CPU: the unmodified reference driver binary
    (oracle/_ref/hb_mcmc_ref, built from mcmc_wrapper2.c with only its /scratch prefix and thread count
    patched) for 2000 iterations on the host cores -- it evaluates 2 likelihoods per rung per step."""
    import re
    import shutil
    import numpy as np
    from hb_mcmc_b200.pt import PTSampler
    gold = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(gold, "lc_102289966_new.txt")) as f:
        n = int(f.readline())
        d = np.loadtxt(f)
    logp = 0.7960497
    ctx.set_data(d[:, 0], d[:, 1], d[:, 2])
    ctx.set_mags([1000, 1, 1, 1, 1], [1e15] * 4, 1, 0)
    s = PTSampler(ctx, 50, 1, logp, seed=3)
    s.init_random()
    s.step(200)
    ctx.sync()
    t0 = time.perf_counter()
    iters = 3000
    s.step(iters)
    ctx.sync()
    gpu_rate = iters / (time.perf_counter() - t0)
    s.close()
    out = {"n_temps": 50, "n_ens": 1, "n_points": n, "gpu_steps_per_sec": gpu_rate, "gpu_iterations": iters}
    exe = os.path.join(ROOT, "oracle", "_ref", "hb_mcmc_ref")
    scr = os.path.join(ROOT, "oracle", "_ref", "scratch")
    if os.path.exists(exe) and os.path.isdir(scr):
        shutil.copy(os.path.join(gold, "lc_102289966_new.txt"),
                    os.path.join(scr, "data", "lightcurves", "folded_lightcurves", "102289966_new.txt"))
        ref_iters = 1500
        t0 = time.perf_counter()
        r = subprocess.run([exe, str(ref_iters), "102289966", repr(logp), "9"], capture_output=True, text=True, cwd=scr)
        dt = time.perf_counter() - t0
        if r.returncode == 0 and re.search(r"Begining main mcmc loop", r.stdout):
            out.update({"cpu_steps_per_sec": ref_iters / dt, "cpu_iterations": ref_iters, "cpu_kind": "reference",
                        "cpu_threads": 8, "cpu_note": "unmodified mcmc_wrapper2.c driver, 2 likelihood evaluations per rung per step"})
        # the same unmodified driver with likelihood3.c replaced by libhb_likelihood3.so on the link line
        # (25 OpenMP threads as in the reference; their concurrent calls are combined into device batches)
        shim_exe = os.path.join(ROOT, "oracle", "_ref", "hb_mcmc_ref_shim")
        if os.path.exists(shim_exe):
            # HB_SHIM_STATS=1 makes the shim print, at exit, the wall-clock span from the end of its first device
            # batch to the end of its last: the run without the process start-up (CUDA context creation takes 1-4 s
            # on a fresh box, as long as the 1500 steps themselves); the whole-process figure is reported next to it
            t0 = time.perf_counter()
            try:
                r = subprocess.run([shim_exe, str(ref_iters), "102289966", repr(logp), "9"], capture_output=True, text=True,
                                   cwd=scr, env=dict(os.environ, HB_SHIM_STATS="1"), timeout=180)
            except subprocess.TimeoutExpired:
                return out
            dt = time.perf_counter() - t0
            m = re.search(r"span ([0-9.]+) s", r.stderr + r.stdout)
            if r.returncode == 0 and re.search(r"Begining main mcmc loop", r.stdout) and m and float(m.group(1)) > 0:
                out.update({"shim_steps_per_sec": ref_iters / float(m.group(1)),
                            "shim_steps_per_sec_whole_process": ref_iters / dt,
                            "shim_note": "unmodified mcmc_wrapper2.c linked against libhb_likelihood3.so (link-level drop-in); "
                                         "rate over the span from the shim's first to its last device batch (start-up excluded)"})
    return out


def cpu_baseline_leg(cfg, t, flux, err, P, logL_gpu):
    """Reference CPU path on the box's host cores over a bounded sample of the SAME inputs
    (about 20 core-seconds); also re-checks parity on that sample."""
    backend, kind = cpu_backend()
    cores = os.cpu_count() or 1
    mags = (None, None)
    if cfg["gaia"]:
        mags = gaia_setup(cfg["truth_vec"], backend.calc_mags(cfg["truth_vec"], 100.0)[0])
    probe_dt, _ = run_cpu_sample(backend, t, flux, err, P[:cores], mags, cores)
    rounds = max(1, min(int(20.0 / max(probe_dt * cores, 1e-3)), len(P) // cores))
    n_s = min(len(P), rounds * cores)
    dt, out = run_cpu_sample(backend, t, flux, err, P[:n_s], mags, cores)
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = np.nanmax(np.abs(out - logL_gpu[:n_s]) / np.abs(out))
    return {"value": n_s * len(t) / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"first {n_s} of {len(P)} chains x {len(t)} points, {cores} threads",
            "max_rel_err_gpu_vs_cpu_on_sample": float(rel)}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = workload_spec(args)
    if args.impl == "reference":
        reference_arm(args, cfg, rank)
        return
    if world == 1 and args.gpus > 1:
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    gpu_arm(args, cfg, rank, local_rank, world)


if __name__ == "__main__":
    main()
