"""In-tree build of the native libraries (nvcc cross-compiles sm_100a without a GPU).

    python -m hb_mcmc_b200.build            # build what is stale
    python -m hb_mcmc_b200.build --force

Outputs (git-ignored, shipped to the GPU box by gpurun):
    hb_mcmc_b200/csrc/libhb_b200.so          CUDA kernels + the C ABI of include/hb_b200.h
    hb_mcmc_b200/csrc/libhb_likelihood3.so   the reference's likelihood3.h symbols on top of it
    host/hb_mcmc                              the C driver (mcmc_wrapper2-compatible CLI)
    host/hb_gaia_mcmc                         the C driver of the Gaia-colour sampler (GAIA_mcmc-compatible CLI)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(CSRC, "libhb_b200.so")
LIB_DEBUG = os.path.join(CSRC, "libhb_b200_dbg.so")  # -DHB_DEBUG_BOUNDS: every indexed access asserted (tests only)
SHIM = os.path.join(CSRC, "libhb_likelihood3.so")
HOST_DIR = os.path.join(ROOT, "host")
DRIVER = os.path.join(HOST_DIR, "hb_mcmc")
GAIA_DRIVER = os.path.join(HOST_DIR, "hb_gaia_mcmc")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libhb_b200.so cannot be built")


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _glob(d: str, exts: tuple[str, ...]) -> list[str]:
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts))


def build_lib(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, f) for f in ("hb_kernels.cu", "hb_capi.cu", "hb_pt.cu", "hb_gaia_pt.cu", "hb_comm.cu")]
    srcs = [s for s in srcs if os.path.exists(s)]
    deps = _glob(CSRC, (".cu", ".cuh", ".h")) + [os.path.join(ROOT, "include", "hb_b200.h")]
    if force or _stale(LIB, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + srcs + ["-ldl"]
        subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB


def build_debug_lib(force: bool = False) -> str:
    """The bounds-asserting build (hb_select.cuh: HB_CHK): same sources, -DHB_DEBUG_BOUNDS."""
    srcs = [os.path.join(CSRC, f) for f in ("hb_kernels.cu", "hb_capi.cu", "hb_pt.cu", "hb_gaia_pt.cu", "hb_comm.cu")]
    deps = _glob(CSRC, (".cu", ".cuh", ".h")) + [os.path.join(ROOT, "include", "hb_b200.h")]
    if force or _stale(LIB_DEBUG, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + ["-DHB_DEBUG_BOUNDS", "-o", LIB_DEBUG] + srcs + ["-ldl"]
        subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB_DEBUG


def build_shim(force: bool = False) -> str | None:
    src = os.path.join(CSRC, "likelihood3_shim.c")
    if not os.path.exists(src):
        return None
    if force or _stale(SHIM, [src, LIB, os.path.join(ROOT, "include", "hb_b200.h")]):
        cmd = ["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-o", SHIM, src,
               "-L", CSRC, "-lhb_b200", "-Wl,-rpath,$ORIGIN", "-lm", "-lpthread"]
        subprocess.run(cmd, check=True, cwd=CSRC)
    return SHIM


def build_driver(force: bool = False) -> str | None:
    built = None
    for name, target in (("hb_mcmc.c", DRIVER), ("hb_gaia_mcmc.c", GAIA_DRIVER)):
        src = os.path.join(HOST_DIR, name)
        if not os.path.exists(src):
            continue
        if force or _stale(target, [src, LIB, os.path.join(ROOT, "include", "hb_b200.h")]):
            cmd = ["gcc", "-O2", "-std=gnu99", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", target, src, "-L", CSRC,
                   "-lhb_b200", "-Wl,-rpath," + CSRC, "-lm", "-lpthread"]
            subprocess.run(cmd, check=True, cwd=HOST_DIR)
        built = built or target
    return built


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_lib(force, verbose)
    build_debug_lib(force)
    build_shim(force)
    build_driver(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", LIB)
