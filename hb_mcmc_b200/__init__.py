"""hb_mcmc_b200 -- B200-native (sm_100a) hot path of sidruns30/HB_MCMC.

The product is the CUDA library ``csrc/libhb_b200.so`` behind the C ABI of
``include/hb_b200.h``; this package is its Python host-side mirror (ctypes, numpy) plus
``torch.distributed`` plumbing for the multi-GPU parallel-tempering driver.
"""
from .lib import ABI_SYMBOLS, BIG_NUM, NPARS, Context, HBError, load_library  # noqa: F401

__all__ = ["Context", "HBError", "load_library", "NPARS", "BIG_NUM", "ABI_SYMBOLS"]
