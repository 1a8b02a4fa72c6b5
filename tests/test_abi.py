"""The C-ABI library loads and exports every symbol include/hb_b200.h declares (no compute)."""
import os
import re

import pytest

import hb_mcmc_b200 as hb
from hb_mcmc_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    L = hb.load_library()
    decl = declared_symbols("hb_b200.h")
    assert len(decl) >= 19
    for s in decl:
        assert hasattr(L, s), f"{s} declared in include/hb_b200.h but not exported"
    assert sorted(hb.ABI_SYMBOLS) == sorted(s for s in decl if s != "hb_ctx")


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", build.LIB], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(hb.HBError):
        hb.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hb_mcmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "hb_oracle" not in text, f
