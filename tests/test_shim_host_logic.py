"""Host logic of the link-level drop-in (likelihood3_shim.c) without a GPU: the shim is compiled against a TEST-ONLY
stand-in for libhb_b200.so (tests/shim_stub/) and driven by an OpenMP team the way the reference's rung loop drives
it (mcmc_wrapper2.c:383,488-489): concurrent callers are combined, every caller gets its own value, repeated calls are
answered from the memo with the same bits.  The product library is not involved."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "tests", "shim_stub")
SHIM_SRC = os.path.join(ROOT, "hb_mcmc_b200", "csrc", "likelihood3_shim.c")


@pytest.fixture(scope="module")
def stress(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("shim_stub"))
    inc = os.path.join(ROOT, "include")
    run = lambda *cmd: subprocess.run(cmd, check=True, cwd=d, capture_output=True, text=True)
    run("gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-I", inc, "-o", "libhb_b200.so", os.path.join(STUB, "stub_hb_b200.c"))
    run("gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-I", inc, "-o", "libhb_likelihood3.so", SHIM_SRC, "-L", d, "-lhb_b200",
        "-lpthread", "-lm")
    run("gcc", "-O2", "-std=gnu99", "-fopenmp", "-o", "stress", os.path.join(STUB, "stress.c"), "-L", d, "-lhb_likelihood3",
        "-lhb_b200", "-lm")

    def go(steps, threads, **env):
        e = dict(os.environ, LD_LIBRARY_PATH=d, HB_SHIM_STATS="1", **env)
        return subprocess.run([os.path.join(d, "stress"), str(steps), str(threads)], capture_output=True, text=True, env=e,
                              timeout=300)
    return go


@pytest.mark.parametrize("threads", [1, 4, 25])
def test_team_of_callers_gets_its_own_values(stress, threads):
    r = stress(400, threads, HB_SHIM_MEMO="0")
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout + r.stderr
    m = re.search(r"(\d+) loglikelihood calls in (\d+) batches", r.stderr)
    calls, batches = int(m.group(1)), int(m.group(2))
    assert calls == 50 + 400 * 100  # every call reached the (stand-in) device
    if threads == 25:
        assert batches < calls / 4  # the team's concurrent calls were combined


def test_repeated_calls_come_from_the_memo(stress):
    r = stress(400, 25)
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout + r.stderr
    m = re.search(r"(\d+) loglikelihood calls in \d+ batches.*; (\d+) more calls answered from the memo", r.stderr)
    evaluated, remembered = int(m.group(1)), int(m.group(2))
    assert evaluated + remembered == 50 + 400 * 100
    # the current state of every rung, every step (same bits: checked by the harness); an eviction only costs a re-evaluation
    assert remembered >= 0.98 * 400 * 50
