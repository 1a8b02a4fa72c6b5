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


def test_proposal_table_reads_the_env_flags_before_any_likelihood(tmp_path):
    """The reference driver calls initialize_proposals (mcmc_wrapper2.c:198) before its first loglikelihood (:342):
    the sigma table must already follow HB_USE_GMAG / HB_USE_COLOR_INFO (likelihood3.c:1158-1179) then, without a
    device context having been created (the stand-in library would abort the process if it were asked for one)."""
    d, inc = str(tmp_path), os.path.join(ROOT, "include")
    run = lambda *cmd: subprocess.run(cmd, check=True, cwd=d, capture_output=True, text=True)
    # a libhb_b200 whose hb_create fails: any attempt to open the device ends the process
    with open(os.path.join(d, "nodev.c"), "w") as f:
        f.write('#include "hb_b200.h"\nint hb_create(hb_ctx** o, int dev) { (void)o; (void)dev; return HB_ERR_CUDA; }\n'
                'const char* hb_global_error(void) { return "no device in this test"; }\n')
    run("gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-I", inc, "-o", "libnodev.so", "nodev.c")
    run("gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-I", inc, "-o", "libhb_b200.so", os.path.join(STUB, "stub_hb_b200.c"))
    run("gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-I", inc, "-o", "libhb_likelihood3.so", SHIM_SRC, "-L", d, "-lhb_b200",
        "-lpthread", "-lm")
    with open(os.path.join(d, "sig.c"), "w") as f:
        f.write('#include <stdio.h>\nvoid initialize_proposals(double*, double***);\n'
                'int main(void) { double s[21]; initialize_proposals(s, 0); printf("%g %g %g\\n", s[0], s[4], s[12]); return 0; }\n')
    run("gcc", "-O2", "-o", "sig", "sig.c", "-L", d, "-lhb_likelihood3", "-lhb_b200", "-lm")
    go = lambda **env: subprocess.run([os.path.join(d, "sig")], capture_output=True, text=True, timeout=60,
                                      env=dict(os.environ, LD_LIBRARY_PATH=d, LD_PRELOAD=os.path.join(d, "libnodev.so"), **env))
    r = go()
    assert r.returncode == 0 and r.stdout.split() == ["0.1", "0.01", "0.1"], r.stdout + r.stderr  # defaults: colours off
    r = go(HB_USE_GMAG="1", HB_USE_COLOR_INFO="1")
    assert r.returncode == 0 and r.stdout.split() == ["0.01", "0.001", "0.01"], r.stdout + r.stderr
    r = go(HB_USE_GMAG="0", HB_USE_COLOR_INFO="1")
    assert r.returncode == 0 and r.stdout.split() == ["0.1", "0.01", "0.1"], r.stdout + r.stderr
