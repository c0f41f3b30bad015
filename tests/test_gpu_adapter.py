"""`Kokkos::B200` attached to the UNMODIFIED reference (the drop-in boundary):

* tests/ref_unit/_build_adapter/ref_unit_kokkos_b200 -- the reference's OWN unit-test sources (tests/ref_unit/adapter.list)
  compiled against the reference's real <Kokkos_Core.hpp> with TEST_EXECSPACE=Kokkos::B200
  (kokkos_b200/adapter/Kokkos_B200_Space.hpp); Kokkos::Cuda and Kokkos::OpenMP live in the same binary;
* benchlib/libkokkos_arms.so -- the headline workloads as Kokkos user lambdas, B200 vs Kokkos::Cuda vs CUB vs the oracle.
Both are built in the build container (they need /root/reference) and travel to the GPU box."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "ref_unit", "_build_adapter", "ref_unit_kokkos_b200")


def _env():
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "kokkos_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    env["OMP_NUM_THREADS"] = "4"
    env["OMP_PROC_BIND"] = "false"
    return env


def test_reference_unit_tests_pass_on_kokkos_b200_adapter():
    if not os.path.exists(BIN):
        pytest.skip("tests/ref_unit/_build_adapter/ref_unit_kokkos_b200 not built (needs /root/reference at build time)")
    p = subprocess.run([BIN, "--gtest_color=no"], capture_output=True, text=True, timeout=900, env=_env())
    out = p.stdout + p.stderr
    tail = "\n".join(out.splitlines()[-60:])
    assert p.returncode == 0, tail
    m = re.search(r"\[  PASSED  \] (\d+) tests", out)
    assert m and int(m.group(1)) >= 150, tail
    assert "FAILED" not in out, tail


def test_kokkos_user_lambdas_on_b200_match_oracle_and_kokkos_cuda():
    if not os.path.exists(os.path.join(ROOT, "benchlib", "libkokkos_arms.so")):
        pytest.skip("benchlib/libkokkos_arms.so not built (needs /root/reference at build time)")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "arms_worker.py")], capture_output=True, text=True, timeout=600, env=_env())
    assert p.returncode == 0 and "arms ok" in p.stdout, (p.stdout + p.stderr)[-3000:]
