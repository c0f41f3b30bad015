"""The reference's OWN unit tests on the B200 execution space.

tests/ref_unit/Makefile compiles the reference test sources named in tests/ref_unit/tests.list (core/unit_test/incremental/*,
core/unit_test/Test*.hpp, core/unit_test/default/TestDefaultDeviceType_*.cpp) UNMODIFIED, from /root/reference and with its
vendored gtest, against the kb200 layer exposed as `Kokkos::`, and links them into
tests/ref_unit/_build/ref_unit_b200 (built in the build container; the binary travels to the GPU box).  This test runs it
and requires every gtest case to pass."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "ref_unit", "_build", "ref_unit_b200")


def test_reference_incremental_unit_tests_pass_on_b200():
    if not os.path.exists(BIN):
        pytest.skip("tests/ref_unit/_build/ref_unit_b200 not built (needs /root/reference at build time)")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(ROOT, "kokkos_b200") + ":" + env.get("LD_LIBRARY_PATH", "")
    p = subprocess.run([BIN, "--gtest_color=no"], capture_output=True, text=True, timeout=600, env=env)
    out = p.stdout + p.stderr
    tail = "\n".join(out.splitlines()[-40:])
    assert p.returncode == 0, tail
    m = re.search(r"\[  PASSED  \] (\d+) tests", out)
    assert m and int(m.group(1)) >= 150, tail
    assert "FAILED" not in out, tail
