"""Host-side checks of the reference-unit-test harness (tests/ref_unit): the source lists and the restarting gtest driver.
No GPU and no compute: the binaries themselves run in tests/test_gpu_ref_unit.py and tests/test_gpu_adapter.py."""
import os
import stat
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_UNIT = os.path.join(ROOT, "tests", "ref_unit")
REF_TESTS = "/root/reference/core/unit_test"


@pytest.mark.parametrize("name", ["tests.list", "adapter.list"])
def test_source_lists_are_well_formed(name):
    path = os.path.join(REF_UNIT, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not present")
    entries = [ln.strip() for ln in open(path) if ln.strip() and not ln.startswith("#")]
    assert entries and len(entries) == len(set(entries)), "duplicate entries"
    assert all(e.endswith((".hpp", ".cpp")) and " " not in e for e in entries)
    if os.path.isdir(REF_TESTS):  # the build container; the GPU box has no reference tree
        missing = [e for e in entries if not os.path.exists(os.path.join(REF_TESTS, e))]
        assert not missing, missing


FAKE = textwrap.dedent('''\
    #!%s
    import os, sys
    tests = ["b200.a", "b200.b", "b200.crash", "b200.c", "b200.fail", "b200.skip", "b200_graph.d"]
    flt = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--gtest_filter=")]
    sel = flt[0].split(":") if flt else tests
    if "--gtest_list_tests" in sys.argv:
        print("runtime chatter before the list")
        print("b200.")
        for t in tests[:-1]:
            print("  " + t.split(".")[1])
        print("b200_graph.")
        print("  d  # TypeParam = int")
        sys.exit(0)
    for t in tests:
        if t not in sel:
            continue
        print(f"[ RUN      ] {t}")
        if t.endswith("crash"):
            sys.stdout.flush()
            os.abort()
        if t.endswith("fail"):
            print("x.cpp:1: Failure")
            print(f"[  FAILED  ] {t} (1 ms)")
        elif t.endswith("skip"):
            print(f"[  SKIPPED ] {t} (0 ms)")
        else:
            print(f"[       OK ] {t} (0 ms)")
    print("[  PASSED  ] n tests.")
    ''') % sys.executable


def test_restarting_driver_survives_a_crashing_test(tmp_path):
    fake = tmp_path / "fake_gtest.py"
    fake.write_text(FAKE)
    fake.chmod(fake.stat().st_mode | stat.S_IXUSR)
    log = tmp_path / "log.txt"
    p = subprocess.run([sys.executable, os.path.join(REF_UNIT, "run_resilient.py"), str(fake), "--limit", "30", "--log", str(log)],
                       capture_output=True, text=True, timeout=120)
    lines = p.stdout.splitlines()
    assert p.returncode == 1
    assert any(ln.split() == ["CRASHED", "b200.crash"] for ln in lines), p.stdout
    assert any(ln.split() == ["FAILED", "b200.fail"] for ln in lines), p.stdout
    assert any(ln.split() == ["SKIPPED", "b200.skip"] for ln in lines), p.stdout
    assert lines[-1] == "summary: 7 tests, 1 CRASHED, 1 FAILED, 4 OK, 1 SKIPPED", lines[-1]
    assert "[ RUN      ] b200_graph.d" in log.read_text()   # the tests behind the crash did run, in a second process
