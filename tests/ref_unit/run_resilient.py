#!/usr/bin/env python3
"""Run a gtest binary to the end of its test list even when single tests abort or hang the process.

The reference's unit tests are written to abort on some errors (Kokkos::abort, death tests, a sticky CUDA error), which ends the
whole gtest process and hides every test after it.  This driver restarts the binary on the remaining tests, marks the test that
was running as CRASHED (or HUNG, when the per-process time limit ended it) and prints one line per test plus a summary:

    python tests/ref_unit/run_resilient.py <binary> [--limit SECONDS_PER_PROCESS] [--filter GTEST_FILTER] [--log FILE]

Exit status 0 when nothing failed, crashed or hung.  Test infrastructure only (used when widening tests/ref_unit/adapter.list)."""
import argparse
import re
import subprocess
import sys


def list_tests(binary, flt):
    cmd = [binary, "--gtest_list_tests"] + ([f"--gtest_filter={flt}"] if flt else [])
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=120).stdout
    tests, suite = [], None
    for line in out.splitlines():
        m = re.match(r"^(\S+\.)\s*(#.*)?$", line)
        if m:
            suite = m.group(1)
            continue
        m = re.match(r"^  (\S+)", line)
        if m and suite:
            tests.append(suite + m.group(1))
        elif line.strip():
            suite = None  # runtime chatter between the lists
    return tests


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("binary")
    ap.add_argument("--limit", type=float, default=120.0)
    ap.add_argument("--filter", default="")
    ap.add_argument("--log", default="")
    a = ap.parse_args()
    remaining = list_tests(a.binary, a.filter)
    total = len(remaining)
    results, log = {}, []
    while remaining:
        try:
            p = subprocess.run([a.binary, "--gtest_color=no", "--gtest_filter=" + ":".join(remaining)], capture_output=True, text=True,
                               timeout=a.limit)
            out, hung = p.stdout + p.stderr, False
        except subprocess.TimeoutExpired as e:
            out = (e.stdout or b"").decode(errors="replace") + (e.stderr or b"").decode(errors="replace")
            hung = True
        log.append(out)
        running = None
        for line in out.splitlines():
            m = re.match(r"\[ RUN      \] (\S+)", line)
            if m:
                running = m.group(1)
                continue
            m = re.match(r"\[\s+(OK|FAILED|SKIPPED)\s+\] (\S+?),? ", line + " ")
            if m and m.group(2) == running:
                results[running] = m.group(1)
                running = None
        if running is not None:
            results[running] = "HUNG" if hung else "CRASHED"
        done = set(results)
        before = len(remaining)
        remaining = [t for t in remaining if t not in done]
        if len(remaining) == before:  # no progress: the binary fails before running anything
            for t in remaining:
                results[t] = "NOT_RUN"
            break
    if a.log:
        with open(a.log, "w") as f:
            f.write("\n".join(log))
    counts = {}
    for t, r in results.items():
        counts[r] = counts.get(r, 0) + 1
        if r not in ("OK",):
            print(f"{r:8s} {t}")
    print(f"summary: {total} tests, " + ", ".join(f"{v} {k}" for k, v in sorted(counts.items())))
    return 0 if all(r in ("OK", "SKIPPED") for r in results.values()) else 1


if __name__ == "__main__":
    sys.exit(main())
