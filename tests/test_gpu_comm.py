"""GPU tests of the one-box communicator and the fused block-cyclic scan (kokkos_b200/csrc/comm.cu,
kb200/impl/ScanChunked.hpp), through the C ABI.  world 1 always runs; world 2 / 4 / 8 run when the box has the GPUs
(one process per GPU, started here; the bootstrap is the library's own shared-memory rendezvous, no torch.distributed)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import workloads as W  # noqa: E402

pytestmark = pytest.mark.gpu


def _run_world(world, extra=(), algo=None):
    import kokkos_b200 as kb
    if kb.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    uid = kb.comm_unique_id()
    env = dict(os.environ)
    if algo is not None:
        env["KB200_COMM_ALGO"] = str(algo)
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "comm_worker.py"), str(r), str(world), uid, *extra],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env) for r in range(world)]
    outs = []
    for r, p in enumerate(procs):
        try:
            out, _ = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for r, p in enumerate(procs):
        assert p.returncode == 0 and f"ok rank {r}" in outs[r], f"rank {r} failed:\n{outs[r][-3000:]}"


def test_comm_world1():
    _run_world(1)


def test_comm_world2():
    _run_world(2)


def test_comm_world2_lockstep_kernel():
    _run_world(2, algo=1)


def test_comm_world4():
    _run_world(4)


def test_comm_world8():
    _run_world(8)


@pytest.mark.parametrize("n", [0, 1, 5, 6911, 6912, 6913, 148 * 6912, 148 * 6912 + 1, 3 * 148 * 6912 + 77, (1 << 24) + 3])
@pytest.mark.parametrize("inclusive", [False, True])
def test_chunked_scan_single_gpu_bit_exact(space, port, n, inclusive):
    """The chunk-synchronous kernel at world 1 (tune key scan.chunked) against the oracle, ragged sizes, with a seed."""
    import kokkos_b200 as kb
    x = W.c3_wrap(n) if n else np.zeros(0, dtype=np.int64)
    vx = space.view_from_host(x) if n else space.view(2, np.int64)
    vy = space.view(max(n, 2), np.int64)
    vx.n = vy.n = n
    kb.tune_set("scan.chunked", 1)
    try:
        total = space.parallel_scan(vx, vy, inclusive=inclusive, seed=-11)
    finally:
        kb.tune_set("scan.chunked", 0)
    py, pt = port.scan(x, inclusive, -11, 4)
    assert total == pt
    if n:
        assert np.array_equal(vy.to_host(), py)


def test_chunked_scan_f64_integer_valued(space, port):
    import kokkos_b200 as kb
    n = 2 * 148 * 6912 + 1001
    x = W.c1_exact(n)
    vx, vy = space.view_from_host(x), space.view(n, np.float64)
    kb.tune_set("scan.chunked", 1)
    try:
        total = space.parallel_scan(vx, vy, inclusive=True)
    finally:
        kb.tune_set("scan.chunked", 0)
    assert np.array_equal(vy.to_host(), np.cumsum(x)) and total == float(x.sum())
