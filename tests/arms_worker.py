"""Parity of the Kokkos-user-code arms (benchlib/libkokkos_arms.so) -- run as a script by tests/test_gpu_adapter.py.

The same KOKKOS_LAMBDA functors, compiled against the UNMODIFIED reference headers, run on Kokkos::B200 (the adapter,
kokkos_b200/adapter) and on the reference's own Kokkos::Cuda; both are checked against the CPU oracle port on the same seeded
inputs (bit-exact: integer-valued doubles / int64 / locations) and against each other.  Prints "arms ok" on success."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import workloads as W  # noqa: E402
from benchlib import arms as A  # noqa: E402
from oracle.bindings import Port  # noqa: E402


def main():
    port = Port()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    arms = A.Arms(0, stream.cuda_stream)
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731

    for n in (1, 1000, 100003, (1 << 22) + 5):
        x = W.c1_exact(n); xi = W.c3_wrap(n)
        dx, dxi = t(x), t(xi)
        exp_sum = port.reduce("sum", x, 4)
        exp_y, exp_tot = port.scan(xi, False, 0, 4)
        for arm in (A.B200, A.CUDA, A.CUB):
            r = torch.zeros(1, dtype=torch.float64, device=dev)
            arms.reduce_sum(arm, dx.data_ptr(), n, r.data_ptr())
            assert r.item() == exp_sum, (A.ARM_NAMES[arm], n, r.item(), exp_sum)
            y = torch.full((n,), -7, dtype=torch.int64, device=dev)
            tot = torch.zeros(1, dtype=torch.int64, device=dev)
            arms.scan_excl(arm, dxi.data_ptr(), y.data_ptr(), n, tot.data_ptr())
            assert np.array_equal(y.cpu().numpy(), exp_y), (A.ARM_NAMES[arm], n, "scan")
            if arm != A.CUB:
                assert tot.item() == exp_tot, (A.ARM_NAMES[arm], n, "scan total")
            y.fill_(-7)
            arms.std_exclusive_scan(arm, dxi.data_ptr(), y.data_ptr(), n)   # Kokkos::Experimental::exclusive_scan (typed route on B200)
            assert np.array_equal(y.cpu().numpy(), exp_y), (A.ARM_NAMES[arm], n, "std exclusive_scan")
        # stream copy / triad
        b, c = t(W.c1_general(n)), t(W.c1_uniform(n))
        for arm in (A.B200, A.CUDA):
            a = torch.zeros(n, dtype=torch.float64, device=dev)
            arms.stream_copy(arm, b.data_ptr(), a.data_ptr(), n)
            assert torch.equal(a, b)
            arms.stream_triad(arm, a.data_ptr(), b.data_ptr(), c.data_ptr(), 3.0, n)
            assert np.array_equal(a.cpu().numpy(), b.cpu().numpy() + 3.0 * c.cpu().numpy()), (A.ARM_NAMES[arm], n, "triad")

    # MDRange 7-point stencil with MinMaxLoc (C4), ragged extents
    for (n0, n1, n2) in ((34, 20, 18), (67, 33, 41), (130, 64, 37)):
        u, _, _ = W.c4_field(n0, n1, n2)
        q, _ = port.stencil7(u, n0, n1, n2, 0.5, 0.125)
        du = t(u)
        for arm in (A.B200, A.CUDA):
            res = torch.zeros(4, dtype=torch.float64, device=dev)
            arms.stencil7_minmaxloc(arm, du.data_ptr(), n0, n1, n2, 0.5, 0.125, res.data_ptr())
            h = res.cpu().numpy()
            got = (h[0], h[1], int(h[2:].view(np.int64)[0]), int(h[2:].view(np.int64)[1]))
            assert got == (q.min_val, q.max_val, q.min_loc, q.max_loc), (A.ARM_NAMES[arm], got, q)

    # GUPS atomic_add (C5a) and TeamPolicy CRS SpMV (C5b)
    table_len, m = 1 << 16, 1 << 20
    idx = W.c5_indices(m, table_len)
    exp = np.zeros(table_len, dtype=np.int64)
    np.add.at(exp, idx, 3)
    didx = t(idx)
    for arm in (A.B200, A.CUDA):
        table = torch.zeros(table_len, dtype=torch.int64, device=dev)
        arms.gups_add(arm, table.data_ptr(), table_len, didx.data_ptr(), m, 3)
        assert np.array_equal(table.cpu().numpy(), exp), (A.ARM_NAMES[arm], "gups")
    nrows = 5000
    rm, ci, va, x = W.c5_crs(nrows, 32, integer_valued=True)
    exp_y = port.spmv(rm, ci, va, x)
    drm, dci, dva, dx = t(rm), t(ci), t(va), t(x)
    for arm in (A.B200, A.CUDA):
        y = torch.zeros(nrows, dtype=torch.float64, device=dev)
        arms.spmv(arm, nrows, drm.data_ptr(), dci.data_ptr(), dva.data_ptr(), dx.data_ptr(), y.data_ptr(), ci.size, x.size)
        assert np.array_equal(y.cpu().numpy(), exp_y), (A.ARM_NAMES[arm], "spmv")
    torch.cuda.synchronize()
    arms.finalize()
    print("arms ok")


if __name__ == "__main__":
    main()
