"""MDRange default-tile fallback for a register-heavy functor (kept in its own, last-sorted file: it is the newest GPU test of the
round and must not stand in front of the established suites under `pytest -x`)."""
import numpy as np
import pytest

from test_gpu_cxx_api import cases, ok, P, c_int64  # noqa: F401  (fixture + helpers)

pytestmark = pytest.mark.gpu


def test_mdrange_default_tile_shrinks_for_register_heavy_functor(cases):
    """A functor with 72 live 64-bit values cannot launch on the default 32x4x4 tile (128 registers per thread); the launcher shrinks
    a DEFAULT tile until the kernel fits (the reference sizes its default tile from cudaFuncAttributes: KokkosExp_MDRangePolicy.hpp:98-128)."""
    n0, n1, n2 = 40, 9, 6
    hy = np.zeros(n0 * n1 * n2, dtype=np.uint64)
    ok(cases, cases.kb200_case_mdrange_heavy(c_int64(n0), c_int64(n1), c_int64(n2), P(hy)))
    k, j, i = np.meshgrid(np.arange(n2, dtype=np.uint64), np.arange(n1, dtype=np.uint64), np.arange(n0, dtype=np.uint64), indexing="ij")
    base = i + np.uint64(3) * j + np.uint64(7) * k
    with np.errstate(over="ignore"):
        a = [base * np.uint64(t + 1) + np.uint64(0x9E3779B97F4A7C15) for t in range(72)]
        for r in range(3):
            for t in range(72):
                a[t] = a[t] * a[(t + 7) % 72] + np.uint64(r + 1)
        s = np.zeros_like(base)
        for t in range(72):
            s ^= a[t] + np.uint64(t)
    assert np.array_equal(hy.reshape(n2, n1, n0), s)
