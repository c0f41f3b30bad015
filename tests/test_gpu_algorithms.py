"""GPU parity tests of the std-algorithm layer, ViewFill and the Crs row-map helper (kokkos_b200/include/kb200/StdAlgorithms.hpp;
SURVEY.md 8f ranks 2-3) through tests/cxx/cases_algorithms.cu.  Expected values are numpy restatements of the std:: semantics
the reference's own tests use as gold (algorithms/unit_tests/TestStdAlgorithms*.cpp compute their gold with plain loops):
integer results and index results bit-exact, double sums exact on the integer-valued inputs used here."""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_uint32, c_void_p

import numpy as np
import pytest

import workloads as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def alg():
    import kokkos_b200 as kb
    kb.load_library()
    L = ctypes.CDLL(kb.CASES_LIB_PATH)
    L.kb200_alg_last_error.restype = c_char_p
    L.kb200_case_last_error.restype = c_char_p
    assert L.kb200_case_init(0) == 0, L.kb200_case_last_error()
    yield L
    L.kb200_case_finalize()


def P(a):
    return a.ctypes.data_as(c_void_p)


def ok(L, rc):
    assert rc == 0, (rc, L.kb200_alg_last_error())


def excl(x, init, op=np.add):
    acc = op.accumulate(np.concatenate(([init], x)).astype(x.dtype))
    return acc[:-1]


@pytest.mark.parametrize("n", (1, 31, 1000, 100003, (1 << 20) + 5))
def test_scans_default_operator_typed_and_generic(alg, n):
    xi = W.c3_wrap(n)
    y = np.empty_like(xi)
    ok(alg, alg.kb200_alg_scan_i64(0, P(xi), P(y), c_int64(n), c_int64(11)))
    assert np.array_equal(y, np.cumsum(xi) - xi + 11)
    ok(alg, alg.kb200_alg_scan_i64(1, P(xi), P(y), c_int64(n), c_int64(0)))
    assert np.array_equal(y, np.cumsum(xi))
    ok(alg, alg.kb200_alg_scan_i64(4, P(xi), P(y), c_int64(n), c_int64(-3)))           # in place
    assert np.array_equal(y, np.cumsum(xi) - xi - 3)
    x32 = (W.c3_small(n) * 5).astype(np.int32)
    y32 = np.empty_like(x32)
    ok(alg, alg.kb200_alg_scan_i32(0, P(x32), P(y32), c_int64(n), c_int(7)))
    assert np.array_equal(y32, (np.cumsum(x32, dtype=np.int64) - x32 + 7).astype(np.int32))
    ok(alg, alg.kb200_alg_scan_i32(1, P(x32), P(y32), c_int64(n), c_int(0)))           # inclusive int32: generic lambda path
    assert np.array_equal(y32, np.cumsum(x32, dtype=np.int64).astype(np.int32))
    xd = W.c1_exact(n)                                                                  # integer-valued doubles: exact in any order
    yd = np.empty_like(xd)
    ok(alg, alg.kb200_alg_scan_f64(0, P(xd), P(yd), c_int64(n), c_double(0.5)))
    assert np.array_equal(yd, np.cumsum(xd) - xd + 0.5)
    ok(alg, alg.kb200_alg_scan_f64(1, P(xd), P(yd), c_int64(n), c_double(0.0)))
    assert np.array_equal(yd, np.cumsum(xd))
    if n <= 100003:                                                                     # float: sums stay below 2^24
        xf = (W.c1_exact(n) % 3).astype(np.float32)
        yf = np.empty_like(xf)
        ok(alg, alg.kb200_alg_scan_f32(0, P(xf), P(yf), c_int64(n), c_float(1.0)))
        assert np.array_equal(yf, (np.cumsum(xf, dtype=np.float64) - xf + 1.0).astype(np.float32))


@pytest.mark.parametrize("n", (1, 257, 70001))
def test_scans_custom_operator(alg, n):
    xd = W.c1_uniform(n)
    yd = np.empty_like(xd)
    ok(alg, alg.kb200_alg_scan_f64(2, P(xd), P(yd), c_int64(n), c_double(-0.25)))       # exclusive, op = max
    assert np.array_equal(yd, excl(xd, -0.25, np.maximum))
    ok(alg, alg.kb200_alg_scan_f64(3, P(xd), P(yd), c_int64(n), c_double(0)))           # inclusive, op = max
    assert np.array_equal(yd, np.maximum.accumulate(xd))
    xu = (W.hash_u32(np.arange(n, dtype=np.uint64)) % 1000 + 2).astype(np.uint32)       # op = modular product: associative, no identity used
    yu = np.empty_like(xu)
    exp = np.empty(n, dtype=np.uint32)
    run = 17
    for inclusive in (0, 1):
        ok(alg, alg.kb200_alg_scan_u32_mulmod(P(xu), P(yu), c_int64(n), c_uint32(17), inclusive))
        run = 17 if not inclusive else None
        for i in range(n):
            if inclusive:
                run = int(xu[i]) if run is None else (run * int(xu[i])) % 1000003
                exp[i] = run
            else:
                exp[i] = run
                run = (run * int(xu[i])) % 1000003
        assert np.array_equal(yu, exp)


@pytest.mark.parametrize("n", (1, 1000, 100003))
def test_reductions_and_element_queries(alg, n):
    x = (W.c1_exact(n) - 37.0) / 8.0                         # multiples of 1/8 in [-4.625, 7.75]: sums exact, many ties
    out = np.zeros(8)
    iout = np.zeros(8, dtype=np.int64)
    ok(alg, alg.kb200_alg_reductions(P(x), c_int64(n), P(out), P(iout)))
    assert out[0] == x.sum() and out[1] == x.sum() + 2.5
    assert out[2] == x.max()
    assert out[3] == float(np.dot(x, x))
    assert out[4] == float((x * x).max())
    assert out[5] == float(np.trunc(x * 8.0).sum())
    assert out[6] == float((x < 0).sum())
    assert iout[0] == int(np.trunc(x * 1000.0).astype(np.int64).sum())
    assert iout[1] == int(np.argmin(x))                      # first smallest
    assert iout[2] == int(np.argmax(x))                      # first largest
    assert iout[3] == int(np.argmin(x))
    assert iout[4] == int(n - 1 - np.argmax(x[::-1]))        # LAST largest (std::minmax_element)
    neg = np.nonzero(x < 0)[0]
    assert iout[5] == (int(neg[0]) if neg.size else n)
    assert iout[6] == int(np.trunc(x * 100.0).astype(np.int64).sum())
    assert iout[7] == n


def test_fill_copy_transform_viewfill(alg):
    n = 100003
    d = np.zeros(n); c = np.zeros(n); sq = np.zeros(n)
    iv = np.zeros(n, dtype=np.int32)
    b = np.zeros(n, dtype=np.uint8)
    ok(alg, alg.kb200_alg_elementwise(c_int64(n), P(d), P(iv), P(c), P(sq), P(b)))
    assert np.all(d == 4.25) and np.all(c == 4.25) and np.all(sq == 4.25 * 4.25)
    assert np.all(iv == 7) and np.all(b == 201)


@pytest.mark.parametrize("n", (1, 1000, 250007))
def test_crs_row_map_from_counts(alg, n):
    counts = (W.hash_u32(np.arange(n, dtype=np.uint64)) % 9).astype(np.int64)
    rm64 = np.zeros(n + 1, dtype=np.int64)
    rm32 = np.zeros(n + 1, dtype=np.int32)
    rmu = np.zeros(n + 1, dtype=np.uint32)
    totals = np.zeros(3, dtype=np.int64)
    ok(alg, alg.kb200_alg_crs_row_map(P(counts), c_int64(n), P(rm64), P(rm32), P(rmu), P(totals)))
    exp = np.concatenate(([0], np.cumsum(counts)))
    assert np.array_equal(rm64, exp) and np.array_equal(rm32.astype(np.int64), exp) and np.array_equal(rmu.astype(np.int64), exp)
    assert list(totals) == [exp[-1]] * 3
