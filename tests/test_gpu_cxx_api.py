"""GPU parity tests of the GENERIC C++ layer (kokkos_b200/include/kb200): Kokkos-style user code -- Views, policies,
KB200_LAMBDA functors, reducers, nested team parallelism, atomics -- compiled into libkokkos_b200_cases.so
(kokkos_b200/csrc/cases_api.cu) and driven here with host buffers; results are compared with the oracle port,
with closed forms of the reference's unit tests, or with a pure-numpy restatement for the cases the C oracle
does not cover (non-commutative scan, MDRange coverage, team collectives)."""
import ctypes
import os
from ctypes import POINTER, c_double, c_int, c_int32, c_int64, c_uint64, c_void_p, c_char_p

import numpy as np
import pytest

import workloads as W

pytestmark = pytest.mark.gpu
T = 4
F64_MAX = np.finfo(np.float64).max


@pytest.fixture(scope="module")
def cases():
    import kokkos_b200 as kb
    kb.load_library()
    assert os.path.exists(kb.CASES_LIB_PATH), "libkokkos_b200_cases.so not built"
    L = ctypes.CDLL(kb.CASES_LIB_PATH)
    L.kb200_case_last_error.restype = c_char_p
    assert L.kb200_case_init(0) == 0, L.kb200_case_last_error()
    yield L
    L.kb200_case_finalize()


def P(a):
    return a.ctypes.data_as(c_void_p)


def ok(L, rc):
    assert rc == 0, (rc, L.kb200_case_last_error())


def reduce_f64(L, op, x, base=0):
    out = np.zeros(4)
    loc = np.zeros(2, dtype=np.int64)
    ok(L, L.kb200_case_reduce_f64(op, P(x), c_int64(x.size), c_int64(base), P(out), P(loc)))
    return out, loc


@pytest.mark.parametrize("n", (0, 1, 33, 1000, 100003, 1 << 21))
def test_range_reduce_lambdas_and_reducers(cases, port, n):
    x = W.c1_exact(n)
    exact = float(x.astype(np.int64).sum())
    for op in (0, 1, 8):                       # scalar Sum, Sum<> reducer, device-View result
        assert reduce_f64(cases, op, x)[0][0] == exact
    assert reduce_f64(cases, 9, x)[0][0] == exact            # work tag + IndexType<int> + Dynamic + LaunchBounds
    assert reduce_f64(cases, 10, x)[0][0] == float((x.astype(np.int64) ** 2).sum()) + 1.0   # init/join/final functor
    out, loc = reduce_f64(cases, 11, x)                      # 32-byte struct value
    assert (out[0], out[1], loc[0]) == (exact, float((x.astype(np.int64) ** 2).sum()), n)
    assert reduce_f64(cases, 12, x)[0][0] == float(x[n // 3:].astype(np.int64).sum())      # non-zero begin
    xu = W.c1_uniform(n)
    assert reduce_f64(cases, 2, xu)[0][0] == port.reduce("min", xu, T)
    assert reduce_f64(cases, 3, xu)[0][0] == port.reduce("max", xu, T)
    xt = (W.hash_u32(np.arange(n, dtype=np.uint64)) % np.uint64(5)).astype(np.float64)    # ties
    for op, kind in ((4, "minloc"), (5, "maxloc")):
        out, loc = reduce_f64(cases, op, xt, 10)
        q = port.reduce_loc(kind, xt, 10, T)
        assert (out[0], loc[0]) == (q.val, q.loc)
    out, _ = reduce_f64(cases, 6, xu)
    q = port.reduce_minmax(xu, T)
    assert (out[0], out[1]) == (q.min_val, q.max_val)
    out, loc = reduce_f64(cases, 7, xt, 3)
    q = port.reduce_loc("minmaxloc", xt, 3, T)
    assert (out[0], out[1], loc[0], loc[1]) == (q.min_val, q.max_val, q.min_loc, q.max_loc)
    if n:
        m = min(n, 20)
        assert reduce_f64(cases, 13, x)[0][0] == float(np.prod([(i % 3 + 2) for i in range(m)], dtype=np.float64))


@pytest.mark.parametrize("n", (0, 7, 4097, 100003))
def test_range_reduce_int32_redux_and_bitwise(cases, port, n):
    x = (W.c3_wrap(n) >> 40).astype(np.int32)
    out = np.zeros(1, dtype=np.int32)
    ok(cases, cases.kb200_case_reduce_i32(0, P(x), c_int64(n), P(out)))
    assert out[0] == port.reduce("sum", x, T)
    ok(cases, cases.kb200_case_reduce_i32(1, P(x), c_int64(n), P(out)))
    assert out[0] == port.reduce("min", x, T)
    ok(cases, cases.kb200_case_reduce_i32(2, P(x), c_int64(n), P(out)))
    assert out[0] == port.reduce("max", x, T)
    ok(cases, cases.kb200_case_reduce_i32(3, P(x), c_int64(n), P(out)))
    assert out[0] == (np.bitwise_and.reduce(x | np.int32(0x0f0f0000)) if n else np.int32(-1))
    ok(cases, cases.kb200_case_reduce_i32(4, P(x), c_int64(n), P(out)))
    assert out[0] == (np.bitwise_or.reduce(x & np.int32(0x00ff00ff)) if n else 0)
    ok(cases, cases.kb200_case_reduce_i32(5, P(x), c_int64(n), P(out)))
    assert out[0] == int(np.all(x != 12345678))
    ok(cases, cases.kb200_case_reduce_i32(6, P(x), c_int64(n), P(out)))
    assert out[0] == int(np.any(x == 3))


@pytest.mark.parametrize("n", (1, 1000, 100003))
def test_combined_multi_result_reduce(cases, n):
    """parallel_reduce(policy, f, r0, r1, ...) -- the CombinedReducer path (row a22): scalars, reducers and a device View."""
    x = W.c3_wrap(n) >> 20
    out = np.zeros(8, dtype=np.int64)
    n0, n1, n2 = 13, 7, 5
    ok(cases, cases.kb200_case_combined_reduce(P(x), c_int64(n), c_int64(n0), c_int64(n1), c_int64(n2), P(out)))
    assert (out[0], out[1], out[2]) == (int(x.sum()), int(x.min()), int(x.max()))
    i, j, k = np.meshgrid(np.arange(n0), np.arange(n1), np.arange(n2), indexing="ij")
    assert out[3] == int((i + 10 * j + 100 * k).sum()) and out[4] == n0 * n1 * n2
    assert out[5] == 2 * int(x.sum()) and out[6] == n


@pytest.mark.parametrize("n", (0, 1, 2303, 2304, 2305, 100003, 1 << 21))
def test_generic_scan_lambda(cases, port, n):
    x = W.c3_wrap(n)
    for incl in (0, 1):
        y = np.zeros(n, dtype=np.int64)
        total = c_int64(-1)
        ok(cases, cases.kb200_case_scan_i64(P(x), P(y), c_int64(n), incl, ctypes.byref(total)))
        py, pt = port.scan(x, bool(incl), 0, T)
        assert total.value == pt and np.array_equal(y, py)
    xd = W.c1_exact(n)                       # exact doubles, total into a device View, IndexType<int>
    y = np.zeros(n)
    total = c_double(-1)
    ok(cases, cases.kb200_case_scan_f64_to_view_total(P(xd), P(y), c_int64(n), ctypes.byref(total)))
    py, pt = port.scan(xd, False, 0.0, T)
    assert total.value == pt and np.array_equal(y, py)


@pytest.mark.parametrize("n", (1, 300, 2305, 30011))
def test_generic_scan_noncommutative_join_is_index_ordered(cases, n):
    """Affine maps over Z/2^64 compose associatively but not commutatively: any mis-ordered combine shows."""
    a = W.hash64(np.arange(n, dtype=np.uint64), 3) | np.uint64(1)
    b = W.hash64(np.arange(n, dtype=np.uint64), 4)
    for incl in (0, 1):
        ya = np.zeros(n, dtype=np.uint64); yb = np.zeros(n, dtype=np.uint64); tot = np.zeros(2, dtype=np.uint64)
        ok(cases, cases.kb200_case_scan_affine(P(a), P(b), c_int64(n), incl, P(ya), P(yb), P(tot)))
        A, B = 1, 0
        M = (1 << 64) - 1
        ea = np.zeros(n, dtype=np.uint64); eb = np.zeros(n, dtype=np.uint64)
        for i in range(n):                      # pure-Python restatement, small n only
            if not incl:
                ea[i], eb[i] = A, B
            A, B = (int(a[i]) * A) & M, (int(a[i]) * B + int(b[i])) & M
            if incl:
                ea[i], eb[i] = A, B
        assert np.array_equal(ya, ea) and np.array_equal(yb, eb) and (int(tot[0]), int(tot[1])) == (A, B)


def test_parallel_for_stream_lambdas(cases, port):
    n = 100003
    a = np.full(n, 1.0); b = np.full(n, 2.0); c = np.zeros(n)
    ha, hb, hc = a.copy(), b.copy(), c.copy()
    ok(cases, cases.kb200_case_stream(P(a), P(b), P(c), c_int64(n), 5, c_double(3.0)))
    for _ in range(5):
        port.stream("copy", ha.ctypes.data, hc.ctypes.data, n)
        port.stream("scale", hb.ctypes.data, hc.ctypes.data, 3.0, n)
        port.stream("add", ha.ctypes.data, hb.ctypes.data, hc.ctypes.data, n)
        port.stream("triad", ha.ctypes.data, hb.ctypes.data, hc.ctypes.data, 3.0, n)
    assert np.array_equal(a, ha) and np.array_equal(b, hb) and np.array_equal(c, hc)


@pytest.mark.parametrize("dims", ((8, 9, 10), (34, 20, 18), (64, 64, 64), (130, 7, 5)))
def test_mdrange_stencil_minmaxloc_lambda(cases, port, dims):
    u, _, _ = W.c4_field(*dims) if min(dims) > 6 else (W.c1_uniform(dims[0] * dims[1] * dims[2]), None, None)
    v = np.zeros_like(u)
    out = np.zeros(2); loc = np.zeros(2, dtype=np.int64)
    ok(cases, cases.kb200_case_mdrange_stencil(P(u), P(v), c_int64(dims[0]), c_int64(dims[1]), c_int64(dims[2]),
                                               c_double(0.5), c_double(0.125), P(out), P(loc)))
    q, pv = port.stencil7(u, *dims, 0.5, 0.125, want_v=True)
    assert (out[0], out[1], loc[0], loc[1]) == (q.min_val, q.max_val, q.min_loc, q.max_loc)
    assert np.array_equal(v, pv)


@pytest.mark.parametrize("rank,lower,upper,tile", [
    (2, (0, 0), (100, 37), None), (2, (-5, 3), (60, 40), (8, 4)), (3, (0, 0, 0), (33, 17, 9), None),
    (3, (1, -2, 3), (40, 11, 20), (16, 2, 4)), (4, (0, 0, 0, 0), (9, 8, 7, 6), None), (4, (0, 1, 2, 3), (12, 9, 8, 7), (4, 2, 2, 3)),
    (5, (0, 0, 0, 0, 0), (7, 6, 5, 4, 3), None), (6, (0, 0, 0, 0, 0, 0), (5, 4, 3, 4, 3, 2), (4, 2, 2, 2, 1, 2)),
    (3, (0, 0, 0), (0, 5, 5), None),                      # zero-length dimension: nothing runs, reduce = 0
    # large enough for the thread-coarsened kernels (4 consecutive slowest-dimension indices per thread), ragged ends
    (2, (-7, 3), (1990, 3002), None), (3, (-3, 2, -5), (125, 66, 1994), None), (3, (0, 0, 0), (64, 33, 2001), (32, 2, 1)),
])
def test_mdrange_every_point_exactly_once(cases, rank, lower, upper, tile):
    ext = [u - l for l, u in zip(lower, upper)]
    total = int(np.prod(ext))
    hits = np.zeros(max(total, 1), dtype=np.int32)
    lo = np.array(lower, dtype=np.int64); up = np.array(upper, dtype=np.int64)
    tl = np.array(tile if tile else [0] * rank, dtype=np.int64)
    poly = c_int64(-1)
    ok(cases, cases.kb200_case_mdrange(rank, P(lo), P(up), P(tl), int(tile is not None), P(hits), c_int64(max(total, 1)), ctypes.byref(poly)))
    if total:
        assert np.all(hits[:total] == 1)
    grids = np.meshgrid(*[np.arange(l, u, dtype=np.int64) for l, u in zip(lower, upper)], indexing="ij")
    coef = (1, 3, 5, 7, 11, 13)
    assert poly.value == int(sum(c * g.sum() for c, g in zip(coef, grids)))


@pytest.mark.parametrize("team,vec,rpt", ((0, 32, 8), (8, 32, 8), (16, 8, 32), (64, 4, 64), (128, 1, 256)))
def test_team_policy_spmv_nested_reduce(cases, port, team, vec, rpt):
    nrows = 3000
    for iv in (True, False):
        rm, ci, va, x = W.c5_crs(nrows, 32, integer_valued=iv)
        y = np.zeros(nrows)
        ok(cases, cases.kb200_case_team_spmv(c_int64(nrows), P(rm), P(ci), P(va), c_int64(va.size), P(x), c_int64(x.size), P(y), rpt, team, vec))
        py = port.spmv(rm, ci, va, x)
        if iv:
            assert np.array_equal(y, py)
        else:
            mag = np.abs(va.reshape(nrows, 32) * x[ci.reshape(nrows, 32)]).sum(axis=1)
            assert np.max(np.abs(y - py) / mag) <= 1e-12


@pytest.mark.parametrize("league,team,vec,n_inner", ((1, 1, 1, 5), (7, 32, 1, 100), (50, 64, 4, 333), (300, 16, 32, 77), (13, 256, 2, 1000)))
def test_team_collectives_scratch_and_nested_scans(cases, league, team, vec, n_inner):
    out = np.zeros(16, dtype=np.int64)
    ok(cases, cases.kb200_case_team_collectives(league, team, vec, n_inner, P(out)))
    assert list(out[:8]) == [0] * 8, out[:8]          # in-kernel checks
    expect = sum(n_inner * (n_inner - 1) // 2 + n_inner * lr for lr in range(league))
    assert out[8] == expect                            # league-level reduce, one contribution per team
    assert out[9] == league * team                     # team_scan global accumulator
    assert out[10] == 0                                # nested TeamThreadRange scan through level-0 scratch


def test_multilevel_scratch_carving(cases):
    """Successive team_scratch(l)/thread_scratch(l) calls carve successive, non-overlapping pieces at both levels (TestTeam.hpp:920-1036)."""
    out = np.zeros(4, dtype=np.int64)
    ok(cases, cases.kb200_case_multilevel_scratch(10, 8, 16, P(out)))
    assert list(out) == [0, 0, 0, 0]


def test_atomics_all_ops(cases):
    n = 100003
    out = np.zeros(32)
    ok(cases, cases.kb200_case_atomics(c_int64(n), P(out)))
    i = np.arange(n, dtype=np.int64)
    raw = out.view(np.int64)
    assert raw[24] == int(i.sum()) and out[1] == n
    assert out[2] == min((1 << 30), int(((i * 7919) % 10007 + 5).min())) and out[3] == int(((i * 7919) % 10007).max())
    assert raw[25] == int(np.bitwise_and.reduce(~(np.int64(1) << (i % 40))))
    assert raw[26] == int(np.bitwise_or.reduce(np.int64(1) << (i % 50)))
    with np.errstate(over="ignore"):
        assert raw[27] == int(np.bitwise_xor.reduce((i.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)).view(np.int64)))
    assert out[8] == n and out[9] == -2 * n
    assert out[10] == int(((i * 31) % 977 + 3).min()) and out[11] == int(((i * 31) % 977).max())
    assert out[16] == 0.5 * n and out[17] == float(((i * 13) % 1000).min()) - 3.25 and out[18] == float(((i * 13) % 1000).max()) + 0.75
    assert out[19] == float(n)                         # float add of ones: exact below 2^24
    assert [int(v) for v in out[20:24]] == [int(np.sum(i % 4 == k)) % 256 for k in range(4)]
    assert out[13] == 77 and out[14] == 99 and out[15] == 5


def test_views_deep_copy_subview_mirrors(cases):
    n = 1000
    out = np.zeros(16, dtype=np.int64)
    ok(cases, cases.kb200_case_views(c_int64(n), P(out)))
    i = np.arange(n, dtype=np.int64)
    assert out[0] == 1 and out[1] == 2
    assert out[2] == int((i[2:n - 1] ** 2).sum()) and out[3] == (n - 1) ** 2 and out[4] == 0
    assert out[5] == 5 and out[6] == 7 and out[7] == 1
    assert out[8] == 2048 * 148 and out[9] == 1


def test_array_policy_arrays_resultless_reduce_user_reducer(cases):
    """Kokkos::Array, MDRangePolicy from Arrays (TestMDRangePolicyConstructors.hpp:38-77), parallel_reduce without a result and a
    user-written reducer whose result_view_type is on the host / device (TestReduceCombinatorical.hpp:27-56,460-490)."""
    n = 100003
    x = np.random.default_rng(5).integers(-1000, 1000, n, dtype=np.int64)
    out = np.zeros(16, dtype=np.int64)
    ok(cases, cases.kb200_case_utilities(P(x), c_int64(n), P(out)))
    assert out[0] == int(x.sum()) + 7 and out[1] == int(x[1:].sum()) + 7
    assert out[2] == int(x.sum()) and out[3] == 2 * int(x.sum())
    assert out[4] == 8 * 10000 + 4 * 100 + 4           # first dimension as given, the rest the device default
    assert out[5] == 32 * 4 * 4 and out[6] == 1024
    i, j, k = np.meshgrid(np.arange(1, 41), np.arange(2, 22), np.arange(3, 13), indexing="ij")
    assert out[7] == int((i + 100 * j + 10000 * k).sum())
    assert out[8] == 0
    assert list(out[9:12]) == [0, 0, 0]                # empty Range / MDRange / Team reduce: identity, no join


def test_strided_subview_copy_fill(cases):
    """subview(view, range, index, ALL) -> LayoutStride View sharing the allocation; element-wise ViewCopy / ViewFill through strides
    (Kokkos_CopyViews.hpp:300-560; TestMDRange_g.hpp:41-72 is the reference's use)."""
    n0, n1, n2 = 37, 11, 23
    out = np.zeros(16, dtype=np.int64)
    ok(cases, cases.kb200_case_strided_subview(c_int64(n0), c_int64(n1), c_int64(n2), P(out)))
    assert out[0] == (n0 - 2) * 1000 + n2 and out[1] == 1 * 1000000 + n0 * n1 and out[2] == 0 and out[3] == 2
    assert out[4] == 0 and out[5] == 0
    assert out[6] == (n0 - 2) * n2 and out[7] == 0
    assert out[8] == sum(10 * i + 4 for i in range(6)) and out[9] == 9


def test_unique_token_and_resize(cases):
    """UniqueToken: every acquire() hands out an index nobody else holds, default- and caller-sized (TestUniqueToken.hpp:36-70); resize keeps the common index box, realloc zero-fills (Kokkos_CopyViews.hpp:1580-1790)."""
    n = 200003
    out = np.zeros(16, dtype=np.int64)
    ok(cases, cases.kb200_case_tokens_resize(c_int64(n), P(out)))
    assert out[0] == 2048 * 148 and out[1] == 0 and out[2] == n
    assert out[3] == 2 * n and out[4] == n * (n + 1) // 2
    h = n // 2
    assert out[5] == h and out[6] == h * (h + 1) // 2
    assert out[7] == 84 and out[8] == sum(10 * i + j for i in range(5) for j in range(4)) and out[9] == 0



@pytest.mark.parametrize("count", [65, 300, 4097])
def test_runtime_length_array_reduce_of_any_length(cases, count):
    """value_type[] reductions above the 64-element register capacity: accumulators in global memory; Range with
    the default init/join (+=), MDRange with the functor's own init/join/final."""
    n, n0, n1 = 200003, 97, 53
    out_r = np.zeros(count, dtype=np.int64)
    out_m = np.zeros(count, dtype=np.int64)
    ok(cases, cases.kb200_case_array_reduce_big(c_int64(n), c_int(count), c_int64(n0), c_int64(n1), P(out_r), P(out_m)))
    i = np.arange(n, dtype=np.int64)
    exp_r = np.bincount(i % count, weights=None, minlength=count) * 0
    np.add.at(exp_r, i % count, i)
    assert np.array_equal(out_r, exp_r)
    ii, jj = np.meshgrid(np.arange(n0), np.arange(n1), indexing="ij")
    exp_m = np.zeros(count, dtype=np.int64)
    np.add.at(exp_m, ((ii * 7 + jj) % count).ravel(), (1 + jj).ravel())
    exp_m[0] += 1000000
    assert np.array_equal(out_m, exp_m)


def test_closure_larger_than_the_kernel_parameter_space(cases):
    """A 40 KB functor cannot be a kernel argument: parallel_for copies it into the instance's functor scratch in stream order and
    the kernel reads it from global memory (the reference's 'global memory launch')."""
    n = 100003
    out = c_int64()
    ok(cases, cases.kb200_case_huge_closure(c_int64(n), ctypes.byref(out)))
    i = np.arange(n, dtype=np.int64)
    ballast = ((np.arange(40000, dtype=np.int64) * 7 + 1) % 256)
    assert out.value == int((i * 3 + ballast[i % 40000]).sum())


def test_mdrange_tile_beyond_blockdim_z(cases):
    """A rank-6 tile {2,2,2,2,2,16}: the four slow extents multiply to 128 > 64 (blockDim.z's limit).  This is the reference's own
    default tiling for a rank-6 Iterate::Right policy, which its ViewFill of a rank >= 6 LayoutRight View launches; the kernel
    layer runs it as a linear block.  parallel_for visits every point exactly once; parallel_reduce agrees."""
    total = 5 * 3 * 4 * 3 * 5 * 37
    s, r, pts = c_int64(), c_int64(), c_int64()
    ok(cases, cases.kb200_case_mdrange_wide_tile(ctypes.byref(s), ctypes.byref(r), ctypes.byref(pts)))
    exp = total * (total + 1) // 2
    assert s.value == exp and r.value == exp and pts.value == total
