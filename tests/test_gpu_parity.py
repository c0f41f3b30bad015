"""GPU parity tests proper: the hand-written sm_100a kernels, called through the C ABI
(include/kokkos_b200.h via kokkos_b200/__init__.py), against

  * the oracle port (oracle/omp_oracle.c) on the same seeded inputs, bit-exact for integer / index /
    scan / elementwise results and for order-independent doubles;
  * the committed golden vectors of the unmodified reference (tests/golden/golden_v1.json);
  * the unmodified reference itself where oracle/_ref travelled to this box;
  * for general double sums: |gpu - exact| <= 1e-12 * |exact| (north_star tolerance; `exact` is the
    correctly rounded sum), with the oracle's own deviation from exact shown beside it;
  * at BASELINE.json's full sizes: size-independent properties (closed forms, linearity, scan
    differences reproduce the input, checksum-of-checksums, idempotent stream recurrences).
"""
import json
import os

import numpy as np
import pytest

import workloads as W

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.json")))
T = GOLD["threads"]
C = GOLD["cases"]
SIZES = (0, 1, 2, 5, 33, 1000, 4097, 100003, 1 << 20)
REL_TOL = 1e-12  # north_star: <= 1e-12 relative for double sums, against reassociation
F64_MAX = np.finfo(np.float64).max
I64_MAX = np.iinfo(np.int64).max


def test_extension_is_the_cuda_library(space):
    import kokkos_b200 as kb
    p = space.props()
    assert p.cc_major == 10 and p.sm_count >= 100, (p.cc_major, p.sm_count)
    assert os.path.exists(kb.LIB_PATH)


# ------------------------------------------------------------------ parallel_reduce
@pytest.mark.parametrize("n", SIZES)
def test_reduce_matches_golden_and_port(space, port, n):
    # integer-valued doubles: bit-exact in any order (the reference's own trick, TestReducers.hpp:472)
    x = W.c1_exact(n)
    got = space.parallel_reduce_sum(space.view_from_host(x))
    assert float(got).hex() == C[f"sum_f64/c1_exact/{n}"] == float(port.reduce("sum", x, T)).hex()
    # int64 with wrap-around: bit-exact
    xi = W.c3_wrap(n)
    v = space.view_from_host(xi)
    assert space.parallel_reduce_sum(v) == C[f"sum_i64/c3_wrap/{n}"]
    assert space.parallel_reduce_min(v) == C[f"min_i64/c3_wrap/{n}"]
    assert space.parallel_reduce_max(v) == C[f"max_i64/c3_wrap/{n}"]
    # min/max/minmaxloc of doubles: no reassociation => bit-exact
    xu = W.c1_uniform(n)
    vu = space.view_from_host(xu)
    assert float(space.parallel_reduce_min(vu)).hex() == C[f"min_f64/c1_uniform/{n}"]
    assert float(space.parallel_reduce_max(vu)).hex() == C[f"max_f64/c1_uniform/{n}"]
    r = space.parallel_reduce_minmaxloc(vu, 0)
    assert [float(r.min_val).hex(), float(r.max_val).hex(), r.min_loc, r.max_loc] == C[f"minmaxloc_f64/c1_uniform/{n}"]
    # ties: OpenMP's effective rule is "lowest index wins"; the B200 join implements it for any grid
    xt = (W.hash_u32(np.arange(n, dtype=np.uint64)) % np.uint64(5)).astype(np.float64)
    r = space.parallel_reduce_minmaxloc(space.view_from_host(xt), 10)
    assert [float(r.min_val).hex(), float(r.max_val).hex(), r.min_loc, r.max_loc] == C[f"minmaxloc_f64/ties/{n}"]
    for kind in ("minloc", "maxloc"):
        a = getattr(space, f"parallel_reduce_{kind}")(space.view_from_host(xt), 3)
        b = port.reduce_loc(kind, xt, 3, T)
        assert (a.val, a.loc) == (b.val, b.loc)
    mm, pm = space.parallel_reduce_minmax(vu), port.reduce_minmax(xu, T)
    assert (mm.min_val, mm.max_val) == (pm.min_val, pm.max_val)


@pytest.mark.parametrize("n", (1000, 100003, 1 << 22))
@pytest.mark.parametrize("gen", ("c1_general", "c1_uniform"))
def test_reduce_general_doubles_within_tolerance(space, port, gen, n):
    x = getattr(W, gen)(n)
    exact = W.exact_sum(x)
    got = space.parallel_reduce_sum(space.view_from_host(x))
    scale = max(abs(exact), float(np.sum(np.abs(x))) * 1e-3)  # uniform(-1,1) sums sit near 0: scale by magnitude
    oracle = port.reduce("sum", x, T)
    assert abs(got - exact) <= REL_TOL * scale, (got, exact, "oracle deviation", abs(oracle - exact) / scale)


@pytest.mark.parametrize("n", (0, 1, 31, 4097, 100003))
def test_reduce_other_types_and_alignment(space, port, n):
    x32 = W.c3_small(n).astype(np.int32)
    v = space.view_from_host(x32)
    assert space.parallel_reduce_sum(v) == port.reduce("sum", x32, T)
    assert space.parallel_reduce_min(v) == port.reduce("min", x32, T)
    assert space.parallel_reduce_max(v) == port.reduce("max", x32, T)
    xf = W.c1_exact(n).astype(np.float32)  # small integers: exact in float32 up to 2^24
    if n <= 100003:
        assert space.parallel_reduce_sum(space.view_from_host(xf)) == port.reduce("sum", xf, T)
    # unaligned subviews exercise the scalar head/tail edges of the 32-byte vector path
    x = W.c1_exact(n + 7)
    big = space.view_from_host(x)
    for off in (1, 2, 3):
        sub = big.subview(off, off + n)
        assert space.parallel_reduce_sum(sub) == float(x[off:off + n].astype(np.int64).sum())
        r = space.parallel_reduce_minmaxloc(sub, 0)
        q = port.reduce_loc("minmaxloc", np.ascontiguousarray(x[off:off + n]), 0, T)
        assert (r.min_val, r.max_val, r.min_loc, r.max_loc) == (q.min_val, q.max_val, q.min_loc, q.max_loc)


def test_reduce_empty_range_yields_identity(space):
    v = space.view(0, np.float64)
    assert space.parallel_reduce_sum(v) == 0.0
    assert space.parallel_reduce_min(v) == F64_MAX and space.parallel_reduce_max(v) == -F64_MAX
    r = space.parallel_reduce_minmaxloc(v)
    assert (r.min_val, r.max_val, r.min_loc, r.max_loc) == (F64_MAX, -F64_MAX, I64_MAX, I64_MAX)


def test_reduce_result_in_device_view_is_async(space):
    x = W.c1_exact(100003)
    v = space.view_from_host(x)
    out = space.view(1, np.float64)
    space.parallel_reduce_sum(v, result_dev=out.ptr, blocking=False)
    space.fence()
    assert out.to_host()[0] == float(x.sum())


def test_reduce_is_bitwise_reproducible(space):
    x = W.c1_uniform(1 << 22)
    v = space.view_from_host(x)
    vals = {float(space.parallel_reduce_sum(v)).hex() for _ in range(5)}
    assert len(vals) == 1


# ------------------------------------------------------------------ parallel_scan
@pytest.mark.parametrize("n", SIZES)
def test_scan_matches_golden_and_port(space, port, n):
    for gen in ("c3_small", "c3_wrap"):
        x = getattr(W, gen)(n)
        vx = space.view_from_host(x)
        for incl in (False, True):
            vy = space.view(n, np.int64)
            total = space.parallel_scan(vx, vy, inclusive=incl, seed=5)
            y = vy.to_host()
            g = C[f"scan_i64/{gen}/{'incl' if incl else 'excl'}/{n}"]
            assert total == g["total"] and W.checksum64(y) == g["checksum"] and [int(t) for t in y[-3:]] == g["tail"]
            py, pt = port.scan(x, incl, 5, T)
            assert np.array_equal(y, py) and total == pt
    # doubles holding small integers: exact, so bit-identical to the reference as well
    x = W.c1_exact(n)
    vy = space.view(n, np.float64)
    total = space.parallel_scan(space.view_from_host(x), vy, inclusive=False, seed=0.0)
    g = C[f"scan_f64/c1_exact/excl/{n}"]
    assert float(total).hex() == g["total"] and W.checksum64(vy.to_host()) == g["checksum"]


@pytest.mark.parametrize("n", (1, 2303, 2304, 2305, 4607, 4608, 4609, 9216, 1 << 20, (1 << 22) + 12345))
def test_scan_closed_form_tile_edges_inplace_and_unaligned(space, n):
    # TestParallelScanRangePolicy.hpp:66-84 closed forms; sizes straddle the 2304- and 4608-element tiles
    i = np.arange(n, dtype=np.int64)
    v = space.view_from_host(i)
    total = space.parallel_scan(v, v, inclusive=False)  # in place
    assert total == n * (n - 1) // 2 and np.array_equal(v.to_host(), i * (i - 1) // 2)
    v.from_host(i)
    assert space.parallel_scan(v, v, inclusive=True) == n * (n - 1) // 2
    assert np.array_equal(v.to_host(), i * (i + 1) // 2)
    # unaligned (8-byte but not 16-byte aligned) input and/or output: the non-TMA path
    big_in = space.view_from_host(np.concatenate([[0], i]))
    big_out = space.view(n + 1, np.int64)
    for sx, sy in ((1, 0), (0, 1), (1, 1)):
        x = big_in.subview(1, n + 1) if sx else space.view_from_host(i)
        y = big_out.subview(1, n + 1) if sy else big_out.subview(0, n)
        assert space.parallel_scan(x, y) == n * (n - 1) // 2
        assert np.array_equal(y.to_host(), i * (i - 1) // 2)
    x32 = (i % 5).astype(np.int32)
    y32 = space.view(n, np.int32)
    t = space.parallel_scan(space.view_from_host(x32), y32)
    ref = np.concatenate([[0], np.cumsum(x32[:-1], dtype=np.int64)]).astype(np.int32)
    assert t == int(x32.sum()) and np.array_equal(y32.to_host(), ref)


def test_scan_uniform_kernel_on_aligned_views(space, port):
    """The non-warp-specialised kernel normally serves unaligned Views only; force it on aligned ones too."""
    import kokkos_b200 as kb
    for ws in (0,):  # ws 1/3/4 are sweep-build variants (kb200/impl/ScanContigSweep.hpp)
      kb.tune_set("scan.ws", ws); kb.tune_set("scan.block", 256); kb.tune_set("scan.nbuf", 2); kb.tune_set("scan.lbw", 2)
      try:
        for n in (4608, 100003, 1 << 20):
            x = W.c3_wrap(n)
            vy = space.view(n, np.int64)
            total = space.parallel_scan(space.view_from_host(x), vy, seed=5)
            py, pt = port.scan(x, False, 5, T)
            assert total == pt and np.array_equal(vy.to_host(), py)
      finally:
        kb.tune_set("scan.ws", 2); kb.tune_set("scan.block", 128); kb.tune_set("scan.nbuf", 4); kb.tune_set("scan.lbw", 1)


def test_scan_many_launches_reuse_descriptor_arena(space):
    # epoch-tagged descriptors are never cleared: alternate sizes so stale entries would be hit
    rng = np.random.default_rng(3)
    for n in (1 << 20, 5000, 1 << 18, 4608 * 3, 1 << 20, 77):
        x = rng.integers(-1000, 1000, n, dtype=np.int64)
        vy = space.view(n, np.int64)
        total = space.parallel_scan(space.view_from_host(x), vy, inclusive=True)
        assert total == int(x.sum()) and np.array_equal(vy.to_host(), np.cumsum(x))


def test_scan_f64_general_within_tolerance(space):
    n = 1 << 20
    x = W.c1_general(n)
    vy = space.view(n, np.float64)
    total = space.parallel_scan(space.view_from_host(x), vy, inclusive=True)
    exact = np.cumsum(x.astype(np.longdouble))
    y = vy.to_host()
    assert np.max(np.abs(y - exact) / np.abs(exact)) <= REL_TOL
    assert abs(total - float(exact[-1])) <= REL_TOL * float(exact[-1])


# ------------------------------------------------------------------ parallel_for: stream
@pytest.mark.parametrize("n", (1, 3, 4099, 1 << 20))
def test_stream_kernels_bit_exact_vs_port_and_golden(space, port, n):
    a = space.view(n, np.float64); b = space.view(n, np.float64); c = space.view(n, np.float64)
    space.stream_set(a, 1.0); space.stream_set(b, 2.0); space.stream_set(c, 0.0)
    ha = np.full(n, 1.0); hb = np.full(n, 2.0); hc = np.zeros(n)
    P = lambda v: v.ctypes.data
    for _ in range(5):
        space.stream_copy(a, c); port.stream("copy", P(ha), P(hc), n)
        space.stream_scale(b, c, 3.0); port.stream("scale", P(hb), P(hc), 3.0, n)
        space.stream_add(a, b, c); port.stream("add", P(ha), P(hb), P(hc), n)
        space.stream_triad(a, b, c, 3.0); port.stream("triad", P(ha), P(hb), P(hc), 3.0, n)
    ga, gb, gc = a.to_host(), b.to_host(), c.to_host()
    assert np.array_equal(ga, ha) and np.array_equal(gb, hb) and np.array_equal(gc, hc)
    if n == 4099:
        assert [float(ga[0]).hex(), float(gb[0]).hex(), float(gc[0]).hex(), W.checksum64(ga), W.checksum64(gb),
                W.checksum64(gc)] == C["stream/5iters"]


def test_stream_triad_random_data_no_fma_contraction(space, port):
    n = 100003
    hb, hc = W.c1_uniform(n, 1), W.c1_uniform(n, 2)
    ha = np.zeros(n)
    port.stream("triad", ha.ctypes.data, hb.ctypes.data, hc.ctypes.data, 1.0 / 3.0, n)
    big = space.view(n + 4, np.float64)
    for off in (0, 1):  # aligned and mis-phased destination
        a = big.subview(off, off + n)
        space.stream_triad(a, space.view_from_host(hb), space.view_from_host(hc), 1.0 / 3.0)
        assert np.array_equal(a.to_host(), ha)


# ------------------------------------------------------------------ MDRange stencil + MinMaxLoc
@pytest.mark.parametrize("dims", ((8, 9, 10), (34, 20, 18), (64, 64, 64), (3, 3, 3), (130, 7, 5)))
def test_stencil_minmaxloc_bit_exact(space, port, dims):
    u, pmax, pmin = W.c4_field(*dims) if min(dims) > 6 else (W.c1_uniform(dims[0] * dims[1] * dims[2]), None, None)
    vout = space.view(u.size, np.float64)
    space.lib.b200_memset_async(space.handle, vout.ptr, 0, vout.nbytes)
    r = space.stencil7_minmaxloc(space.view_from_host(u), *dims, 0.5, 0.125, v_out=vout)
    q, pv = port.stencil7(u, *dims, 0.5, 0.125, want_v=True)
    assert (r.min_val, r.max_val, r.min_loc, r.max_loc) == (q.min_val, q.max_val, q.min_loc, q.max_loc)
    assert np.array_equal(vout.to_host(), pv)
    key = f"stencil7/{dims[0]}x{dims[1]}x{dims[2]}"
    if key in C:
        assert [float(r.min_val).hex(), float(r.max_val).hex(), r.min_loc, r.max_loc] == C[key]["minmaxloc"]
        assert W.checksum64(vout.to_host()) == C[key]["v_checksum"]


@pytest.mark.parametrize("dims", ((512, 11, 9), (4, 4, 4), (6, 3, 3), (258, 19, 37), (64, 10, 70), (33, 12, 9), (514, 5, 4)))
@pytest.mark.parametrize("store", (False, True))
def test_stencil_both_kernels_edge_shapes(space, port, dims, store):
    """Even n0 <= 512 takes the TMA plane-marching kernel (partial j tiles, short k chunks, i edges); odd / wide n0 the
    row-per-warp kernel.  Random data: every value and both locations bit-exact vs the oracle."""
    n = dims[0] * dims[1] * dims[2]
    u = W.c1_uniform(n, seed=dims[0] * 131 + dims[1])
    vout = space.view(n, np.float64) if store else None
    if store:
        space.lib.b200_memset_async(space.handle, vout.ptr, 0, vout.nbytes)
    r = space.stencil7_minmaxloc(space.view_from_host(u), *dims, 0.5, 0.125, v_out=vout)
    q, pv = port.stencil7(u, *dims, 0.5, 0.125, want_v=True)
    assert (r.min_val, r.max_val, r.min_loc, r.max_loc) == (q.min_val, q.max_val, q.min_loc, q.max_loc)
    if store:
        assert np.array_equal(vout.to_host(), pv)


@pytest.mark.parametrize("dims", ((64, 20, 18), (33, 9, 9)))
def test_stencil_ties_keep_the_lowest_location(space, port, dims):
    """Equal extrema: the first location in the reference's host iteration order (i slowest, k fastest) wins, i.e. the
    lowest flattened location -- a constant field (every interior point ties) and a field with duplicated extrema."""
    n = dims[0] * dims[1] * dims[2]
    for u in (np.full(n, 1.5), np.tile(np.array([1.0, -2.0, 3.0, 0.5, 3.0, -2.0]), n // 6 + 1)[:n].copy()):
        r = space.stencil7_minmaxloc(space.view_from_host(u), *dims, 0.5, 0.125)
        q, _ = port.stencil7(u, *dims, 0.5, 0.125)
        assert (r.min_val, r.max_val, r.min_loc, r.max_loc) == (q.min_val, q.max_val, q.min_loc, q.max_loc)


def test_stencil_unaligned_view_falls_back(space, port):
    dims = (64, 12, 10)
    n = dims[0] * dims[1] * dims[2]
    u = W.c1_uniform(n + 1)
    base = space.view_from_host(u)
    r = space.stencil7_minmaxloc(base.subview(1, n + 1), *dims, 0.5, 0.125)   # 8-byte aligned only
    q, _ = port.stencil7(u[1:], *dims, 0.5, 0.125)
    assert (r.min_val, r.max_val, r.min_loc, r.max_loc) == (q.min_val, q.max_val, q.min_loc, q.max_loc)


def test_stencil_empty_interior_yields_identity(space):
    u = space.view_from_host(np.ones(2 * 5 * 5))
    r = space.stencil7_minmaxloc(u, 2, 5, 5, 1.0, 1.0)
    assert (r.min_val, r.max_val, r.min_loc, r.max_loc) == (F64_MAX, -F64_MAX, I64_MAX, I64_MAX)


# ------------------------------------------------------------------ atomics
def test_gups_final_table_bit_exact(space, port):
    tl, m = 1 << 12, 1 << 15  # heavy collisions: 8 updates per entry on average
    idx = W.c5_indices(m, tl)
    for op, d in (("add", 7), ("xor", -1)):
        t = np.full(tl, 10101010101, dtype=np.int64)
        vt = space.view_from_host(t)
        space.gups(vt, space.view_from_host(idx), d, op)
        space.fence()
        port.gups(t, idx, d, op)
        got = vt.to_host()
        assert np.array_equal(got, t) and W.checksum64(got) == C[f"gups/{op}"]
    # all updates on ONE entry: worst-case contention
    idx0 = np.zeros(1 << 16, dtype=np.int64)
    vt = space.view_from_host(np.zeros(4, dtype=np.int64))
    space.gups(vt, space.view_from_host(idx0), 3, "add")
    assert vt.to_host()[0] == 3 * (1 << 16)


def test_atomic_add_f64_integer_valued_exact(space, port):
    tl, m = 1000, 1 << 16
    idx = W.c5_indices(m, tl, 5)
    vals = (W.hash_u32(np.arange(m, dtype=np.uint64)) % np.uint64(9)).astype(np.float64) - 4.0
    t = np.zeros(tl)
    port.atomic_add_f64(t, idx, vals)
    vt = space.view_from_host(np.zeros(tl))
    space.atomic_add_f64(vt, space.view_from_host(idx), space.view_from_host(vals))
    space.fence()
    assert np.array_equal(vt.to_host(), t)


# ------------------------------------------------------------------ TeamPolicy SpMV
@pytest.mark.parametrize("nnz_per_row", (3, 8, 32, 50))
def test_spmv_crs(space, port, nnz_per_row):
    nrows = 1000
    for iv in (True, False):
        rm, ci, va, x = W.c5_crs(nrows, nnz_per_row, integer_valued=iv)
        y = space.view(nrows, np.float64)
        space.spmv_crs(space.view_from_host(rm), space.view_from_host(ci), space.view_from_host(va), space.view_from_host(x), y)
        space.fence()
        py = port.spmv(rm, ci, va, x)
        if iv:  # integer-valued: every association order gives the same bits
            assert np.array_equal(y.to_host(), py)
            if nnz_per_row == 32:
                assert W.checksum64(y.to_host()) == C["spmv/int"]["checksum"]
        else:   # rows reassociate: <= 1e-12 relative to the row's magnitude
            mag = np.array([np.sum(np.abs(va[rm[r]:rm[r + 1]] * x[ci[rm[r]:rm[r + 1]]])) for r in range(nrows)])
            assert np.max(np.abs(y.to_host() - py) / np.maximum(mag, 1e-300)) <= REL_TOL


def test_spmv_ragged_rows_and_empty_rows(space, port):
    rng = np.random.default_rng(5)
    lens = rng.integers(0, 70, 500)
    lens[::7] = 0
    rm = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    nnz = int(rm[-1])
    ci = rng.integers(0, 300, nnz).astype(np.int32)
    va = rng.integers(-5, 6, nnz).astype(np.float64)
    x = rng.integers(-3, 4, 300).astype(np.float64)
    y = space.view(500, np.float64)
    space.spmv_crs(space.view_from_host(rm), space.view_from_host(ci), space.view_from_host(va), space.view_from_host(x), y)
    space.fence()
    assert np.array_equal(y.to_host(), port.spmv(rm, ci, va, x))


# ------------------------------------------------------------------ live reference, where it travelled
def test_against_live_reference(space, ref):
    n = 1 << 21
    x = W.c1_exact(n)
    assert space.parallel_reduce_sum(space.view_from_host(x)) == ref.reduce("sum", x)
    xi = W.c3_wrap(n)
    vy = space.view(n, np.int64)
    total = space.parallel_scan(space.view_from_host(xi), vy, seed=3)
    ry, rt = ref.scan(xi, False, 3)
    assert total == rt and np.array_equal(vy.to_host(), ry)
    u, _, _ = W.c4_field(48, 40, 36)
    r = space.stencil7_minmaxloc(space.view_from_host(u), 48, 40, 36, 0.5, 0.125)
    q, _ = ref.stencil7(u, 48, 40, 36, 0.5, 0.125)
    assert (r.min_val, r.max_val, r.min_loc, r.max_loc) == (q.min_val, q.max_val, q.min_loc, q.max_loc)


# ------------------------------------------------------------------ BASELINE.json full sizes, by property
def test_full_size_c1_reduce_2_27(space):
    n = 1 << 27
    x = W.c1_exact(n)
    assert space.parallel_reduce_sum(space.view_from_host(x)) == float(x.astype(np.int64).sum())


def test_full_size_c3_scan_2_30_properties(space):
    """N = 2^30 int64: generated on the device by a scan of ones (closed form y[i] = i), then the C3 input
    is produced per 2^26 chunk on the host; properties: total == sum, y[0] == seed, adjacent difference
    reproduces the input on sampled windows, last element closed form."""
    n = 1 << 30
    x = space.view(n, np.int64)
    y = space.view(n, np.int64)
    chunk = 1 << 26
    host_total = 0
    for c in range(n // chunk):
        h = (W.hash_u32(np.arange(c * chunk, (c + 1) * chunk, dtype=np.uint64)) % np.uint64(7)).astype(np.int64) - 3
        host_total += int(h.sum())
        x.subview(c * chunk, (c + 1) * chunk).from_host(h)
    total = space.parallel_scan(x, y, seed=11)
    assert total == host_total
    for start in (0, 4607, (1 << 29) - 5, n - 4096):
        w = min(4096, n - start)
        ys = y.subview(start, start + w).to_host()
        xs = x.subview(start, start + w).to_host()
        assert np.array_equal(np.diff(ys), xs[:-1])
    assert y.subview(0, 1).to_host()[0] == 11
    last_x = x.subview(n - 1, n).to_host()[0]
    assert y.subview(n - 1, n).to_host()[0] + last_x == 11 + host_total
    # checksum of checksums: an inclusive scan of y's adjacent differences is the identity on x (sampled by sum)
    assert space.parallel_reduce_sum(x) == host_total


def test_full_size_c2_stream_2_28(space):
    n = 1 << 28
    a = space.view(n, np.float64); b = space.view(n, np.float64); c = space.view(n, np.float64)
    space.stream_set(a, 1.0); space.stream_set(b, 2.0); space.stream_set(c, 0.0)
    ga, gb, gc = 1.0, 2.0, 0.0
    for _ in range(3):
        space.stream_copy(a, c); gc = ga
        space.stream_scale(b, c, 3.0); gb = 3.0 * gc
        space.stream_add(a, b, c); gc = ga + gb
        space.stream_triad(a, b, c, 3.0); ga = gb + 3.0 * gc
    for v, g in ((a, ga), (b, gb), (c, gc)):
        mm = space.parallel_reduce_minmax(v)
        assert mm.min_val == g and mm.max_val == g  # every element equals the analytic value


def test_full_size_c4_stencil_512(space):
    dims = (512, 512, 512)
    u, pmax, pmin = W.c4_field(*dims)
    r = space.stencil7_minmaxloc(space.view_from_host(u), *dims, 0.5, 0.125)
    # the planted extrema dominate: the stencil max/min sit at the planted points (c0 > 6*c1*|neighbour|)
    assert r.max_loc == (pmax[0] * 512 + pmax[1]) * 512 + pmax[2]
    assert r.min_loc == (pmin[0] * 512 + pmin[1]) * 512 + pmin[2]


def test_full_size_c5_gups_2_30_table(space):
    tl, m = 1 << 30, 1 << 24
    table = space.view(tl, np.int64)
    space.lib.b200_memset_async(space.handle, table.ptr, 0, table.nbytes)
    idx = W.c5_indices(m, tl)
    vi = space.view_from_host(idx)
    space.gups(table, vi, 1, "add")
    space.fence()
    assert space.parallel_reduce_sum(table) == m          # every update landed exactly once
    space.gups(table, vi, -1, "add")
    space.fence()
    assert space.parallel_reduce_max(table) == 0 and space.parallel_reduce_min(table) == 0  # add/sub round trip
