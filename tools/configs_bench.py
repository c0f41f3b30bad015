"""tools/configs_bench.py -- every BASELINE.json config (C1..C5) on one B200: the typed C-ABI fast path and
the generic Kokkos-style lambda path (tests/cxx/cases_perf.cu) side by side, CUDA-event timed (inputs larger than L2,
3 warm-ups, best + median of `reps`), with the algorithmic-byte roofline fraction against MEASURED_PEAKS.json.
Writes gpurun_out/configs_bench.json; a copy is committed under profiles/ per round.  Not the bench.py headline."""
import ctypes
import json
import os
import sys
from ctypes import POINTER, c_double, c_int, c_int64, c_void_p, c_longlong

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import kokkos_b200 as kb  # noqa: E402


def time_it(fn, stream, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    only = set(sys.argv[1].split(",")) if len(sys.argv) > 1 and sys.argv[1] != "all" else None
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    space = kb.B200(0, stream=side.cuda_stream)
    cases = ctypes.CDLL(kb.CASES_LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    cases.kb200_perf_last_error.restype = ctypes.c_char_p
    assert cases.kb200_case_init(0) == 0
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    rows = []
    log2ns = tuple(int(x) for x in os.environ.get("KB200_LOG2N", "27,30").split(","))  # ncu runs use 27 only

    def want(tag):
        return only is None or tag in only

    def report(cfg, path, what, nbytes, best_ms, med_ms, extra=None):
        row = {"config": cfg, "path": path, "what": what, "bytes": nbytes, "best_ms": best_ms, "med_ms": med_ms,
               "best_GBs": nbytes / best_ms / 1e6, "med_GBs": nbytes / med_ms / 1e6}
        row["frac_of_measured_peak"] = row["med_GBs"] / peak
        if extra:
            row.update(extra)
        rows.append(row)
        print(f"{cfg:4s} {path:7s} {what:44s} {med_ms:9.3f} ms  {row['med_GBs']:8.1f} GB/s  ({row['frac_of_measured_peak']*100:5.1f}% of measured {peak:.0f})"
              + (f"  {extra}" if extra else ""), flush=True)

    def perf(fn, *args):
        out = (c_double * 2)()
        rc = fn(*args, out)
        if rc != 0:
            raise RuntimeError(f"{fn.__name__}: rc={rc} {cases.kb200_perf_last_error().decode()}")
        return out[0], out[1]

    def perf_chk(fn, nchk, *args):
        out = (c_double * 2)()
        chk = (c_double * nchk)()
        rc = fn(*args, out, chk)
        if rc != 0:
            raise RuntimeError(f"{fn.__name__}: rc={rc} {cases.kb200_perf_last_error().decode()}")
        return out[0], out[1], list(chk)

    # ------------------------------------------------------------------ C1 reduce
    if want("c1"):
        for log2n in log2ns:
            n = 1 << log2n
            x = torch.empty(n, dtype=torch.float64, device=dev)
            CH = 1 << 26
            for c in range(0, n, CH):
                idx = torch.arange(c, min(n, c + CH), dtype=torch.int64, device=dev)
                x[c:c + idx.numel()] = (((idx * 2654435761) >> 7) % 100).double()
                del idx
            v = space.wrap(x.data_ptr(), n, np.float64)
            rd = torch.zeros(4, dtype=torch.float64, device=dev)
            b, m = time_it(lambda: space.parallel_reduce_sum(v, result_dev=rd.data_ptr(), blocking=False), side, reps)
            report("C1", "typed", f"reduce Sum<double> 2^{log2n} (device result)", 8 * n, b, m)
            b, m = time_it(lambda: space.parallel_reduce_sum(v), side, reps)
            report("C1", "typed", f"reduce Sum<double> 2^{log2n} (scalar, fenced)", 8 * n, b, m)
            b, m = time_it(lambda: space.parallel_reduce_minmaxloc(v), side, reps)
            report("C1", "typed", f"reduce MinMaxLoc<double,int64> 2^{log2n}", 8 * n, b, m)
            del x, v
            torch.cuda.empty_cache()
            for op, nm in ((0, "Sum scalar"), (1, "MinMaxLoc"), (2, "Sum -> View")):
                cases.kb200_perf_reduce.argtypes = [c_int, c_longlong, c_int, c_int, POINTER(c_double), POINTER(c_double)]
                b, m, chk = perf_chk(cases.kb200_perf_reduce, 4, op, n, 3, reps)
                report("C1", "lambda", f"parallel_reduce lambda {nm} 2^{log2n}", 8 * n, b, m)

    # ------------------------------------------------------------------ C2 stream
    if want("c2"):
        n = 1 << 28
        a = torch.full((n,), 1.0, dtype=torch.float64, device=dev)
        bb = torch.full((n,), 2.0, dtype=torch.float64, device=dev)
        cc = torch.full((n,), 0.5, dtype=torch.float64, device=dev)
        va, vb, vc = (space.wrap(t.data_ptr(), n, np.float64) for t in (a, bb, cc))
        b, m = time_it(lambda: space.stream_copy(va, vc), side, reps)
        report("C2", "typed", "stream copy 2^28", 16 * n, b, m)
        b, m = time_it(lambda: space.stream_triad(va, vb, vc, 3.0), side, reps)
        report("C2", "typed", "stream triad 2^28", 24 * n, b, m)
        b, m = time_it(lambda: cc.copy_(a), side, reps)
        report("C2", "torch", "torch copy_ 2^28 (comparator)", 16 * n, b, m)
        del a, bb, cc
        torch.cuda.empty_cache()
        cases.kb200_perf_stream.argtypes = [c_int, c_longlong, c_int, c_int, POINTER(c_double)]
        b, m = perf(cases.kb200_perf_stream, 0, n, 3, reps)
        report("C2", "lambda", "parallel_for lambda copy 2^28", 16 * n, b, m)
        b, m = perf(cases.kb200_perf_stream, 1, n, 3, reps)
        report("C2", "lambda", "parallel_for lambda triad 2^28", 24 * n, b, m)

    # ------------------------------------------------------------------ C3 scan
    if want("c3"):
        for log2n in log2ns:
            n = 1 << log2n
            xi = torch.empty(n, dtype=torch.int64, device=dev)
            yi = torch.empty(n, dtype=torch.int64, device=dev)
            CH = 1 << 26
            for c in range(0, n, CH):
                idx = torch.arange(c, min(n, c + CH), dtype=torch.int64, device=dev)
                xi[c:c + idx.numel()] = (((idx * 2654435761) >> 7) % 7) - 3
                del idx
            vx, vy = space.wrap(xi.data_ptr(), n, np.int64), space.wrap(yi.data_ptr(), n, np.int64)
            td = torch.zeros(1, dtype=torch.int64, device=dev)
            b, m = time_it(lambda: space.parallel_scan(vx, vy, total_dev=td.data_ptr(), blocking=False), side, reps)
            report("C3", "typed", f"scan exclusive int64 2^{log2n}", 16 * n, b, m)
            del xi, yi
            torch.cuda.empty_cache()
            cases.kb200_perf_scan.argtypes = [c_longlong, c_int, c_int, POINTER(c_double), POINTER(c_longlong)]
            out = (c_double * 2)()
            tot = c_longlong()
            rc = cases.kb200_perf_scan(n, 3, reps, out, ctypes.byref(tot))
            assert rc == 0, cases.kb200_perf_last_error()
            report("C3", "lambda", f"parallel_scan lambda int64 2^{log2n}", 16 * n, out[0], out[1])

    # ------------------------------------------------------------------ C4 stencil
    if want("c4"):
        n0 = n1 = n2 = 512
        n = n0 * n1 * n2
        u = torch.rand(n, dtype=torch.float64, device=dev)
        vout = torch.empty(n, dtype=torch.float64, device=dev)
        vu, vv = space.wrap(u.data_ptr(), n, np.float64), space.wrap(vout.data_ptr(), n, np.float64)
        b, m = time_it(lambda: space.stencil7_minmaxloc(vu, n0, n1, n2, 0.5, 0.125), side, reps)
        report("C4", "typed", "stencil7+MinMaxLoc 512^3 (reduce only)", 8 * n, b, m)
        b, m = time_it(lambda: space.stencil7_minmaxloc(vu, n0, n1, n2, 0.5, 0.125, v_out=vv), side, reps)
        report("C4", "typed", "stencil7+MinMaxLoc 512^3 (+ store v)", 8 * n + 8 * 510 ** 3, b, m)
        del u, vout
        torch.cuda.empty_cache()
        cases.kb200_perf_mdrange_stencil.argtypes = [c_longlong] * 3 + [c_int] * 3 + [POINTER(c_double), POINTER(c_double)]
        b, m, chk = perf_chk(cases.kb200_perf_mdrange_stencil, 4, n0, n1, n2, 0, 3, reps)
        report("C4", "lambda", "MDRange<3> lambda stencil7+MinMaxLoc 512^3", 8 * n, b, m)
        b, m, chk = perf_chk(cases.kb200_perf_mdrange_stencil, 4, n0, n1, n2, 1, 3, reps)
        report("C4", "lambda", "MDRange<3> lambda stencil7 (+ store v)", 8 * n + 8 * 510 ** 3, b, m)

    # ------------------------------------------------------------------ C5a GUPS
    if want("c5a"):
        tl = 1 << 30
        table = torch.full((tl,), 10101010101, dtype=torch.int64, device=dev)
        vt = space.wrap(table.data_ptr(), tl, np.int64)
        for log2m in (26, 28):
            m_ = 1 << log2m
            g = torch.Generator(device=dev)
            g.manual_seed(20230913)
            idx = torch.randint(0, tl, (m_,), dtype=torch.int64, device=dev, generator=g)
            vi = space.wrap(idx.data_ptr(), m_, np.int64)
            for op in ("add", "xor"):
                b, m = time_it(lambda: space.gups(vt, vi, -1, op), side, reps)
                report("C5a", "typed", f"gups atomic_{op} table 2^30, M=2^{log2m}", 16 * m_, b, m,
                       {"GUPS": m_ / m / 1e6, "sector_model_GBs": 72 * m_ / m / 1e6})
            del idx
        del table
        torch.cuda.empty_cache()
        cases.kb200_perf_gups.argtypes = [c_int, c_longlong, c_longlong, c_int, c_int, POINTER(c_double)]
        for op, nm in ((0, "atomic_add"), (1, "atomic_fetch_xor")):
            b, m = perf(cases.kb200_perf_gups, op, tl, 1 << 26, 3, reps)
            report("C5a", "lambda", f"parallel_for lambda {nm} table 2^30, M=2^26", 16 * (1 << 26), b, m,
                   {"GUPS": (1 << 26) / m / 1e6, "sector_model_GBs": 72 * (1 << 26) / m / 1e6})

    # ------------------------------------------------------------------ C5b SpMV
    if want("c5b"):
        R, K = 1 << 22, 32
        nnz = R * K
        row_map = torch.arange(0, nnz + 1, K, dtype=torch.int64, device=dev)
        g = torch.Generator(device=dev)
        g.manual_seed(7)
        # banded + random columns: half of each row within +-64 of the diagonal, half anywhere
        rws = torch.arange(R, dtype=torch.int64, device=dev).repeat_interleave(K)
        band = (rws + torch.randint(-64, 65, (nnz,), device=dev, generator=g)).clamp_(0, R - 1)
        rnd = torch.randint(0, R, (nnz,), device=dev, generator=g)
        sel = (torch.arange(nnz, device=dev) % 2) == 0
        col = torch.where(sel, band, rnd).to(torch.int32)
        del rws, band, rnd, sel
        val = torch.rand(nnz, dtype=torch.float64, device=dev, generator=g)
        x = torch.rand(R, dtype=torch.float64, device=dev, generator=g)
        y = torch.empty(R, dtype=torch.float64, device=dev)
        nbytes = nnz * 12 + R * 16 + 8 * R
        V = lambda t, dt: space.wrap(t.data_ptr(), t.numel(), dt)  # noqa: E731
        b, m = time_it(lambda: space.spmv_crs(V(row_map, np.int64), V(col, np.int32), V(val, np.float64), V(x, np.float64), V(y, np.float64)), side, reps)
        report("C5b", "typed", "spmv_crs 2^22 rows x 32 nnz (banded+random)", nbytes, b, m)
        yref = y.clone()
        cases.kb200_perf_team_spmv.argtypes = [c_longlong, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                               POINTER(c_double)]
        for rpt, ts, vl in ((8, 8, 32), (16, 16, 16), (32, 0, 32), (4, 8, 32)):
            y.zero_()
            b, m = perf(cases.kb200_perf_team_spmv, R, row_map.data_ptr(), col.data_ptr(), val.data_ptr(), x.data_ptr(), y.data_ptr(), rpt, ts, vl, 3, reps)
            ok = bool(torch.allclose(y, yref, rtol=1e-12, atol=0))
            report("C5b", "lambda", f"TeamPolicy lambda spmv rows/team={rpt} team={ts or 'AUTO'} vec={vl}", nbytes, b, m, {"matches_typed": ok})

    # ------------------------------------------------------------------ launch latency
    if want("lat"):
        cases.kb200_perf_launch_latency.argtypes = [c_int, c_longlong, c_int, c_int, POINTER(c_double)]
        for op, nm in ((0, "parallel_for"), (1, "parallel_reduce -> scalar"), (2, "parallel_reduce -> View")):
            for n in (1, 1 << 10, 1 << 16):
                out = (c_double * 1)()
                rc = cases.kb200_perf_launch_latency(op, n, 1000, 3, out)
                assert rc == 0, cases.kb200_perf_last_error()
                rows.append({"config": "lat", "path": "lambda", "what": f"{nm} n={n}", "us_per_call": out[0]})
                print(f"lat  lambda  {nm:28s} n={n:6d}  {out[0]:8.2f} us/call", flush=True)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs_bench.json"), "w") as f:
        json.dump({"peak_GBs": peak, "rows": rows}, f, indent=1)
    cases.kb200_case_finalize()
    space.finalize()


if __name__ == "__main__":
    main()
