"""Generate tests/golden/golden_v1.json from the UNMODIFIED reference (oracle/_ref, Kokkos::OpenMP).

Run in the build container (needs /root/reference to have been compiled by oracle/Makefile):
    python tests/golden/make_golden.py
Inputs are regenerated from tests/workloads.py seeds, so only outputs are stored: scalars as C99 hex
floats / decimal ints, arrays as a 64-bit position-weighted checksum (workloads.checksum64) plus a few
sampled elements.
The thread count the reference ran with is recorded: double-precision sums depend on it
(SURVEY.md section 3.2); the port reproduces the order for that count.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import workloads as W  # noqa: E402
from oracle.bindings import Ref  # noqa: E402

THREADS = 4


cks = W.checksum64


def main():
    ref = Ref(THREADS)
    assert ref.threads == THREADS, ref.threads
    G = {"reference": ref.L.ref_version().decode(), "threads": THREADS, "cases": {}}
    C = G["cases"]
    for n in (0, 1, 2, 5, 33, 1000, 4097, 100003, 1 << 20):
        for gen in ("c1_exact", "c1_general", "c1_uniform"):
            x = getattr(W, gen)(n)
            C[f"sum_f64/{gen}/{n}"] = float(ref.reduce("sum", x)).hex()
        xi = W.c3_wrap(n)
        C[f"sum_i64/c3_wrap/{n}"] = int(ref.reduce("sum", xi))
        C[f"min_i64/c3_wrap/{n}"] = int(ref.reduce("min", xi))
        C[f"max_i64/c3_wrap/{n}"] = int(ref.reduce("max", xi))
        xu = W.c1_uniform(n)
        C[f"min_f64/c1_uniform/{n}"] = float(ref.reduce("min", xu)).hex()
        C[f"max_f64/c1_uniform/{n}"] = float(ref.reduce("max", xu)).hex()
        r = ref.reduce_loc("minmaxloc", xu, 0)
        C[f"minmaxloc_f64/c1_uniform/{n}"] = [float(r.min_val).hex(), float(r.max_val).hex(), int(r.min_loc), int(r.max_loc)]
        # ties: few distinct values => many equal extrema; OpenMP static schedule gives the lowest index
        xt = (W.hash_u32(np.arange(n, dtype=np.uint64)) % np.uint64(5)).astype(np.float64)
        r = ref.reduce_loc("minmaxloc", xt, 10)
        C[f"minmaxloc_f64/ties/{n}"] = [float(r.min_val).hex(), float(r.max_val).hex(), int(r.min_loc), int(r.max_loc)]
        for gen in ("c3_small", "c3_wrap"):
            x = getattr(W, gen)(n)
            for incl in (False, True):
                y, total = ref.scan(x, inclusive=incl, seed=5)
                C[f"scan_i64/{gen}/{'incl' if incl else 'excl'}/{n}"] = {
                    "total": int(total), "checksum": cks(y), "tail": [int(v) for v in y[-3:]]}
        y, total = ref.scan(W.c1_exact(n), inclusive=False, seed=0.0)
        C[f"scan_f64/c1_exact/excl/{n}"] = {"total": float(total).hex(), "checksum": cks(y)}
    # stream: analytic a/b/c after k iterations of the benchmark loop (stream-kokkos.cpp:79-130) on n elements
    n = 4099
    a = np.full(n, 1.0); b = np.full(n, 2.0); c = np.zeros(n)
    P = lambda v: v.ctypes.data
    for it in range(5):
        ref.stream("copy", P(a), P(c), n)        # c = a
        ref.stream("scale", P(b), P(c), 3.0, n)  # b = 3*c
        ref.stream("add", P(a), P(b), P(c), n)   # c = a+b
        ref.stream("triad", P(a), P(b), P(c), 3.0, n)  # a = b+3*c
    C["stream/5iters"] = [float(a[0]).hex(), float(b[0]).hex(), float(c[0]).hex(), cks(a), cks(b), cks(c)]
    # stencil
    for dims in ((8, 9, 10), (34, 20, 18), (64, 64, 64)):
        u, pmax, pmin = W.c4_field(*dims)
        r, v = ref.stencil7(u, *dims, 0.5, 0.125, want_v=True)
        C[f"stencil7/{dims[0]}x{dims[1]}x{dims[2]}"] = {
            "minmaxloc": [float(r.min_val).hex(), float(r.max_val).hex(), int(r.min_loc), int(r.max_loc)],
            "v_checksum": cks(v), "pmax": list(pmax), "pmin": list(pmin)}
    # gups
    tl, m = 1 << 12, 1 << 15
    idx = W.c5_indices(m, tl)
    for op, d in (("add", 7), ("xor", -1)):
        t = np.full(tl, 10101010101, dtype=np.int64)
        ref.gups(t, idx, d, op)
        C[f"gups/{op}"] = cks(t)
    # spmv (integer-valued => order independent, and general)
    for iv in (True, False):
        rm, ci, va, x = W.c5_crs(1000, 32, integer_valued=iv)
        y = ref.spmv(rm, ci, va, x)
        C[f"spmv/{'int' if iv else 'real'}"] = {"checksum": cks(y), "y0": float(y[0]).hex(), "y999": float(y[999]).hex()}
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.json")
    with open(out, "w") as f:
        json.dump(G, f, indent=0, sort_keys=True)
    print("wrote", out, len(C), "cases")


if __name__ == "__main__":
    main()
