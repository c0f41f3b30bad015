"""oracle/bindings.py -- TEST INFRASTRUCTURE ONLY (checker, never the product).

ctypes access to the two CPU checkers built by oracle/Makefile:
  * ``Port``  -> oracle/liboracle_port.so  (plain-C restatement, omp_oracle.c)
  * ``Ref``   -> oracle/_ref/libkokkos_ref_omp.so  (the unmodified reference, Kokkos::OpenMP)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, byref, c_double, c_float, c_int, c_int32, c_int64, c_void_p, c_char_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_PATH = os.path.join(HERE, "liboracle_port.so")
REF_PATH = os.path.join(HERE, "_ref", "libkokkos_ref_omp.so")


class ValLoc(Structure):
    _fields_ = [("val", c_double), ("loc", c_int64)]


class MinMaxLoc(Structure):
    _fields_ = [("min_val", c_double), ("max_val", c_double), ("min_loc", c_int64), ("max_loc", c_int64)]


class MinMax(Structure):
    _fields_ = [("min_val", c_double), ("max_val", c_double)]


_CT = {np.dtype(np.float64): ("f64", c_double), np.dtype(np.float32): ("f32", c_float),
       np.dtype(np.int64): ("i64", c_int64), np.dtype(np.int32): ("i32", c_int32)}


def build(ref: bool = True) -> None:
    """(Re)build the checkers with oracle/Makefile.  `ref` is only possible where /root/reference exists."""
    subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)
    if ref and os.path.exists("/root/reference/core/src/Kokkos_Core.hpp"):
        subprocess.run(["make", "-s", "-j8", "-C", HERE, "ref"], check=True)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(c_void_p)


class Port:
    """The plain-C restatement; `threads` selects which OpenMP thread count's association order to reproduce."""

    def __init__(self):
        if not os.path.exists(PORT_PATH):
            build(ref=False)
        L = ctypes.CDLL(PORT_PATH)
        self.L = L
        for sfx, ct in (("f64", c_double), ("f32", c_float), ("i64", c_int64), ("i32", c_int32)):
            f = getattr(L, f"oracle_reduce_sum_{sfx}")
            f.argtypes, f.restype = [c_void_p, c_int64, c_int], ct
        for op in ("min", "max"):
            for sfx, ct in (("f64", c_double), ("i64", c_int64), ("i32", c_int32)):
                f = getattr(L, f"oracle_reduce_{op}_{sfx}")
                f.argtypes, f.restype = [c_void_p, c_int64, c_int], ct
        L.oracle_reduce_minmax_f64.argtypes, L.oracle_reduce_minmax_f64.restype = [c_void_p, c_int64, c_int], MinMax
        for n, rt in (("minloc", ValLoc), ("maxloc", ValLoc), ("minmaxloc", MinMaxLoc)):
            f = getattr(L, f"oracle_reduce_{n}_f64")
            f.argtypes, f.restype = [c_void_p, c_int64, c_int64, c_int], rt
        for sfx, ct in (("f64", c_double), ("i64", c_int64), ("i32", c_int32)):
            f = getattr(L, f"oracle_scan_{sfx}")
            f.argtypes, f.restype = [c_void_p, c_void_p, c_int64, ct, c_int, c_int], ct
        L.oracle_stream_set_f64.argtypes = [c_void_p, c_double, c_int64]
        L.oracle_stream_copy_f64.argtypes = [c_void_p, c_void_p, c_int64]
        L.oracle_stream_scale_f64.argtypes = [c_void_p, c_void_p, c_double, c_int64]
        L.oracle_stream_add_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_int64]
        L.oracle_stream_triad_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_double, c_int64]
        L.oracle_stencil7_minmaxloc_f64.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_double, c_double]
        L.oracle_stencil7_minmaxloc_f64.restype = MinMaxLoc
        L.oracle_gups_add_i64.argtypes = [c_void_p, c_void_p, c_int64, c_int64]
        L.oracle_gups_xor_i64.argtypes = [c_void_p, c_void_p, c_int64, c_int64]
        L.oracle_atomic_add_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_int64]
        L.oracle_spmv_crs_f64.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        for f in (L.oracle_reduce_partition, L.oracle_scan_partition):
            f.argtypes = [c_int64, c_int, c_int, POINTER(c_int64), POINTER(c_int64)]
        L.oracle_auto_chunk.argtypes, L.oracle_auto_chunk.restype = [c_int64, c_int], c_int64

    def reduce(self, op: str, x: np.ndarray, threads: int = 8):
        sfx, _ = _CT[x.dtype]
        return getattr(self.L, f"oracle_reduce_{op}_{sfx}")(_ptr(x), x.size, threads)

    def reduce_minmax(self, x, threads=8):
        return self.L.oracle_reduce_minmax_f64(_ptr(x), x.size, threads)

    def reduce_loc(self, kind: str, x, base=0, threads=8):
        return getattr(self.L, f"oracle_reduce_{kind}_f64")(_ptr(x), x.size, base, threads)

    def scan(self, x: np.ndarray, inclusive=False, seed=0, threads=8):
        sfx, ct = _CT[x.dtype]
        y = np.empty_like(x)
        total = getattr(self.L, f"oracle_scan_{sfx}")(_ptr(x), _ptr(y), x.size, ct(seed), int(inclusive), threads)
        return y, total

    def stencil7(self, u: np.ndarray, n0, n1, n2, c0, c1, want_v=False):
        v = np.zeros_like(u) if want_v else None
        r = self.L.oracle_stencil7_minmaxloc_f64(_ptr(u), _ptr(v) if want_v else None, n0, n1, n2, c0, c1)
        return r, v

    def gups(self, table: np.ndarray, idx: np.ndarray, datum: int, op="add"):
        getattr(self.L, f"oracle_gups_{op}_i64")(_ptr(table), _ptr(idx), idx.size, datum)

    def atomic_add_f64(self, table, idx, vals):
        self.L.oracle_atomic_add_f64(_ptr(table), _ptr(idx), _ptr(vals), idx.size)

    def spmv(self, row_map, col_idx, values, x):
        y = np.empty(row_map.size - 1, dtype=np.float64)
        self.L.oracle_spmv_crs_f64(y.size, _ptr(row_map), _ptr(col_idx), _ptr(values), _ptr(x), _ptr(y))
        return y

    def stream(self, name: str, *args):
        getattr(self.L, f"oracle_stream_{name}_f64")(*args)

    def partition(self, kind: str, n: int, threads: int, rank: int):
        b, e = c_int64(), c_int64()
        getattr(self.L, f"oracle_{kind}_partition")(n, threads, rank, byref(b), byref(e))
        return b.value, e.value


def ref_available() -> bool:
    return os.path.exists(REF_PATH)


class Ref:
    """The unmodified reference (Kokkos::OpenMP).  One thread count per process (Kokkos::initialize)."""

    _instance = None

    def __new__(cls, threads: int = 0):
        if cls._instance is not None:
            return cls._instance
        if not ref_available():
            raise FileNotFoundError(f"{REF_PATH} not built (oracle/Makefile target `ref` needs /root/reference)")
        os.environ.setdefault("OMP_PROC_BIND", "spread")
        os.environ.setdefault("OMP_PLACES", "threads")
        self = super().__new__(cls)
        L = ctypes.CDLL(REF_PATH)
        self.L = L
        L.ref_init.argtypes, L.ref_init.restype = [c_int], c_int
        L.ref_concurrency.restype = c_int
        L.ref_version.restype = c_char_p
        for sfx, ct in (("f64", c_double), ("f32", c_float), ("i64", c_int64), ("i32", c_int32)):
            f = getattr(L, f"ref_reduce_sum_{sfx}")
            f.argtypes, f.restype = [c_void_p, c_int64], ct
        for op in ("min", "max"):
            for sfx, ct in (("f64", c_double), ("i64", c_int64), ("i32", c_int32)):
                f = getattr(L, f"ref_reduce_{op}_{sfx}")
                f.argtypes, f.restype = [c_void_p, c_int64], ct
        L.ref_reduce_minmax_f64.argtypes, L.ref_reduce_minmax_f64.restype = [c_void_p, c_int64], MinMax
        for n, rt in (("minloc", ValLoc), ("maxloc", ValLoc), ("minmaxloc", MinMaxLoc)):
            f = getattr(L, f"ref_reduce_{n}_f64")
            f.argtypes, f.restype = [c_void_p, c_int64, c_int64], rt
        for sfx, ct in (("f64", c_double), ("i64", c_int64), ("i32", c_int32)):
            f = getattr(L, f"ref_scan_{sfx}")
            f.argtypes, f.restype = [c_void_p, c_void_p, c_int64, ct, c_int], ct
        L.ref_stream_set_f64.argtypes = [c_void_p, c_double, c_int64]
        L.ref_stream_copy_f64.argtypes = [c_void_p, c_void_p, c_int64]
        L.ref_stream_scale_f64.argtypes = [c_void_p, c_void_p, c_double, c_int64]
        L.ref_stream_add_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_int64]
        L.ref_stream_triad_f64.argtypes = [c_void_p, c_void_p, c_void_p, c_double, c_int64]
        L.ref_stencil7_minmaxloc_f64.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_double, c_double]
        L.ref_stencil7_minmaxloc_f64.restype = MinMaxLoc
        L.ref_gups_add_i64.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int64]
        L.ref_gups_xor_i64.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int64]
        L.ref_spmv_crs_f64.argtypes = [c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p]
        L.ref_time_reduce_sum_f64.argtypes, L.ref_time_reduce_sum_f64.restype = [c_int64, c_int, POINTER(c_double)], c_double
        L.ref_time_scan_excl_i64.argtypes, L.ref_time_scan_excl_i64.restype = [c_int64, c_int, POINTER(c_int64)], c_double
        L.ref_time_stream_triad_f64.argtypes, L.ref_time_stream_triad_f64.restype = [c_int64, c_int], c_double
        self.threads = L.ref_init(threads)
        cls._instance = self
        return self

    def reduce(self, op: str, x: np.ndarray):
        sfx, _ = _CT[x.dtype]
        return getattr(self.L, f"ref_reduce_{op}_{sfx}")(_ptr(x), x.size)

    def reduce_minmax(self, x):
        return self.L.ref_reduce_minmax_f64(_ptr(x), x.size)

    def reduce_loc(self, kind: str, x, base=0):
        return getattr(self.L, f"ref_reduce_{kind}_f64")(_ptr(x), x.size, base)

    def scan(self, x: np.ndarray, inclusive=False, seed=0):
        sfx, ct = _CT[x.dtype]
        y = np.empty_like(x)
        total = getattr(self.L, f"ref_scan_{sfx}")(_ptr(x), _ptr(y), x.size, ct(seed), int(inclusive))
        return y, total

    def stencil7(self, u, n0, n1, n2, c0, c1, want_v=False):
        v = np.zeros_like(u) if want_v else None
        r = self.L.ref_stencil7_minmaxloc_f64(_ptr(u), _ptr(v) if want_v else None, n0, n1, n2, c0, c1)
        return r, v

    def gups(self, table, idx, datum, op="add"):
        getattr(self.L, f"ref_gups_{op}_i64")(_ptr(table), table.size, _ptr(idx), idx.size, datum)

    def spmv(self, row_map, col_idx, values, x):
        y = np.empty(row_map.size - 1, dtype=np.float64)
        self.L.ref_spmv_crs_f64(y.size, _ptr(row_map), _ptr(col_idx), _ptr(values), values.size, _ptr(x), x.size, _ptr(y))
        return y

    def stream(self, name: str, *args):
        getattr(self.L, f"ref_stream_{name}_f64")(*args)

    def time_reduce_sum_f64(self, n: int, reps: int):
        r = c_double()
        return self.L.ref_time_reduce_sum_f64(n, reps, byref(r)), r.value

    def time_scan_excl_i64(self, n: int, reps: int):
        t = c_int64()
        return self.L.ref_time_scan_excl_i64(n, reps, byref(t)), t.value

    def time_stream_triad_f64(self, n: int, reps: int):
        return self.L.ref_time_stream_triad_f64(n, reps)
