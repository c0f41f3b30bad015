"""kokkos_b200 -- Python binding (ctypes) of the B200-native Kokkos execution space.

The product is ``libkokkos_b200.so`` (hand-written sm_100a kernels behind the C ABI declared in
``include/kokkos_b200.h``) plus the header-only C++ layer under ``kokkos_b200/include/kb200``.
This module only loads the library and mirrors the reference's vocabulary for the hot path:

    ``B200``      the execution-space instance (Kokkos::Cuda, core/src/Cuda/Kokkos_Cuda.hpp:95-247)
    ``View``      a rank-1 device allocation (Kokkos::View<T*>, CudaSpace::allocate)
    ``parallel_reduce_* / parallel_scan_* / stream_* ...``  typed fast paths

There is no CPU fallback: if the CUDA library is missing or no B200 is visible, calls raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char, c_char_p, c_double, c_float, c_int, c_int32, c_int64,
                    c_size_t, c_uint32, c_uint64, c_void_p)

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KOKKOS_B200_LIB") or os.path.join(_HERE, "libkokkos_b200.so")  # override: tools/sweep.py only
CASES_LIB_PATH = os.path.join(_HERE, "libkokkos_b200_cases.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "kokkos_b200.h")


class B200Error(RuntimeError):
    """Raised for every non-zero return of the C ABI (std::runtime_error in the reference's terms)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[kokkos_b200 rc={code}] {message}")
        self.code = code


class ValLoc(Structure):
    _fields_ = [("val", c_double), ("loc", c_int64)]


class MinMaxLocVal(Structure):
    _fields_ = [("min_val", c_double), ("max_val", c_double), ("min_loc", c_int64), ("max_loc", c_int64)]


class MinMaxVal(Structure):
    _fields_ = [("min_val", c_double), ("max_val", c_double)]


class Props(Structure):
    _fields_ = [("device", c_int), ("cc_major", c_int), ("cc_minor", c_int), ("sm_count", c_int),
                ("max_threads_per_sm", c_int), ("warp_size", c_int), ("smem_per_block_optin", c_size_t),
                ("smem_per_sm", c_size_t), ("l2_bytes", c_size_t), ("total_mem", c_size_t), ("concurrency", c_int),
                ("name", c_char * 64)]


_lib = None


def load_library() -> ctypes.CDLL:
    """Load libkokkos_b200.so (built in-tree by ``__graft_entry__.build()``); fail loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(-2, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    P = c_void_p
    sig = {
        "b200_init": [c_int, POINTER(P)],
        "b200_instance_create": [c_int, P, POINTER(P)],
        "b200_finalize": [P],
        "b200_fence": [P, c_char_p],
        "b200_device_props": [P, POINTER(Props)],
        "b200_device_count": [POINTER(c_int)],
        "b200_malloc": [P, c_size_t, POINTER(P)],
        "b200_free": [P, P],
        "b200_malloc_host_pinned": [c_size_t, POINTER(P)],
        "b200_free_host_pinned": [P],
        "b200_memset_async": [P, P, c_int, c_size_t],
        "b200_memcpy_h2d_async": [P, P, P, c_size_t],
        "b200_memcpy_d2h_async": [P, P, P, c_size_t],
        "b200_memcpy_d2d_async": [P, P, P, c_size_t],
        "b200_tune_set": [c_char_p, c_int],
        "b200_tune_get": [c_char_p, POINTER(c_int)],
        "b200_stream_set_f64": [P, P, c_double, c_int64],
        "b200_stream_copy_f64": [P, P, P, c_int64],
        "b200_stream_scale_f64": [P, P, P, c_double, c_int64],
        "b200_stream_add_f64": [P, P, P, P, c_int64],
        "b200_stream_triad_f64": [P, P, P, P, c_double, c_int64],
        "b200_stencil7_minmaxloc_f64": [P, P, P, c_int64, c_int64, c_int64, c_double, c_double, P, P],
        "b200_gups_add_i64": [P, P, c_int64, P, c_int64, c_int64],
        "b200_gups_xor_i64": [P, P, c_int64, P, c_int64, c_int64],
        "b200_atomic_add_f64": [P, P, c_int64, P, P, c_int64],
        "b200_spmv_crs_f64": [P, c_int64, P, P, P, P, P],
        "b200_scan_excl_i64_seed_dev": [P, P, P, c_int64, P, P],
        "b200_scan_excl_i64_seeds_dev": [P, P, P, c_int64, P, c_int, P],
        "b200_result_slot": [P, c_size_t, POINTER(P), POINTER(P), POINTER(P), POINTER(c_uint64)],
        "b200_result_wait": [P, P, c_uint64, c_char_p],
        "b200_reduce_sum_f64_host": [P, P, c_int64, P],
        "b200_scan_excl_i64_host": [P, P, P, c_int64, c_int64, P],
        "b200_comm_unique_id": [ctypes.c_char_p, c_size_t],
        "b200_comm_init": [P, c_int, c_int, c_char_p, POINTER(P)],
        "b200_comm_finalize": [P],
        "b200_comm_barrier": [P],
        "b200_comm_host_barrier": [P],
        "b200_allgather_bytes": [P, P, P, c_size_t],
        "b200_allreduce_minloc_f64": [P, P],
        "b200_allreduce_maxloc_f64": [P, P],
        "b200_allreduce_minmaxloc_f64": [P, P],
        "b200_comm_cyclic_layout": [P, c_int, c_int64, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)],
        "b200_comm_scan_excl_i64": [P, P, P, c_int64, P, P],
        "b200_comm_scan_incl_i64": [P, P, P, c_int64, P, P],
        "b200_comm_scan_excl_f64": [P, P, P, c_int64, P, P],
    }
    for op in ("sum", "min", "max"):
        for sfx in ("f64", "i64"):
            sig[f"b200_allreduce_{op}_{sfx}"] = [P, P, c_int]
    for name in ("sum_f64", "sum_f32", "sum_i64", "sum_i32", "min_f64", "max_f64", "min_i64", "max_i64", "min_i32",
                 "max_i32", "minmax_f64"):
        sig[f"b200_reduce_{name}"] = [P, P, c_int64, P, P]
    for name in ("minloc_f64", "maxloc_f64", "minmaxloc_f64"):
        sig[f"b200_reduce_{name}"] = [P, P, c_int64, c_int64, P, P]
    for name, ct in (("excl_i64", c_int64), ("incl_i64", c_int64), ("excl_f64", c_double), ("incl_f64", c_double),
                     ("excl_i32", c_int32)):
        sig[f"b200_scan_{name}"] = [P, P, P, c_int64, ct, P, P]
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = c_int
    lib.b200_last_error_string.restype = c_char_p
    lib.b200_version.restype = c_char_p
    lib.b200_instance_stream.argtypes = [P]
    lib.b200_instance_stream.restype = P
    lib.b200_instance_id.argtypes = [P]
    lib.b200_instance_id.restype = c_uint32
    lib.b200_instance_sm_count.argtypes = [P]
    lib.b200_instance_sm_count.restype = c_int
    for name in ("b200_comm_rank", "b200_comm_world"):
        getattr(lib, name).argtypes = [P]
        getattr(lib, name).restype = c_int
    lib.b200_comm_error.argtypes = [P]
    lib.b200_comm_error.restype = c_uint32
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != 0:
        raise B200Error(rc, load_library().b200_last_error_string().decode(errors="replace"))


def tune_set(key: str, value: int) -> None:
    _check(load_library().b200_tune_set(key.encode(), int(value)))


_DTYPES = {"f64": np.float64, "f32": np.float32, "i64": np.int64, "i32": np.int32}
_SUFFIX = {np.dtype(v): k for k, v in _DTYPES.items()}
_CT = {"f64": c_double, "f32": c_float, "i64": c_int64, "i32": c_int32}


class View:
    """Rank-1 contiguous device allocation, the analogue of ``Kokkos::View<T*, B200Space>``."""

    def __init__(self, space: "B200", n: int, dtype, label: str = "", ptr: int | None = None, owner=None):
        self.space = space
        self.dtype = np.dtype(dtype)
        self.n = int(n)
        self.label = label
        self._owner = owner
        self._owns = ptr is None
        if ptr is None:
            p = c_void_p()
            _check(space.lib.b200_malloc(space.handle, self.nbytes, byref(p)))
            self.ptr = p.value or 0
        else:
            self.ptr = int(ptr)

    @property
    def nbytes(self) -> int:
        return self.n * self.dtype.itemsize

    def extent(self, r: int = 0) -> int:
        return self.n

    def subview(self, begin: int, end: int) -> "View":
        if not (0 <= begin <= end <= self.n):
            raise B200Error(-1, "subview bounds out of range")
        return View(self.space, end - begin, self.dtype, self.label, self.ptr + begin * self.dtype.itemsize, owner=self)

    def from_host(self, arr: np.ndarray) -> "View":  # deep_copy(view, host)
        arr = np.ascontiguousarray(arr, dtype=self.dtype)
        if arr.size != self.n:
            raise B200Error(-1, "deep_copy extent mismatch")
        _check(self.space.lib.b200_memcpy_h2d_async(self.space.handle, self.ptr, arr.ctypes.data, self.nbytes))
        self.space.fence("deep_copy h2d")
        return self

    def to_host(self) -> np.ndarray:  # create_mirror_view_and_copy(HostSpace, view)
        out = np.empty(self.n, dtype=self.dtype)
        if self.n:
            _check(self.space.lib.b200_memcpy_d2h_async(self.space.handle, out.ctypes.data, self.ptr, self.nbytes))
            self.space.fence("deep_copy d2h")
        return out

    def free(self) -> None:
        if self._owns and self.ptr and self.space.handle:
            _check(self.space.lib.b200_free(self.space.handle, self.ptr))
        self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class B200:
    """Execution-space instance = device + stream + scratch (``Kokkos::Cuda`` in the reference)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load_library()
        h = c_void_p()
        if stream is None:
            _check(self.lib.b200_init(device, byref(h)))
        else:
            _check(self.lib.b200_instance_create(device, c_void_p(stream), byref(h)))
        self.handle = h
        self.device = device

    # ---- instance ----
    @staticmethod
    def name() -> str:
        return "B200"

    def fence(self, label: str = "B200::fence") -> None:
        _check(self.lib.b200_fence(self.handle, label.encode()))

    def props(self) -> Props:
        p = Props()
        _check(self.lib.b200_device_props(self.handle, byref(p)))
        return p

    def concurrency(self) -> int:
        return self.props().concurrency

    @property
    def stream(self) -> int:
        return self.lib.b200_instance_stream(self.handle) or 0

    def finalize(self) -> None:
        if self.handle:
            _check(self.lib.b200_finalize(self.handle))
            self.handle = None

    # ---- views ----
    def view(self, n: int, dtype, label: str = "") -> View:
        return View(self, n, dtype, label)

    def view_from_host(self, arr: np.ndarray, label: str = "") -> View:
        arr = np.ascontiguousarray(arr)
        return View(self, arr.size, arr.dtype, label).from_host(arr)

    def wrap(self, ptr: int, n: int, dtype) -> View:
        """Unmanaged View over caller-owned device memory (e.g. a torch tensor's data_ptr())."""
        return View(self, n, dtype, "unmanaged", ptr=ptr)

    # ---- parallel_reduce, typed fast paths ----
    def _reduce_scalar(self, op: str, v: View, result_dev: int = 0, blocking: bool = True):
        sfx = _SUFFIX[v.dtype]
        fn = getattr(self.lib, f"b200_reduce_{op}_{sfx}")
        out = _CT[sfx]()
        _check(fn(self.handle, v.ptr, v.n, byref(out) if blocking else None, result_dev or None))
        return out.value if blocking else None

    def parallel_reduce_sum(self, v: View, result_dev: int = 0, blocking: bool = True):
        return self._reduce_scalar("sum", v, result_dev, blocking)

    def parallel_reduce_min(self, v: View, result_dev: int = 0, blocking: bool = True):
        return self._reduce_scalar("min", v, result_dev, blocking)

    def parallel_reduce_max(self, v: View, result_dev: int = 0, blocking: bool = True):
        return self._reduce_scalar("max", v, result_dev, blocking)

    def parallel_reduce_minmax(self, v: View) -> MinMaxVal:
        out = MinMaxVal()
        _check(self.lib.b200_reduce_minmax_f64(self.handle, v.ptr, v.n, byref(out), None))
        return out

    def parallel_reduce_minloc(self, v: View, index_base: int = 0) -> ValLoc:
        out = ValLoc()
        _check(self.lib.b200_reduce_minloc_f64(self.handle, v.ptr, v.n, index_base, byref(out), None))
        return out

    def parallel_reduce_maxloc(self, v: View, index_base: int = 0) -> ValLoc:
        out = ValLoc()
        _check(self.lib.b200_reduce_maxloc_f64(self.handle, v.ptr, v.n, index_base, byref(out), None))
        return out

    def parallel_reduce_minmaxloc(self, v: View, index_base: int = 0) -> MinMaxLocVal:
        out = MinMaxLocVal()
        _check(self.lib.b200_reduce_minmaxloc_f64(self.handle, v.ptr, v.n, index_base, byref(out), None))
        return out

    def parallel_reduce_minmaxloc_dev(self, v: View, index_base: int, result_dev: int) -> None:
        """Asynchronous form: the 32-byte b200_minmaxloc_f64 goes to device memory at `result_dev`."""
        _check(self.lib.b200_reduce_minmaxloc_f64(self.handle, v.ptr, v.n, index_base, None, result_dev))

    # ---- parallel_scan ----
    def parallel_scan(self, x: View, y: View, inclusive: bool = False, seed=0, total_dev: int = 0, blocking: bool = True):
        if x.dtype != y.dtype or x.n != y.n:
            raise B200Error(-1, "parallel_scan: x and y must have the same type and extent")
        sfx = _SUFFIX[x.dtype]
        name = f"b200_scan_{'incl' if inclusive else 'excl'}_{sfx}"
        if not hasattr(self.lib, name):
            raise B200Error(-3, f"{name} is not provided")
        ct = _CT[sfx]
        out = ct()
        _check(getattr(self.lib, name)(self.handle, x.ptr, y.ptr, x.n, ct(seed), byref(out) if blocking else None,
                                       total_dev or None))
        return out.value if blocking else None

    def parallel_scan_seed_dev(self, x: View, y: View, seed_dev: int, total_dev: int = 0) -> None:
        _check(self.lib.b200_scan_excl_i64_seed_dev(self.handle, x.ptr, y.ptr, x.n, seed_dev, total_dev or None))

    def parallel_scan_seeds_dev(self, x: View, y: View, seeds_dev: int, nseeds: int, total_dev: int = 0) -> None:
        """Exclusive int64 scan seeded with sum(seeds_dev[0:nseeds]) read on the device (distributed scan, rank = nseeds)."""
        _check(self.lib.b200_scan_excl_i64_seeds_dev(self.handle, x.ptr, y.ptr, x.n, seeds_dev or None, nseeds, total_dev or None))

    # ---- host-buffer forms (chunked H2D / kernel / D2H pipeline inside the library) ----
    def parallel_reduce_sum_host(self, host_ptr: int, n: int) -> float:
        out = c_double()
        _check(self.lib.b200_reduce_sum_f64_host(self.handle, host_ptr, n, byref(out)))
        return out.value

    def parallel_scan_host(self, host_x_ptr: int, host_y_ptr: int, n: int, seed: int = 0) -> int:
        out = c_int64()
        _check(self.lib.b200_scan_excl_i64_host(self.handle, host_x_ptr, host_y_ptr, n, seed, byref(out)))
        return out.value

    # ---- parallel_for (stream) ----
    def stream_set(self, a: View, value: float) -> None:
        _check(self.lib.b200_stream_set_f64(self.handle, a.ptr, value, a.n))

    def stream_copy(self, a: View, b: View) -> None:
        _check(self.lib.b200_stream_copy_f64(self.handle, a.ptr, b.ptr, a.n))

    def stream_scale(self, b: View, c: View, s: float) -> None:
        _check(self.lib.b200_stream_scale_f64(self.handle, b.ptr, c.ptr, s, b.n))

    def stream_add(self, a: View, b: View, c: View) -> None:
        _check(self.lib.b200_stream_add_f64(self.handle, a.ptr, b.ptr, c.ptr, a.n))

    def stream_triad(self, a: View, b: View, c: View, s: float) -> None:
        _check(self.lib.b200_stream_triad_f64(self.handle, a.ptr, b.ptr, c.ptr, s, a.n))

    # ---- MDRange stencil ----
    def stencil7_minmaxloc(self, u: View, n0: int, n1: int, n2: int, c0: float, c1: float, v_out: View | None = None) -> MinMaxLocVal:
        if u.n != n0 * n1 * n2:
            raise B200Error(-1, "stencil7: extent mismatch")
        out = MinMaxLocVal()
        _check(self.lib.b200_stencil7_minmaxloc_f64(self.handle, u.ptr, v_out.ptr if v_out else None, n0, n1, n2, c0, c1,
                                                    byref(out), None))
        return out

    def stencil7_minmaxloc_dev(self, u: View, n0: int, n1: int, n2: int, c0: float, c1: float, result_dev: int, v_out: View | None = None) -> None:
        """Asynchronous form: the 32-byte b200_minmaxloc_f64 goes to device memory at `result_dev`."""
        if u.n != n0 * n1 * n2:
            raise B200Error(-1, "stencil7: extent mismatch")
        _check(self.lib.b200_stencil7_minmaxloc_f64(self.handle, u.ptr, v_out.ptr if v_out else None, n0, n1, n2, c0, c1, None, result_dev))

    # ---- atomics ----
    def gups(self, table: View, indices: View, datum: int, op: str = "add") -> None:
        fn = {"add": self.lib.b200_gups_add_i64, "xor": self.lib.b200_gups_xor_i64}[op]
        _check(fn(self.handle, table.ptr, table.n, indices.ptr, indices.n, datum))

    def atomic_add_f64(self, table: View, indices: View, values: View) -> None:
        _check(self.lib.b200_atomic_add_f64(self.handle, table.ptr, table.n, indices.ptr, values.ptr, indices.n))

    # ---- TeamPolicy SpMV ----
    def spmv_crs(self, row_map: View, col_idx: View, values: View, x: View, y: View) -> None:
        _check(self.lib.b200_spmv_crs_f64(self.handle, y.n, row_map.ptr, col_idx.ptr, values.ptr, x.ptr, y.ptr))

    def __del__(self):
        try:
            self.finalize()
        except Exception:
            pass


def comm_unique_id() -> str:
    """Rank 0 creates the id and hands it to the other ranks out of band (the role of ncclGetUniqueId)."""
    buf = ctypes.create_string_buffer(64)
    _check(load_library().b200_comm_unique_id(buf, 64))
    return buf.value.decode()


class Comm:
    """One-box communicator: one process per GPU, peer-mapped mailboxes over NVLink, the library's own kernels
    (``b200_comm_*`` / ``b200_allreduce_*`` / ``b200_comm_scan_*`` in include/kokkos_b200.h).  Every method is a
    collective: all ranks call it, in the same order, on their instance's stream."""

    def __init__(self, space: B200, rank: int, world: int, unique_id: str = ""):
        self.space = space
        self.lib = space.lib
        h = c_void_p()
        _check(self.lib.b200_comm_init(space.handle, rank, world, unique_id.encode(), byref(h)))
        self.handle = h
        self.rank, self.world = rank, world

    def barrier(self) -> None:
        _check(self.lib.b200_comm_barrier(self.handle))

    def host_barrier(self) -> None:
        _check(self.lib.b200_comm_host_barrier(self.handle))

    def error(self) -> int:
        return int(self.lib.b200_comm_error(self.handle))

    def allgather(self, src_dev: int, dst_dev: int, bytes_per_rank: int) -> None:
        _check(self.lib.b200_allgather_bytes(self.handle, src_dev, dst_dev, bytes_per_rank))

    def allreduce(self, op: str, buf_dev: int, count: int, dtype) -> None:
        """In-place fold (sum/min/max) of `count` f64 or i64 values per rank, in rank order."""
        _check(getattr(self.lib, f"b200_allreduce_{op}_{_SUFFIX[np.dtype(dtype)]}")(self.handle, buf_dev, count))

    def allreduce_loc(self, kind: str, buf_dev: int) -> None:
        """kind: minloc / maxloc (b200_valloc_f64) or minmaxloc (b200_minmaxloc_f64), in place."""
        _check(getattr(self.lib, f"b200_allreduce_{kind}_f64")(self.handle, buf_dev))

    def cyclic_layout(self, n_global: int, dtype) -> tuple[int, int, int]:
        """(block_elems, n_local, nsteps) of the block-cyclic distribution the fused scan works on."""
        b, nl, ns = c_int64(), c_int64(), c_int64()
        _check(self.lib.b200_comm_cyclic_layout(self.handle, np.dtype(dtype).itemsize, n_global, byref(b), byref(nl), byref(ns)))
        return b.value, nl.value, ns.value

    def parallel_scan(self, x: View, y: View, n_global: int, inclusive: bool = False, total_dev: int = 0, blocking: bool = True):
        """Distributed parallel_scan over a block-cyclic View (local blocks in x / y); returns the GLOBAL total if blocking."""
        sfx = _SUFFIX[x.dtype]
        name = f"b200_comm_scan_{'incl' if inclusive else 'excl'}_{sfx}"
        if not hasattr(self.lib, name):
            raise B200Error(-3, f"{name} is not provided")
        out = _CT[sfx]()
        _check(getattr(self.lib, name)(self.handle, x.ptr, y.ptr, n_global, byref(out) if blocking else None, total_dev or None))
        return out.value if blocking else None

    def finalize(self) -> None:
        if self.handle:
            _check(self.lib.b200_comm_finalize(self.handle))
            self.handle = None


def version() -> str:
    return load_library().b200_version().decode()


def device_count() -> int:
    n = c_int()
    _check(load_library().b200_device_count(byref(n)))
    return n.value
