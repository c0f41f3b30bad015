"""ctypes binding of benchlib/libkokkos_arms.so: the headline workloads as Kokkos user code (KOKKOS_LAMBDA functors on the
UNMODIFIED reference headers) dispatched to Kokkos::B200 (this repository, through kokkos_b200/adapter), Kokkos::Cuda (the
reference's own backend, comparator) or CUB.  All calls are asynchronous on the stream given to `Arms`."""
import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libkokkos_arms.so")
B200, CUDA, CUB = 0, 1, 2
ARM_NAMES = {B200: "Kokkos::B200", CUDA: "Kokkos::Cuda", CUB: "CUB"}


def available() -> bool:
    return os.path.exists(LIB)


class ArmsError(RuntimeError):
    pass


class Arms:
    def __init__(self, device: int, stream: int):
        # libkokkos_b200.so first, by absolute path, so the arms library binds to the in-tree copy
        ctypes.CDLL(os.path.join(os.path.dirname(HERE), "kokkos_b200", "libkokkos_b200.so"), mode=ctypes.RTLD_GLOBAL)
        self.lib = ctypes.CDLL(LIB)
        L = self.lib
        L.kka_last_error.restype = c_char_p
        L.kka_init.argtypes = [c_int, c_void_p]
        L.kka_reduce_sum_f64.argtypes = [c_int, c_void_p, c_int64, c_void_p]
        L.kka_scan_excl_i64.argtypes = [c_int, c_void_p, c_void_p, c_int64, c_void_p]
        L.kka_std_exclusive_scan_i64.argtypes = [c_int, c_void_p, c_void_p, c_int64]
        L.kka_stream_copy_f64.argtypes = [c_int, c_void_p, c_void_p, c_int64]
        L.kka_stream_triad_f64.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_double, c_int64]
        L.kka_stencil7_minmaxloc_f64.argtypes = [c_int, c_void_p, c_int64, c_int64, c_int64, c_double, c_double, c_void_p]
        L.kka_stencil7_minmaxloc_f64_tiled.argtypes = [c_int, c_void_p, c_int64, c_int64, c_int64, c_double, c_double, c_void_p, c_int64, c_int64, c_int64]
        L.kka_launch_latency_us.argtypes = [c_int, c_int, c_int64, c_int, c_void_p, ctypes.POINTER(c_double)]
        L.kka_gups_add_i64.argtypes = [c_int, c_void_p, c_int64, c_void_p, c_int64, c_int64]
        L.kka_spmv_crs_f64.argtypes = [c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64]
        self._ok(L.kka_init(device, stream))
        self.live = True

    def _ok(self, rc):
        if rc != 0:
            raise ArmsError(f"kokkos_arms rc={rc}: {self.lib.kka_last_error().decode(errors='replace')}")

    def reduce_sum(self, arm, x_ptr, n, result_dev_ptr):
        self._ok(self.lib.kka_reduce_sum_f64(arm, x_ptr, n, result_dev_ptr))

    def scan_excl(self, arm, x_ptr, y_ptr, n, total_dev_ptr):
        self._ok(self.lib.kka_scan_excl_i64(arm, x_ptr, y_ptr, n, total_dev_ptr))

    def std_exclusive_scan(self, arm, x_ptr, y_ptr, n):
        self._ok(self.lib.kka_std_exclusive_scan_i64(arm, x_ptr, y_ptr, n))

    def stream_copy(self, arm, a_ptr, c_ptr, n):
        self._ok(self.lib.kka_stream_copy_f64(arm, a_ptr, c_ptr, n))

    def stream_triad(self, arm, a_ptr, b_ptr, c_ptr, scalar, n):
        self._ok(self.lib.kka_stream_triad_f64(arm, a_ptr, b_ptr, c_ptr, scalar, n))

    def stencil7_minmaxloc(self, arm, u_ptr, n0, n1, n2, c0, c1, result_dev_ptr):
        self._ok(self.lib.kka_stencil7_minmaxloc_f64(arm, u_ptr, n0, n1, n2, c0, c1, result_dev_ptr))

    def stencil7_minmaxloc_tiled(self, arm, u_ptr, n0, n1, n2, c0, c1, result_dev_ptr, tile):
        self._ok(self.lib.kka_stencil7_minmaxloc_f64_tiled(arm, u_ptr, n0, n1, n2, c0, c1, result_dev_ptr, *tile))

    def launch_latency_us(self, arm, op, n, batch, scratch_dev_ptr):
        out = c_double()
        self._ok(self.lib.kka_launch_latency_us(arm, op, n, batch, scratch_dev_ptr, ctypes.byref(out)))
        return out.value

    def gups_add(self, arm, table_ptr, table_len, idx_ptr, m, datum):
        self._ok(self.lib.kka_gups_add_i64(arm, table_ptr, table_len, idx_ptr, m, datum))

    def spmv(self, arm, nrows, row_map_ptr, col_ptr, val_ptr, x_ptr, y_ptr, nnz, ncols):
        self._ok(self.lib.kka_spmv_crs_f64(arm, nrows, row_map_ptr, col_ptr, val_ptr, x_ptr, y_ptr, nnz, ncols))

    def finalize(self):
        if self.live:
            self.live = False
            self._ok(self.lib.kka_finalize())
