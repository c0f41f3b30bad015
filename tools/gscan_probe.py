"""tools/gscan_probe.py -- generic (lambda) parallel_scan tile-shape probe on a B200; prints GB/s per variant."""
import ctypes
import os
import sys
from ctypes import POINTER, c_double, c_int, c_longlong

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kokkos_b200 as kb  # noqa: E402

kb.load_library()
cases = ctypes.CDLL(kb.CASES_LIB_PATH, mode=ctypes.RTLD_GLOBAL)
cases.kb200_perf_last_error.restype = ctypes.c_char_p
assert cases.kb200_case_init(0) == 0
cases.kb200_perf_scan_variant.argtypes = [c_int, c_longlong, c_int, c_int, POINTER(c_double), POINTER(c_longlong)]
names = ["default", "256x13", "256x17", "512x9", "512x13", "128x17", "128x9", "1024x9", "256x21", "256x25", "512x17", "512x21", "1024x13", "1024x17", "128x25"]
for log2n in (30,):
    n = 1 << log2n
    for v, nm in enumerate(names):
        out = (c_double * 2)()
        tot = c_longlong()
        rc = cases.kb200_perf_scan_variant(v, n, 3, 10, out, ctypes.byref(tot))
        if rc != 0:
            print(nm, "rc", rc, cases.kb200_perf_last_error())
            continue
        print(f"2^{log2n} {nm:8s} best {out[0]:7.3f} ms med {out[1]:7.3f} ms  {16 * n / out[1] / 1e6:8.1f} GB/s  total={tot.value}", flush=True)
cases.kb200_case_finalize()
