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
names = ["default(1024x17 lbw2)", "1024x17 lbw1", "1024x13 lbw4", "512x13 lbw4", "512x17 lbw4", "512x13 lbw8", "1024x19 lbw2", "1024x21 lbw2", "256x13 lbw8", "512x21 lbw4", "768x13 lbw4", "1024x13 lbw8", "512x25 lbw4", "1024x21 lbw4"]
only = [int(x) for x in os.environ.get('KB200_VARIANTS', '').split(',') if x]
for log2n in (int(os.environ.get('KB200_LOG2N', '30')),):
    n = 1 << log2n
    for v, nm in enumerate(names):
        if only and v not in only:
            continue
        out = (c_double * 2)()
        tot = c_longlong()
        rc = cases.kb200_perf_scan_variant(v, n, 3, 10, out, ctypes.byref(tot))
        if rc != 0:
            print(nm, "rc", rc, cases.kb200_perf_last_error())
            continue
        print(f"2^{log2n} {nm:22s} best {out[0]:7.3f} ms med {out[1]:7.3f} ms  {16 * n / out[1] / 1e6:8.1f} GB/s  total={tot.value}", flush=True)
cases.kb200_case_finalize()
