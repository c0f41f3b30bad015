"""tools/for_probe.py -- generic (lambda) parallel_for launch-shape probe on a B200: copy lambda over 2^28 doubles."""
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
cases.kb200_perf_for_variant.argtypes = [c_int, c_longlong, c_int, c_int, POINTER(c_double)]
names = ["256x4 (shipped)", "256x1", "256x2", "256x8", "128x1", "512x1", "1024x1", "256x4 persistent", "256x1 persistent", "512x2"]
n = 1 << 28
for v, nm in enumerate(names):
    out = (c_double * 2)()
    rc = cases.kb200_perf_for_variant(v, n, 3, 10, out)
    if rc != 0:
        print(nm, "rc", rc, cases.kb200_perf_last_error())
        continue
    print(f"{nm:18s} best {out[0]:7.3f} ms med {out[1]:7.3f} ms  {16 * n / out[1] / 1e6:8.1f} GB/s", flush=True)
cases.kb200_case_finalize()
