"""tools/stencil_probe.py -- C4 stencil tuning probe on a B200 (gpurun): times b200_stencil7_minmaxloc_f64 at 512^3 for the
compiled-in TMA configurations and the row-per-warp fallback.  Not a reported bench value."""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if "--sweep" in sys.argv:
    os.environ.setdefault("KOKKOS_B200_LIB", os.path.join(ROOT, "kokkos_b200", "libkokkos_b200_sweep.so"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import kokkos_b200 as kb  # noqa: E402
from tools.configs_bench import time_it  # noqa: E402


def main():
    torch.cuda.set_device(0)
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    space = kb.B200(0, stream=side.cuda_stream)
    n0 = n1 = n2 = 512
    n = n0 * n1 * n2
    u = torch.rand(n, dtype=torch.float64, device="cuda")
    vout = torch.empty(n, dtype=torch.float64, device="cuda")
    vu, vv = space.wrap(u.data_ptr(), n, np.float64), space.wrap(vout.data_ptr(), n, np.float64)
    ref = None
    if "--dbg" in sys.argv:  # needs --sweep: the probe branches are compiled into the sweep library only
        for dbg in (0, 1, 2):
            kb.tune_set("stencil.dbg", dbg)
            b, m = time_it(lambda: space.stencil7_minmaxloc(vu, n0, n1, n2, 0.5, 0.125), side, 10)
            print(f"dbg={dbg} best {b:.3f} med {m:.3f} ms", flush=True)
        kb.tune_set("stencil.dbg", 0)
        return
    cts = (224, 256, 352, 384, 512) if "--sweep" in sys.argv else (224, 352, 512)
    nss = (5, 4, 3) if "--sweep" in sys.argv else (5,)
    for tma, ct, ns, kc in itertools.chain([(0, 0, 0, 0)], itertools.product((1,), cts, nss, (0, 32))):
        kb.tune_set("stencil.tma", tma)
        if tma:
            kb.tune_set("stencil.ct", ct); kb.tune_set("stencil.ns", ns); kb.tune_set("stencil.kc", kc)
        for store in (False, True):
            try:
                fn = lambda: space.stencil7_minmaxloc(vu, n0, n1, n2, 0.5, 0.125, v_out=vv if store else None)  # noqa: E731
                r = fn()
                got = (r.min_val, r.max_val, r.min_loc, r.max_loc)
                ref = ref or got
                b, m = time_it(fn, side, 10)
            except kb.B200Error as e:
                if e.code == -3:
                    continue
                raise
            nbytes = 8 * n + (8 * 510 ** 3 if store else 0)
            print(f"tma={tma} ct={ct:3d} ns={ns} kc={kc:3d} store={int(store)}  best {b:7.3f} ms  med {m:7.3f} ms  {nbytes / m / 1e6:8.1f} GB/s"
                  f"  {'OK' if got == ref else 'MISMATCH ' + str(got) + ' vs ' + str(ref)}", flush=True)


if __name__ == "__main__":
    main()
