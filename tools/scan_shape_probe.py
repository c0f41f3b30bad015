"""tools/scan_shape_probe.py -- tile shape x ring depth of the warp-specialised scan WITH the L2 prefetch, sweep build, one B200."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("KOKKOS_B200_LIB", os.path.join(ROOT, "kokkos_b200", "libkokkos_b200_sweep.so"))
import numpy as np, torch  # noqa: E402,E401
import kokkos_b200 as kb  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tools"))
from sweep import time_it  # noqa: E402

torch.cuda.set_device(0)
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
space = kb.B200(0, stream=side.cuda_stream)
n = 1 << 30
x = torch.randint(-3, 4, (n,), device="cuda", dtype=torch.int64); y = torch.empty_like(x)
vx, vy = space.wrap(x.data_ptr(), n, np.int64), space.wrap(y.data_ptr(), n, np.int64)
tot = torch.zeros(1, device="cuda", dtype=torch.int64)
fn = lambda: space.parallel_scan(vx, vy, total_dev=tot.data_ptr(), blocking=False)  # noqa: E731
ref = torch.cumsum(x[: 1 << 24], 0) - x[: 1 << 24]
kb.tune_set("scan.ws", 2); kb.tune_set("scan.lbw", 1)
shapes = [(128, 9, 4), (128, 9, 5), (128, 9, 6), (128, 7, 6), (128, 7, 8), (128, 5, 8), (128, 11, 4), (128, 13, 4), (256, 7, 4), (256, 9, 3), (256, 9, 4), (256, 5, 6), (512, 9, 3), (512, 5, 4)]
for (bl, nv, nb) in shapes:
    kb.tune_set("scan.block", bl); kb.tune_set("scan.nv", nv); kb.tune_set("scan.nbuf", nb)
    for pfd in (-1, 0):
        kb.tune_set("scan.pfd", pfd)
        try:
            fn(); torch.cuda.synchronize()
            ok = bool(torch.equal(y[: 1 << 24], ref))
            best, med = time_it(fn)
            print(f"block={bl:3d} nv={nv:2d} stages={nb}  pfd={'auto' if pfd < 0 else 'off ':4s} best {16*n/best/1e9:8.1f} med {16*n/med/1e9:8.1f} GB/s  {'ok' if ok else 'MISMATCH'}", flush=True)
        except kb.B200Error as e:
            print(f"block={bl} nv={nv} stages={nb}: {str(e)[:60]}", flush=True)
