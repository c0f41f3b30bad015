"""tools/mdtile_probe.py -- MDRangePolicy<Rank<3>> tile-shape probe for the generic (lambda) stencil + MinMaxLoc reduce on a B200:
the Kokkos user lambda of benchlib/kokkos_arms.cu on Kokkos::B200 (and Kokkos::Cuda for the default), 512^3 doubles."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from benchlib import arms as A  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
arms = A.Arms(0, side.cuda_stream)
n = 512
u = torch.rand(n * n * n, dtype=torch.float64, device=dev)
out = torch.zeros(4, dtype=torch.float64, device=dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side); fn(); e1.record(side); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


arms.stencil7_minmaxloc(A.B200, u.data_ptr(), n, n, n, 0.5, 0.125, out.data_ptr())
torch.cuda.synchronize()
ref = out.clone()
tiles = [(0, 0, 0), (32, 4, 4), (64, 2, 4), (64, 4, 2), (128, 2, 2), (128, 1, 4), (128, 4, 1), (256, 1, 2), (256, 2, 1), (512, 1, 1), (32, 8, 2), (32, 2, 8), (64, 8, 1), (64, 1, 8), (32, 16, 1), (16, 8, 4), (64, 2, 2), (128, 1, 2), (32, 4, 2)]
for t in tiles:
    try:
        ms = timed(lambda: arms.stencil7_minmaxloc_tiled(A.B200, u.data_ptr(), n, n, n, 0.5, 0.125, out.data_ptr(), t))
        ok = bool(torch.equal(out, ref))
        print(f"B200 tile {str(t):14s} {ms:7.3f} ms  {8.0 * n**3 / ms / 1e6:7.1f} GB/s  {'ok' if ok else 'MISMATCH'}", flush=True)
    except A.ArmsError as e:
        print(f"B200 tile {str(t):14s} error {str(e)[:80]}", flush=True)
ms = timed(lambda: arms.stencil7_minmaxloc(A.CUDA, u.data_ptr(), n, n, n, 0.5, 0.125, out.data_ptr()), reps=3)
print(f"Kokkos::Cuda default tile {ms:7.3f} ms  {8.0 * n**3 / ms / 1e6:7.1f} GB/s  {'ok' if bool(torch.equal(out, ref)) else 'MISMATCH'}")
arms.finalize()
