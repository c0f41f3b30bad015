"""tools/scan_pfd_probe.py -- L2 prefetch distance of the typed scan (tune key scan.pfd, in tiles of 2304 int64) on a B200.
Prints GB/s per distance for 2^30 int64; every element checked against torch.cumsum once per distance.  Diagnostic only."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch  # noqa: E402,E401
import kokkos_b200 as kb  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tools"))
from sweep import time_it  # noqa: E402

torch.cuda.set_device(0)
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
space = kb.B200(0, stream=side.cuda_stream)
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 30)
x = torch.randint(-3, 4, (n,), device="cuda", dtype=torch.int64); y = torch.empty_like(x)
vx, vy = space.wrap(x.data_ptr(), n, np.int64), space.wrap(y.data_ptr(), n, np.int64)
tot = torch.zeros(1, device="cuda", dtype=torch.int64)
fn = lambda: space.parallel_scan(vx, vy, total_dev=tot.data_ptr(), blocking=False)  # noqa: E731


def check():
    y.fill_(-12345)
    fn(); torch.cuda.synchronize()
    run = torch.zeros((), dtype=torch.int64, device="cuda")
    CH = 1 << 26
    for c in range(0, n, CH):
        xb = x[c:c + CH]
        if not torch.equal(torch.cumsum(xb, 0) - xb + run, y[c:c + CH]):
            return False
        run = run + xb.sum()
    return int(tot.item()) == int(run.item())


for d in [int(a) for a in (sys.argv[2].split(",") if len(sys.argv) > 2 else "0,148,444,888,1776,3552,7104".split(","))]:
    kb.tune_set("scan.pfd", d)
    ok = check()
    best, med = time_it(fn)
    print(f"scan.pfd={d:5d} tiles ({d * 18432 / 1e6:6.1f} MB ahead)  best {16*n/best/1e9:8.1f} med {16*n/med/1e9:8.1f} GB/s  parity={'ok' if ok else 'MISMATCH'}", flush=True)
kb.tune_set("scan.pfd", 0)
