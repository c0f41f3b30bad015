"""tools/gups_probe.py -- what bounds GUPS (C5a) on a B200: the same b200_gups_add_i64 kernel, the same 2^26 updates into the same
2^30-entry int64 table (8 GiB), with the update stream (a) random, (b) binned by 2 MiB region (random inside a region), (c) binned by
4 KiB, (d) fully sorted.  Atomic adds commute, so every ordering produces the same table (asserted).  Diagnostic only."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch  # noqa: E402,E401
import kokkos_b200 as kb  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tools"))
from sweep import time_it  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
space = kb.B200(0, stream=side.cuda_stream)
tl, m = 1 << 30, 1 << 26
g = torch.Generator(device=dev); g.manual_seed(20230913)
idx = torch.randint(0, tl, (m,), dtype=torch.int64, device=dev, generator=g)
table = torch.zeros(tl, dtype=torch.int64, device=dev)
vt = space.wrap(table.data_ptr(), tl, np.int64)


def binned(shift):
    key = idx >> shift
    order = torch.argsort(key, stable=True)
    return idx[order].contiguous()


ref = None
for name, stream in (("random", idx), ("binned by 2 MiB region", binned(18)), ("binned by 4 KiB", binned(9)), ("binned by 32 B sector", binned(2)), ("sorted", torch.sort(idx).values)):
    vi = space.wrap(stream.data_ptr(), m, np.int64)
    table.zero_()
    space.gups(vt, vi, 1, "add"); torch.cuda.synchronize()
    if ref is None:
        ref = table.clone()
    ok = bool(torch.equal(table, ref))
    best, med = time_it(lambda: space.gups(vt, vi, 1, "add"))
    print(f"{name:26s} med {med*1e3:7.3f} ms  {m/med/1e9:7.1f} GUP/s  {'same table' if ok else 'MISMATCH'}", flush=True)
