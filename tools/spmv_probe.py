"""tools/spmv_probe.py -- C5b SpMV tuning probe on a B200: vector length x rows-per-group of b200_spmv_crs_f64."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import kokkos_b200 as kb  # noqa: E402
from tools.configs_bench import time_it  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
side = torch.cuda.Stream()
torch.cuda.set_stream(side)
space = kb.B200(0, stream=side.cuda_stream)
R, K = 1 << 22, 32
nnz = R * K
row_map = torch.arange(0, nnz + 1, K, dtype=torch.int64, device=dev)
g = torch.Generator(device=dev)
g.manual_seed(7)
rws = torch.arange(R, dtype=torch.int64, device=dev).repeat_interleave(K)
band = (rws + torch.randint(-64, 65, (nnz,), device=dev, generator=g)).clamp_(0, R - 1)
rnd = torch.randint(0, R, (nnz,), device=dev, generator=g)
sel = (torch.arange(nnz, device=dev) % 2) == 0
col = torch.where(sel, band, rnd).to(torch.int32)
del rws, band, rnd, sel
val = torch.rand(nnz, dtype=torch.float64, device=dev, generator=g)
x = torch.rand(R, dtype=torch.float64, device=dev, generator=g)
y = torch.empty(R, dtype=torch.float64, device=dev)
nbytes = nnz * 12 + R * 16 + 8 * R
V = lambda t, dt: space.wrap(t.data_ptr(), t.numel(), dt)  # noqa: E731
ref = None
for vl, ur in ((32, 1), (16, 1), (8, 1), (8, 2), (4, 1), (4, 2)):
    kb.tune_set("spmv.vl", vl); kb.tune_set("spmv.ur", ur)
    fn = lambda: space.spmv_crs(V(row_map, np.int64), V(col, np.int32), V(val, np.float64), V(x, np.float64), V(y, np.float64))  # noqa: E731
    fn(); torch.cuda.synchronize()
    if ref is None:
        ref = y.clone()
    ok = bool(torch.allclose(y, ref, rtol=1e-12, atol=0))
    b, m = time_it(fn, side, 10)
    print(f"vl={vl:2d} ur={ur}  best {b:.3f} med {m:.3f} ms  {nbytes / m / 1e6:7.1f} GB/s  {'OK' if ok else 'MISMATCH'}", flush=True)
