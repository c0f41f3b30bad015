"""tools/scan_probe.py -- scan bottleneck experiments on a B200 (sweep build): full kernel vs no look-back vs
pure bulk-copy pipeline, and spin back-off.  Diagnostic only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("KOKKOS_B200_LIB", os.path.join(ROOT, "kokkos_b200", "libkokkos_b200_sweep.so"))
import numpy as np, torch
import kokkos_b200 as kb
sys.path.insert(0, os.path.join(ROOT, "tools"))
from sweep import time_it

torch.cuda.set_device(0)
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
space = kb.B200(0, stream=side.cuda_stream)
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
x = torch.randint(-3, 4, (n,), device="cuda", dtype=torch.int64); y = torch.empty_like(x)
vx, vy = space.wrap(x.data_ptr(), n, np.int64), space.wrap(y.data_ptr(), n, np.int64)
tot = torch.zeros(1, device="cuda", dtype=torch.int64)
fn = lambda: space.parallel_scan(vx, vy, total_dev=tot.data_ptr(), blocking=False)
import ctypes
def show(tag):
    best, med = time_it(fn)
    line = f"{tag:60s} best {16*n/best/1e9:8.1f}  med {16*n/med/1e9:8.1f} GB/s"
    if hasattr(space.lib, "b200_debug_scan_stats"):
        st = (ctypes.c_ulonglong * 16)()
        space.lib.b200_debug_scan_stats(None, 1)
        fn(); torch.cuda.synchronize()
        space.lib.b200_debug_scan_stats(st, 1)
        lb, steps, polls, cyc, wpre, wdat, tiles, wagg = [int(v) for v in st[:8]]
        if lb:
            line += (f" | per tile: steps {steps/lb:5.2f} miss-polls {polls/lb:6.2f} lookback {cyc/lb/1.9e3:5.2f}us"
                     f" compute-wait-prefix {wpre/max(tiles,1)/1.9e3:5.2f}us compute-wait-data {wdat/max(tiles,1)/1.9e3:5.2f}us lb-wait-agg {wagg/max(lb,1)/1.9e3:5.2f}us")
    print(line, flush=True)
configs = {
    4: ((256, 9, 3), (256, 7, 4), (128, 9, 4), (512, 9, 3), (128, 9, 6)),
    3: ((256, 9, 3), (128, 9, 6)),
    2: ((128, 9, 4), (256, 7, 4), (256, 9, 3)),
}
for ws in (4, 3, 2):
    kb.tune_set("scan.ws", ws)
    for (bl, nv, nb) in configs[ws]:
        kb.tune_set("scan.block", bl); kb.tune_set("scan.nv", nv); kb.tune_set("scan.nbuf", nb)
        for lbw in (1, 2, 4):
            kb.tune_set("scan.lbw", lbw)
            for dbg, sl in ((0, 0), (8, 0), (0, 300)):
                kb.tune_set("scan.dbg", dbg); kb.tune_set("scan.sleep", sl)
                try:
                    show(f"ws={ws} block={bl} nv={nv} nbuf={nb} lbw={lbw} dbg={dbg} sleep={sl}")
                except kb.B200Error as e:
                    if e.code not in (-3, 9, 1): raise
kb.tune_set("scan.dbg", 0); kb.tune_set("scan.sleep", 0)
