"""tools/sweep.py -- tuning sweep on a real B200 (run under gpurun).  Uses the sweep build of the library
(make -C kokkos_b200/csrc sweep) in which every tuning combination is compiled in, times each with CUDA events
on the launching stream (inputs larger than L2, 3 warm-ups, best and median of `reps`), prints a table and
writes gpurun_out/sweep_<what>.json.  Never used for reported bench values."""
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("KOKKOS_B200_LIB", os.path.join(ROOT, "kokkos_b200", "libkokkos_b200_sweep.so"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import kokkos_b200 as kb  # noqa: E402


def time_it(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 28
    torch.cuda.set_device(0)
    # torch's default stream has handle 0 (= "make me a new stream" for b200_instance_create):
    # run everything on an explicit side stream so the CUDA events bracket our kernels
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    space = kb.B200(0, stream=side.cuda_stream)
    n = 1 << log2n
    out = {}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)

    def run_grid(name, knobs, fn, bytes_per_call):
        rows = []
        keys = list(knobs)
        for combo in itertools.product(*[knobs[k] for k in keys]):
            for k, v in zip(keys, combo):
                kb.tune_set(k, v)
            try:
                best, med = time_it(fn)
            except kb.B200Error as e:
                if e.code in (-3, 9, 1):  # not compiled in / invalid launch configuration (too many threads or smem)
                    continue
                raise
            rows.append({**dict(zip(keys, combo)), "best_ms": best * 1e3, "med_ms": med * 1e3,
                         "best_GBs": bytes_per_call / best / 1e9, "med_GBs": bytes_per_call / med / 1e9})
        rows.sort(key=lambda r: -r["med_GBs"])
        print(f"== {name}: n=2^{log2n}, {bytes_per_call/1e9:.3f} GB/call; top 12 of {len(rows)}")
        for r in rows[:12]:
            print("  ", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items()})
        out[name] = rows
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"sweep_{what}.json"), "w"), indent=1)

    if what in ("all", "reduce"):
        x = torch.randint(0, 100, (n,), device="cuda", dtype=torch.int64).double()
        v = space.wrap(x.data_ptr(), n, np.float64)
        res = torch.zeros(1, device="cuda", dtype=torch.float64)
        run_grid("reduce_sum_f64", {"reduce.vbytes": [8, 16, 32], "reduce.block": [128, 256, 512],
                                    "reduce.unroll": [1, 2, 4, 8, 16], "reduce.bps": [0, 2, 4]},
                 lambda: space.parallel_reduce_sum(v, result_dev=res.data_ptr(), blocking=False), 8 * n)
        expect = float(x.sum().item())
        assert res.item() == expect, (res.item(), expect)
        del x
    if what in ("all", "scan"):
        x = torch.randint(-3, 4, (n,), device="cuda", dtype=torch.int64)
        y = torch.empty_like(x)
        vx, vy = space.wrap(x.data_ptr(), n, np.int64), space.wrap(y.data_ptr(), n, np.int64)
        tot = torch.zeros(1, device="cuda", dtype=torch.int64)
        run_grid("scan_excl_i64", {"scan.ws": [0, 1], "scan.block": [128, 256, 512, 1024], "scan.nv": [3, 5, 7, 9, 11, 13],
                                   "scan.nbuf": [2, 3, 4], "scan.lbw": [1, 2, 4, 8], "scan.bps": [0]},
                 lambda: space.parallel_scan(vx, vy, total_dev=tot.data_ptr(), blocking=False), 16 * n)
        assert tot.item() == int(x.sum().item())
        ref = torch.cumsum(x, 0) - x
        assert torch.equal(ref, y)
        del x, y, ref
    if what in ("all", "stream"):
        a = torch.ones(n, device="cuda", dtype=torch.float64)
        b = torch.full((n,), 2.0, device="cuda", dtype=torch.float64)
        c = torch.zeros(n, device="cuda", dtype=torch.float64)
        va, vb, vc = (space.wrap(t.data_ptr(), n, np.float64) for t in (a, b, c))
        knobs = {"stream.vbytes": [8, 16, 32], "stream.block": [128, 256, 512, 1024], "stream.unroll": [1, 2, 4, 8],
                 "stream.bps": [0, 8]}
        run_grid("stream_copy", knobs, lambda: space.stream_copy(va, vb), 16 * n)
        run_grid("stream_triad", knobs, lambda: space.stream_triad(va, vb, vc, 3.0), 24 * n)
        # comparator: torch's copy kernel, the thing MEASURED_PEAKS.json was taken with
        best, med = time_it(lambda: b.copy_(a))
        print(f"== torch copy_: best {16*n/best/1e9:.1f} GB/s, median {16*n/med/1e9:.1f} GB/s")
        out["torch_copy"] = {"best_GBs": 16 * n / best / 1e9, "med_GBs": 16 * n / med / 1e9}
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"sweep_{what}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
