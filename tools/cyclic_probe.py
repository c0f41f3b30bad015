"""Throughput probe of the fused block-cyclic scan (b200_comm_scan_excl_i64) at WORLD ranks, one process per GPU.
usage: python tools/cyclic_probe.py --world 2 [--log2n 30] [--reps 10]     (spawns the ranks itself)
Each rank prints its device-timed GB/s (16 B per local element); the parent prints the slowest rank = the job's number."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, world, uid, log2n, reps):
    import numpy as np
    import torch
    import kokkos_b200 as kb
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    side = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(side)
    space = kb.B200(rank, stream=side.cuda_stream)
    comm = kb.Comm(space, rank, world, uid)
    for kv in os.environ.get("KB200_TUNE", "").split(","):  # e.g. KB200_TUNE=comm.scan_algo=1,comm.tpr=512
        if "=" in kv:
            kb.tune_set(kv.split("=")[0], int(kv.split("=")[1]))
    n_global = (1 << log2n) * world
    block, n_local, nsteps = comm.cyclic_layout(n_global, np.int64)
    x = torch.randint(-3, 4, (n_local,), dtype=torch.int64, device=dev)
    y = torch.empty_like(x)
    tot = torch.zeros(1, dtype=torch.int64, device=dev)
    vx, vy = space.wrap(x.data_ptr(), n_local, np.int64), space.wrap(y.data_ptr(), n_local, np.int64)
    res = {"rank": rank, "n_local": n_local, "block": block, "nsteps": nsteps}

    def timed(fn, label):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev)
        comm.host_barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(side)
        for _ in range(reps):
            fn()
        e1.record(side)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / reps
        res[label + "_ms"] = ms
        res[label + "_GBs"] = 16.0 * n_local / ms / 1e6

    timed(lambda: comm.parallel_scan(vx, vy, n_global, total_dev=tot.data_ptr(), blocking=False), "cyclic")
    # parity of what was timed: the global total and this rank's first block against torch
    s = torch.zeros(1, dtype=torch.int64, device=dev)
    s[0] = x.sum()
    comm.allreduce("sum", s.data_ptr(), 1, np.int64)
    torch.cuda.synchronize(dev)
    assert int(tot.item()) == int(s.item()), (int(tot.item()), int(s.item()))
    m = min(n_local, block)
    loc = torch.cumsum(x[:m], 0) - x[:m]
    assert torch.equal(y[:m] - y[0], loc), "first local block mismatch"
    if world == 1:
        timed(lambda: space.parallel_scan(vx, vy, total_dev=tot.data_ptr(), blocking=False), "ws2")
        assert torch.equal(y[:m], loc)
        kb.tune_set("scan.chunked", 1)
        timed(lambda: space.parallel_scan(vx, vy, total_dev=tot.data_ptr(), blocking=False), "chunked_entry")
        kb.tune_set("scan.chunked", 0)
    if hasattr(space.lib, "b200_debug_chunk_stats"):  # sweep build: per-step cycle counts of CTA 0 and CTA G-1 from the LAST launch
        import ctypes
        buf = (ctypes.c_uint64 * 32)()
        comm.parallel_scan(vx, vy, n_global, total_dev=tot.data_ptr(), blocking=False)
        torch.cuda.synchronize(dev)
        space.lib.b200_debug_chunk_stats(buf)
        names = ["load_wait_empty", "store_wait_out", "store_drain", "sync_local", "sync_remote", "red_wait_full", "red_work",
                 "scan_wait_full", "scan_wait_prefix", "scan_finish"]
        for base, who in ((0, "cta0"), (16, "ctaLast")):
            steps = max(1, buf[base + 10])
            res["cyc_per_step_" + who] = {nm: round(buf[base + i] / steps) for i, nm in enumerate(names)}
    if hasattr(space.lib, "b200_debug_comm_scan_stats") and not os.environ.get("KB200_TUNE", "").count("scan_algo=1"):
        import ctypes
        st = (ctypes.c_ulonglong * 16)()
        space.lib.b200_debug_comm_scan_stats(None, 1)
        comm.host_barrier()
        comm.parallel_scan(vx, vy, n_global, total_dev=tot.data_ptr(), blocking=False)
        torch.cuda.synchronize(dev)
        space.lib.b200_debug_comm_scan_stats(st, 1)
        v = [int(a) for a in st[:12]]
        lb, tiles = max(v[0], 1), max(v[6], 1)
        res["rounds_stats_us"] = {"lookback_steps_per_tile": round(v[1] / lb, 2), "lookback": round(v[3] / lb / 1.9e3, 2), "compute_wait_prefix": round(v[4] / tiles / 1.9e3, 2),
                                  "compute_wait_data": round(v[5] / tiles / 1.9e3, 2), "lb_wait_agg": round(v[7] / lb / 1.9e3, 2),
                                  "base_nonleader_per_tile": round(v[8] / tiles / 1.9e3, 3), "base_leader_each": round(v[9] / max(v[10], 1) / 1.9e3, 2),
                                  "leaders": v[10], "late_bases": v[11], "tiles": v[6]}
    if hasattr(space.lib, "b200_debug_round_ts"):
        import ctypes
        ts = np.zeros((4, 8192), dtype=np.uint64)
        space.lib.b200_debug_round_ts(ctypes.c_void_p(ts.ctypes.data))
        nr = int(min(nsteps, 8192))
        t = ts[:, 8:nr - 8].astype(np.int64)  # steady state
        res["round_ns"] = {"period": float(np.mean(np.diff(t[3]))), "first_to_last_landed": float(np.mean(t[1] - t[0])),
                           "last_landed_to_agg_sent": float(np.mean(t[2] - t[1])), "agg_sent_to_next_base": float(np.mean(t[3][1:] - t[2][:-1])),
                           "first_landed_to_base": float(np.mean(t[3] - t[0]))}
    res["comm_error"] = comm.error()
    print("RESULT " + json.dumps(res), flush=True)
    comm.finalize()
    space.finalize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=1)
    ap.add_argument("--log2n", type=int, default=30, help="elements per GPU = 2^log2n")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--rank", type=int, default=-1)
    ap.add_argument("--uid", default="")
    a = ap.parse_args()
    if a.rank >= 0:
        worker(a.rank, a.world, a.uid, a.log2n, a.reps)
        return
    import kokkos_b200 as kb
    uid = kb.comm_unique_id()
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--world", str(a.world), "--log2n", str(a.log2n), "--reps", str(a.reps),
                               "--rank", str(r), "--uid", uid], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(a.world)]
    rows = []
    for p in procs:
        out, _ = p.communicate(timeout=900)
        for line in out.splitlines():
            if line.startswith("RESULT "):
                rows.append(json.loads(line[7:]))
        if p.returncode != 0:
            print(out[-2000:])
    if rows:
        slow = max(r["cyclic_ms"] for r in rows)
        print(json.dumps({"world": a.world, "log2n_per_gpu": a.log2n, "cyclic_ms_max": slow, "per_gpu_GBs": 16.0 * rows[0]["n_local"] / slow / 1e6,
                          "job_GBs": 16.0 * sum(r["n_local"] for r in rows) / slow / 1e6, "ranks": rows}))


if __name__ == "__main__":
    main()
