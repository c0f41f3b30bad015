#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 execution space (contract: see the task statement / DESIGN.md section 7).

Workload (BASELINE.json metric "parallel_reduce/scan HBM GB/s", target sentence: 2^30-element Views):
  one STEP = parallel_reduce Sum<double> over a View<double*> of n elements
           + parallel_scan (exclusive, with total) over a View<int64_t*> of n elements, n = 2^30 PER GPU.
  Algorithmic bytes per step and GPU = 8 n (reduce) + 16 n (scan: 8 read + 8 written)  [SURVEY.md 8(d) C1, C3].
  value = (24 n * n_gpus) / (time of K steps / K), in GB/s; weak scaling (per-GPU shard fixed).
Multi-GPU (one process per GPU, torchrun): the index range is sharded contiguously; the reduce partial is combined
with an NCCL all-reduce; the distributed scan is reduce-totals -> NCCL all-gather -> seeded local scan.

Legs:
  value / roofline : inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e              : same step through the host-buffer C-ABI entry points (pinned host memory, chunked H2D/D2H inside)
  cpu_baseline     : the UNMODIFIED reference (Kokkos::OpenMP, oracle/_ref) on this box's host cores, bounded sample
  --impl reference : that CPU implementation alone, same JSON shape
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG2N_DEFAULT = 30
CPU_SAMPLE_LOG2N = 27


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return max(1, os.cpu_count() or 1)


class ClockSampler:
    """nvidia-smi sampler for the timed region (B200_PROFILING.md 'clocks' line)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.rows = []
        self.proc = None
        self.device = device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_leg(log2n, reps):
    """Time the unmodified reference (Kokkos::OpenMP) on this host: reduce Sum<double> + exclusive scan int64 over
    2^log2n elements each (a bounded sample of the same step).  Returns dict for the JSON line."""
    from oracle.bindings import Ref, Port, ref_available
    n = 1 << log2n
    if ref_available():
        # every host core this process may run on, stated explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers and
        # the OpenMP runtime would inherit it (round-1 SCALE runs measured a 1-thread reference at N > 1)
        cores = host_cores()
        os.environ["OMP_NUM_THREADS"] = str(cores)
        ref = Ref(cores)
        t_red, s = ref.time_reduce_sum_f64(n, reps)
        t_scan, total = ref.time_scan_excl_i64(n, reps)
        kind, cores = "reference", ref.threads
    else:  # the plain-C restatement, one thread
        import numpy as np
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import workloads as W
        port = Port()
        x = W.c1_exact(n); xi = W.c3_small(n)
        t_red = t_scan = 1e30
        for _ in range(max(reps, 1)):
            t0 = time.perf_counter(); port.reduce("sum", x, 1); t_red = min(t_red, time.perf_counter() - t0)
            t0 = time.perf_counter(); port.scan(xi, False, 0, 1); t_scan = min(t_scan, time.perf_counter() - t0)
        kind, cores = "port", 1
    gbs = 24.0 * n / (t_red + t_scan) / 1e9
    return {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind,
            "sample": f"reduce Sum<double> + exclusive scan int64 over 2^{log2n} elements each, best of {reps} after 1 warm-up, "
                      f"Kokkos::OpenMP {cores} threads" if kind == "reference" else f"oracle port, 1 thread, 2^{log2n} elements",
            "reduce_GBs": 8.0 * n / t_red / 1e9, "scan_GBs": 16.0 * n / t_scan / 1e9}, (t_red + t_scan)


def run_reference(args, rank):
    if rank != 0:
        return
    log2n = CPU_SAMPLE_LOG2N
    n = 1 << log2n
    from oracle.bindings import Ref, ref_available
    for _ in range(max(args.warmup, 0)):
        pass  # warm-up is inside the reference driver (1 untimed call per measurement)
    t0 = time.perf_counter()
    cb, step_s = cpu_reference_leg(log2n, max(args.steps, 1))
    line = {"impl": "reference", "metric": "parallel_reduce+parallel_scan HBM throughput", "value": cb["value"], "unit": "GB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+int64", "data": "synthetic",
            "config": {"workload": f"per GPU: parallel_reduce Sum<double> over View<double*> 2^{LOG2N_DEFAULT} + parallel_scan exclusive over "
                                   f"View<int64_t*> 2^{LOG2N_DEFAULT} (BASELINE.json configs[0] functor at the target size + configs[2])",
                       "sample": f"each step = the same two calls on 2^{log2n} elements per View on the host (bounded sample of the workload); "
                                 "throughput = algorithmic bytes / time, directly comparable",
                       "policy": "RangePolicy<Kokkos::OpenMP>", "elements_per_view_in_sample": n, "algorithmic_bytes_per_step": 24 * n,
                       "timing": "Kokkos::Timer around each call, best of `steps` after 1 warm-up; all host threads"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=LOG2N_DEFAULT, help="elements per GPU and per View = 2^log2n")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3

    import numpy as np
    import torch
    import torch.distributed as dist
    import kokkos_b200 as kb

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = 1 << args.log2n
    side = torch.cuda.Stream(device=dev)   # the stream our kernels, the NCCL calls and the timing events share
    torch.cuda.set_stream(side)
    space = kb.B200(local_rank, stream=side.cuda_stream)

    # ---- synthetic shards, generated on the device (C1 (i) integer-valued doubles, C3 hash%7-3)
    xd = torch.empty(n, dtype=torch.float64, device=dev)
    xi = torch.empty(n, dtype=torch.int64, device=dev)
    yi = torch.empty(n, dtype=torch.int64, device=dev)
    base = rank * n
    CH = 1 << 26
    for c in range(0, n, CH):
        m = min(CH, n - c)
        idx = torch.arange(base + c, base + c + m, dtype=torch.int64, device=dev)
        h = (idx * 2654435761) >> 7
        xd[c:c + m] = (h % 100).double()
        xi[c:c + m] = (h % 7) - 3
        del idx, h
    vxd, vxi, vyi = space.wrap(xd.data_ptr(), n, np.float64), space.wrap(xi.data_ptr(), n, np.int64), space.wrap(yi.data_ptr(), n, np.int64)
    from kokkos_b200.sharded import ShardedB200
    sp = ShardedB200(space, coll_device=dev)   # the range-sharded layer: local kernels + NCCL combines (world 1: no collectives)
    red_dev = torch.zeros(1, dtype=torch.float64, device=dev)
    tot_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    # our kernels per step: reduce(xd) + scan at N=1; reduce(xd) + shard-total reduce(xi) + seeded scan at N>1
    launches_per_step = 2 if not distributed else 3

    scan_ev = []

    def step(record=False):
        # parallel_reduce Sum<double>: result stays on the device (View result => asynchronous); N>1: + NCCL all-reduce
        sp.reduce_sum_async(vxd, out=red_dev)
        # parallel_scan: N>1 = shard totals -> NCCL all-gather -> seeded local scan (the kernel sums the lower ranks' totals)
        hooks = None
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            hooks = (lambda: e0.record(side), lambda: e1.record(side))
            scan_ev.append((e0, e1))
        totals = sp.scan_exclusive_async(vxi, vyi, total_out=tot_dev, around_scan_kernel=hooks)
        return totals

    def sync_all():
        torch.cuda.synchronize(dev)
        if distributed:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    sync_all()
    # correctness of what is being timed (cheap closed-form checks, outside the timed region)
    exp_red = float(xd.sum().item())
    got_red = float(red_dev.item())
    if distributed:
        t = torch.tensor([exp_red], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        exp_red = float(t.item())
    assert got_red == exp_red, ("reduce mismatch", got_red, exp_red)
    totals = step()
    torch.cuda.synchronize(dev)
    seed_exp = int(totals[:rank].sum().item()) if distributed else 0
    chk = torch.cumsum(xi[: 1 << 20], 0) - xi[: 1 << 20] + seed_exp
    assert torch.equal(chk, yi[: 1 << 20]), "scan mismatch in the first 2^20 outputs"
    assert int(tot_dev.item()) == int(xi.sum().item()), "scan total mismatch"
    del chk

    sampler = ClockSampler(local_rank)
    sync_all()
    sampler.start()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(side)
    for _ in range(args.steps):
        step(record=True)
    t1.record(side)
    sync_all()
    clocks = sampler.stop()
    ms_total = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    scan_ms = sum(a.elapsed_time(b) for a, b in scan_ev) / len(scan_ev)
    if distributed:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms_total.item()) / args.steps
    bytes_per_step = 24.0 * n * world
    value = bytes_per_step / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (the scan: 2/3 of the step's bytes), measured live with CUDA events
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 (of fallback)"
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_scan_ncu_summary.json"))).get("dram_bytes_per_launch_at_2^30")
    except Exception:
        pass
    achieved = 16.0 * n / (scan_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "contig_scan_ws2_kernel<int64> (single-pass scan)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": 16.0 * n, "avg_launch_ms": scan_ms, "frac_of_8TBs_nominal": achieved / 8000.0}

    # ---- e2e: the same step through the host-buffer C-ABI calls (pinned host memory; copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        try:
            hx = torch.empty(n, dtype=torch.float64, pin_memory=True)
            hi = torch.empty(n, dtype=torch.int64, pin_memory=True)
            hy = torch.empty(n, dtype=torch.int64, pin_memory=True)
            hx.copy_(xd); hi.copy_(xi)
            torch.cuda.synchronize(dev)
            e2e_steps = max(2, min(args.steps, 3))

            def e2e_step():
                r = space.parallel_reduce_sum_host(hx.data_ptr(), n)
                t = space.parallel_scan_host(hi.data_ptr(), hy.data_ptr(), n, 0)
                return r, t
            e2e_step()  # warm-up (allocates the staging buffers)
            sync_all()
            w0 = time.perf_counter()
            for _ in range(e2e_steps):
                r, t = e2e_step()
            torch.cuda.synchronize(dev)
            w = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
            if distributed:
                dist.all_reduce(w, op=dist.ReduceOp.MAX)
            assert r == float(xd.sum().item()) and t == int(xi.sum().item())
            assert torch.equal(hy[: 1 << 20].to(dev), torch.cumsum(xi[: 1 << 20], 0) - xi[: 1 << 20])
            e2e = {"value": bytes_per_step / (float(w.item()) / e2e_steps) / 1e9, "unit": "GB/s",
                   "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 8 * n + 16, "steps": e2e_steps,
                   "api": "b200_reduce_sum_f64_host + b200_scan_excl_i64_host (pinned host buffers, 64 MiB chunks, double-buffered)",
                   "note": "per-GPU shards are independent on this leg (no cross-GPU seed): PCIe-bound"}
            del hx, hi, hy
        except Exception as ex:  # e.g. not enough pinnable host memory
            e2e = {"value": None, "unit": "GB/s", "error": repr(ex)[:200]}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu_baseline, _ = cpu_reference_leg(CPU_SAMPLE_LOG2N, 5)
        except Exception as ex:
            cpu_baseline = {"value": None, "error": repr(ex)[:200]}

    if rank == 0:
        line = {"metric": "parallel_reduce+parallel_scan HBM throughput", "value": value, "unit": "GB/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64+int64", "data": "synthetic",
                "config": {"workload": f"per GPU: parallel_reduce Sum<double> over View<double*> 2^{args.log2n} + parallel_scan exclusive over "
                                       f"View<int64_t*> 2^{args.log2n} (BASELINE.json configs[0] functor at the target size + configs[2])",
                           "policy": "RangePolicy", "elements_per_gpu_per_view": n, "algorithmic_bytes_per_step_per_gpu": 24 * n,
                           "parallelism": f"index-range sharded x{world}" + ("; NCCL all-reduce (reduce), all-gather of shard totals + seeded scan (scan)" if distributed else ""),
                           "l2": "inputs (8 GiB per View) are far larger than the 126 MB L2; no flush needed",
                           "timing": "CUDA events on the launching stream, barrier+synchronize both sides, max over ranks"},
                "frac_of_measured_peak": value / (peak * world), "frac_of_8TBs_nominal": value / (8000.0 * world),
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
                "clocks": clocks,
                "breakdown": {"scan_GBs_per_gpu": achieved, "reduce_plus_collectives_ms": ms_per_step - scan_ms, "scan_ms": scan_ms}}
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
