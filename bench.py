#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 execution space (contract: see the task statement / DESIGN.md section 7).

Workload (BASELINE.json metric "parallel_reduce/scan HBM GB/s", target sentence: 2^30-element Views):
  one STEP = parallel_reduce Sum<double> over a View<double*> of n elements
           + parallel_scan (exclusive, with total) over a View<int64_t*> of n elements, n = 2^30 PER GPU.
  Algorithmic bytes per step and GPU = 8 n (reduce) + 16 n (scan: 8 read + 8 written)  [SURVEY.md 8(d) C1, C3].
  value = (24 n * n_gpus) / (time of K steps / K), in GB/s; weak scaling (per-GPU shard fixed).
Multi-GPU (one process per GPU, torchrun): the reduce View is sharded contiguously and its partial combined with an NCCL
all-reduce; the scan View is distributed BLOCK-CYCLICALLY and scanned by ONE fused kernel per rank whose round aggregates
travel over peer-mapped NVLink mailboxes (b200_comm_scan_excl_i64, 16 B/element at any N; kokkos_b200/csrc/comm.cu).

Legs on the JSON line:
  value / roofline : inputs resident in HBM, CUDA events on the launching stream, max over ranks
  breakdown.configs: every other BASELINE.json config (C2 stream copy/triad, C4 MDRange stencil + MinMaxLoc, C5a GUPS,
                     C5b TeamPolicy SpMV) at this N, per-GPU shards, bit-exact in-bench asserts
  kokkos_api       : the same workloads as Kokkos USER CODE (KOKKOS_LAMBDA on the unmodified reference headers) on Kokkos::B200
  comparators      : that same user code on the reference's Kokkos::Cuda (sm_100) and on CUB, same buffers, same stream
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
HASH_MUL = 2654435761


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


def workload_text(log2n):
    return (f"per GPU: parallel_reduce Sum<double> over View<double*> 2^{log2n} + parallel_scan exclusive over "
            f"View<int64_t*> 2^{log2n} (BASELINE.json configs[0] functor at the target size + configs[2])")


def run_reference(args, rank):
    if rank != 0:
        return
    log2n = CPU_SAMPLE_LOG2N
    n = 1 << log2n
    t0 = time.perf_counter()
    cb, step_s = cpu_reference_leg(log2n, max(args.steps, 1))  # warm-up is inside the reference driver (1 untimed call per measurement)
    line = {"impl": "reference", "metric": "parallel_reduce+parallel_scan HBM throughput", "value": cb["value"], "unit": "GB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+int64", "data": "synthetic",
            "config": {"workload": workload_text(LOG2N_DEFAULT),
                       "sample": f"each step = the same two calls on 2^{log2n} elements per View on the host (bounded sample of the workload); "
                                 "throughput = algorithmic bytes / time, directly comparable",
                       "policy": "RangePolicy<Kokkos::OpenMP>", "elements_per_view_in_sample": n, "algorithmic_bytes_per_step": 24 * n,
                       "timing": "Kokkos::Timer around each call, best of `steps` after 1 warm-up; all host threads"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


class Bench:
    """State shared by the legs of one rank."""

    def __init__(self, args):
        import numpy as np
        import torch
        import torch.distributed as dist
        import kokkos_b200 as kb
        from kokkos_b200.sharded import ShardedB200
        self.np, self.torch, self.dist, self.kb = np, torch, dist, kb
        self.args = args
        self.rank, self.world, self.local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.distributed = self.world > 1
        if self.distributed:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        # the stream our kernels, the NCCL calls and the timing events share
        self.side = torch.cuda.Stream(device=self.dev)
        torch.cuda.set_stream(self.side)
        for kv in os.environ.get("KB200_TUNE", "").split(","):  # probe knobs, e.g. KB200_TUNE=for.bps=0,comm.tpr=256
            if "=" in kv:
                kb.tune_set(kv.split("=")[0], int(kv.split("=")[1]))
        self.space = kb.B200(self.local_rank, stream=self.side.cuda_stream)
        self.comm = None
        if self.distributed:
            uid = [kb.comm_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            self.comm = kb.Comm(self.space, self.rank, self.world, uid[0])
        self.sp = ShardedB200(self.space, coll_device=self.dev, comm=self.comm)
        self.steps, self.warmup = args.steps, max(args.warmup, 3)  # timing rule: W >= 3
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.peak = float(peaks.get("hbm_gbs", 6650.0))
        self.peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 (of fallback)"
        self.arms = None
        if not args.no_arms:
            try:
                from benchlib import arms as A
                if A.available():
                    self.A = A
                    self.arms = A.Arms(self.local_rank, self.side.cuda_stream)
            except Exception as ex:  # the comparators are optional evidence; the product path never depends on them
                self.arms_error = repr(ex)[:200]

    # ---- helpers -----------------------------------------------------------------------------------------------------
    def wrap(self, t):
        np = self.np
        dt = {self.torch.float64: np.float64, self.torch.int64: np.int64, self.torch.int32: np.int32}[t.dtype]
        return self.space.wrap(t.data_ptr(), t.numel(), dt)

    def sync_all(self):
        self.torch.cuda.synchronize(self.dev)
        if self.distributed:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.distributed:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps=None, warmup=None, barrier=True):
        """avg ms per call of `fn` over `steps` calls after `warmup`, CUDA events on the launching stream, max over ranks."""
        torch = self.torch
        steps = steps or self.steps
        for _ in range(self.warmup if warmup is None else warmup):
            fn()
        if barrier:
            self.sync_all()
        else:
            torch.cuda.synchronize(self.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.side)
        for _ in range(steps):
            fn()
        e1.record(self.side)
        if barrier:
            self.sync_all()
            return self.max_over_ranks(e0.elapsed_time(e1)) / steps
        torch.cuda.synchronize(self.dev)
        return e0.elapsed_time(e1) / steps

    def hashed(self, gidx):
        return (gidx * HASH_MUL) >> 7

    # ---- headline data -----------------------------------------------------------------------------------------------
    def make_headline(self):
        """xd: contiguous shard of the reduce View (C1 (i): integer-valued doubles).  xi/yi: this rank's blocks of the
        block-cyclic scan View (C3: hash % 7 - 3); at world 1 the block-cyclic layout IS the contiguous one."""
        torch, np = self.torch, self.np
        n = 1 << self.args.log2n
        self.n = n
        self.n_global = n * self.world
        if self.comm is not None:
            self.block, self.n_local, _ = self.comm.cyclic_layout(self.n_global, np.int64)
        else:
            self.block, self.n_local = 1 << 20, n
        self.xd = torch.empty(n, dtype=torch.float64, device=self.dev)
        self.xi = torch.empty(self.n_local, dtype=torch.int64, device=self.dev)
        self.yi = torch.empty(self.n_local, dtype=torch.int64, device=self.dev)
        CH = 1 << 26
        for c in range(0, n, CH):
            m = min(CH, n - c)
            idx = torch.arange(self.rank * n + c, self.rank * n + c + m, dtype=torch.int64, device=self.dev)
            self.xd[c:c + m] = (self.hashed(idx) % 100).double()
            del idx
        for c in range(0, self.n_local, CH):
            m = min(CH, self.n_local - c)
            l = torch.arange(c, c + m, dtype=torch.int64, device=self.dev)
            g = (torch.div(l, self.block, rounding_mode="floor") * self.world + self.rank) * self.block + l % self.block
            self.xi[c:c + m] = (self.hashed(g) % 7) - 3
            del l, g
        self.vxd, self.vxi, self.vyi = self.wrap(self.xd), self.wrap(self.xi), self.wrap(self.yi)
        self.red_dev = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.tot_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)

    def check_scan_all_elements(self, y, what):
        """Every output element of the (distributed) exclusive scan against torch: per-block sums of every rank are gathered,
        the exclusive prefix of each block start follows, and inside a block the expected values are a local cumsum."""
        torch, dist = self.torch, self.dist
        B, w = self.block, self.world
        nblk_g = -(-self.n_global // B)
        nblk_max = -(-nblk_g // w)
        sums = torch.zeros(nblk_max, dtype=torch.int64, device=self.dev)
        nfull = self.n_local // B
        if nfull:
            sums[:nfull] = self.xi[:nfull * B].view(nfull, B).sum(1)
        if self.n_local > nfull * B:
            sums[nfull] = self.xi[nfull * B:].sum()
        allsums = torch.zeros(w * nblk_max, dtype=torch.int64, device=self.dev)
        if self.distributed:
            dist.all_gather_into_tensor(allsums, sums)
        else:
            allsums.copy_(sums)
        flat = allsums.view(w, nblk_max).t().contiguous().view(-1)  # global block order: (local block, rank)
        starts = (torch.cumsum(flat, 0) - flat).view(nblk_max, w)[:, self.rank]
        total = int(flat.sum().item())
        G = 64
        for b0 in range(0, nfull, G):
            b1 = min(nfull, b0 + G)
            xb = self.xi[b0 * B:b1 * B].view(b1 - b0, B)
            exp = torch.cumsum(xb, 1) - xb + starts[b0:b1, None]
            assert torch.equal(exp, y[b0 * B:b1 * B].view(b1 - b0, B)), f"{what}: mismatch in local blocks {b0}..{b1}"
            del exp
        if self.n_local > nfull * B:
            xb = self.xi[nfull * B:]
            assert torch.equal(torch.cumsum(xb, 0) - xb + starts[nfull], y[nfull * B:]), f"{what}: mismatch in the last (short) block"
        return total

    # ---- the headline step --------------------------------------------------------------------------------------------
    def headline(self):
        torch = self.torch
        sp = self.sp
        scan_ev = []

        # two execution-space instances (as a Kokkos program would use two Kokkos::Cuda instances): the reduce and its all-reduce
        # can run on stream B while the scan runs on the main stream; every step then forks and joins with events.  Whether that
        # pays depends on N: on one GPU each kernel saturates HBM by itself, with the distributed scan (latency bound) it does.
        self.streamB = torch.cuda.Stream(device=self.dev)
        with torch.cuda.stream(self.streamB):
            self.spaceB = self.kb.B200(self.local_rank, stream=self.streamB.cuda_stream)
            from kokkos_b200.sharded import ShardedB200
            spB = ShardedB200(self.spaceB, coll_device=self.dev)
            vxdB = self.spaceB.wrap(self.xd.data_ptr(), self.n, self.np.float64)
        ev_fork, ev_join = torch.cuda.Event(), torch.cuda.Event()
        mode = {"overlap": self.args.overlap}

        def step(record=False):
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if mode["overlap"]:
                ev_fork.record(self.side)
                self.streamB.wait_event(ev_fork)
                if record:
                    e0.record(self.side)
                sp.cyclic_scan_async(self.vxi, self.vyi, self.n_global, self.tot_dev)   # launched first: it needs the shared memory
                if record:
                    e1.record(self.side)
                with torch.cuda.stream(self.streamB):
                    spB.reduce_sum_async(vxdB, out=self.red_dev)
                    ev_join.record(self.streamB)
                self.side.wait_event(ev_join)
            else:
                # parallel_reduce Sum<double>: result stays on the device (View result => asynchronous); N>1: + NCCL all-reduce
                sp.reduce_sum_async(self.vxd, out=self.red_dev)
                if record:
                    e0.record(self.side)
                # parallel_scan: one kernel per rank (N>1: the fused block-cyclic scan over NVLink mailboxes)
                sp.cyclic_scan_async(self.vxi, self.vyi, self.n_global, self.tot_dev)
                if record:
                    e1.record(self.side)
            if record:
                scan_ev.append((e0, e1))

        if mode["overlap"] < 0:  # auto: 3 untimed-for-the-record steps of each structure, the faster one (max over ranks) is used
            cand = {}
            for ov in (0, 1):
                mode["overlap"] = ov
                cand[ov] = self.timed(step, steps=3, warmup=2)
            mode["overlap"] = 1 if cand[1] < cand[0] else 0
            self.overlap_probe_ms = {"sequential": cand[0], "concurrent": cand[1]}
        self.overlap = mode["overlap"]
        for _ in range(self.warmup):
            step()
        self.sync_all()
        # correctness of what is being timed (outside the timed region): the reduce against torch, EVERY scan element
        exp_red = self.xd.sum().reshape(1)
        if self.distributed:
            self.dist.all_reduce(exp_red)
        assert float(self.red_dev.item()) == float(exp_red.item()), ("reduce mismatch", float(self.red_dev.item()), float(exp_red.item()))
        total = self.check_scan_all_elements(self.yi, "typed scan")
        assert int(self.tot_dev.item()) == total, ("scan total mismatch", int(self.tot_dev.item()), total)
        if self.comm is not None:
            assert self.comm.error() == 0

        sampler = ClockSampler(self.local_rank)
        self.sync_all()
        sampler.start()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(self.side)
        for _ in range(self.steps):
            step(record=True)
        t1.record(self.side)
        self.sync_all()
        self.clocks = sampler.stop()
        self.ms_per_step = self.max_over_ranks(t0.elapsed_time(t1)) / self.steps
        self.scan_ms_in_step = self.max_over_ranks(sum(a.elapsed_time(b) for a, b in scan_ev) / len(scan_ev))
        self.scan_ms = self.scan_ms_in_step
        if self.overlap:
            # the dominant kernel timed ALONE (inside a step it shares the machine with the reduce): same launches, same data
            self.scan_ms = self.timed(lambda: sp.cyclic_scan_async(self.vxi, self.vyi, self.n_global, self.tot_dev), warmup=1)
            self.reduce_ms_alone = self.timed(lambda: sp.reduce_sum_async(self.vxd, out=self.red_dev), warmup=1)
        self.bytes_per_step = 24.0 * self.n * self.world
        self.value = self.bytes_per_step / (self.ms_per_step * 1e-3) / 1e9
        self.launches_per_step = 2  # range_reduce_kernel + contig_scan_ws2_kernel (plain at N=1, rounds form at N>1)

    def roofline(self):
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_scan_ncu_summary.json"))).get("dram_bytes_per_launch_at_2^30")
        except Exception:
            pass
        achieved = 16.0 * self.n_local / (self.scan_ms * 1e-3) / 1e9
        kname = "contig_scan_ws2_kernel<int64,128,9,4,1> (single-pass look-back scan" + (", ROUNDS form: block-cyclic over NVLink mailboxes)" if self.distributed else ")")
        return {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": self.peak, "unit": "GB/s", "frac": achieved / self.peak,
                "traffic": traffic if not self.distributed else None, "peak_source": self.peak_src,
                "algorithmic_bytes_per_launch": 16.0 * self.n_local, "avg_launch_ms": self.scan_ms, "frac_of_8TBs_nominal": achieved / 8000.0,
                "timing": ("kernel timed alone, %d launches with CUDA events right after the timed steps (inside a step it runs concurrently with the "
                           "reduce kernel: %.3f ms there)" % (self.steps, self.scan_ms_in_step)) if self.overlap else "CUDA events around each launch inside the timed steps"}

    # ---- Kokkos user code on Kokkos::B200 (adapter) and the comparators --------------------------------------------
    def kokkos_api_and_comparators(self):
        """The headline pair as Kokkos USER CODE on three arms, per rank on its own shard (no collectives: these legs compare
        kernels), max over ranks.  Results of every arm are checked bit-exact against the typed path's."""
        if self.arms is None:
            return None, None
        A, torch = self.A, self.torch
        n = self.n
        nl = self.n_local if not self.distributed else min(self.n_local, n)  # a contiguous local array of about n elements
        exp_red = float(self.xd.sum().item())
        xs = self.xi[:nl]
        ysc = torch.empty(nl, dtype=torch.int64, device=self.dev)
        # expected local scan: the typed kernel on the same local array (checked against torch in blocks just below)
        self.space.parallel_scan(self.space.wrap(xs.data_ptr(), nl, self.np.int64), self.space.wrap(ysc.data_ptr(), nl, self.np.int64),
                                 total_dev=self.tot_dev.data_ptr(), blocking=False)
        run = torch.zeros((), dtype=torch.int64, device=self.dev)
        CH = 1 << 26
        for c in range(0, nl, CH):
            xb = xs[c:c + CH]
            assert torch.equal(torch.cumsum(xb, 0) - xb + run, ysc[c:c + CH]), "typed local scan mismatch"
            run = run + xb.sum()
        exp_total = int(run.item())
        out = {}
        y2 = torch.empty(nl, dtype=torch.int64, device=self.dev)
        for arm in (A.B200, A.CUDA, A.CUB):
            r = torch.zeros(1, dtype=torch.float64, device=self.dev)
            t = torch.zeros(1, dtype=torch.int64, device=self.dev)
            red_ms = self.timed(lambda: self.arms.reduce_sum(arm, self.xd.data_ptr(), n, r.data_ptr()))
            assert float(r.item()) == exp_red, (A.ARM_NAMES[arm], "reduce", float(r.item()), exp_red)
            y2.fill_(-1)
            scan_ms = self.timed(lambda: self.arms.scan_excl(arm, xs.data_ptr(), y2.data_ptr(), nl, t.data_ptr()))
            assert torch.equal(y2, ysc), (A.ARM_NAMES[arm], "scan output differs from the typed path")
            if arm != A.CUB:
                assert int(t.item()) == exp_total, (A.ARM_NAMES[arm], "scan total")
            y2.fill_(-1)
            std_ms = self.timed(lambda: self.arms.std_exclusive_scan(arm, xs.data_ptr(), y2.data_ptr(), nl))
            assert torch.equal(y2, ysc), (A.ARM_NAMES[arm], "std exclusive_scan output differs from the typed path")
            out[arm] = {"reduce_ms": red_ms, "scan_ms": scan_ms, "reduce_GBs": 8.0 * n / red_ms / 1e6, "scan_GBs": 16.0 * nl / scan_ms / 1e6,
                        "step_GBs_per_gpu": (8.0 * n + 16.0 * nl) / (red_ms + scan_ms) / 1e6,
                        "std_exclusive_scan_ms": std_ms, "std_exclusive_scan_GBs": 16.0 * nl / std_ms / 1e6,
                        "step_GBs_per_gpu_with_std_exclusive_scan": (8.0 * n + 16.0 * nl) / (red_ms + std_ms) / 1e6}
        del y2, ysc
        api = dict(out[A.B200])
        api["what"] = ("Kokkos::parallel_reduce / Kokkos::parallel_scan with KOKKOS_LAMBDA functors over RangePolicy<Kokkos::B200> on the UNMODIFIED "
                       "reference headers (kokkos_b200/adapter; benchlib/kokkos_arms.cu), per GPU on its own shard, results bit-identical to the typed path; "
                       "std_exclusive_scan = Kokkos::Experimental::exclusive_scan(exec, in, out, 0) (blocking: the reference fences), which the adapter "
                       "routes to the typed look-back kernel the way the reference routes its Cuda inclusive_scan to thrust/CUB")
        comp = {"Kokkos::Cuda (reference backend built for sm_100a, same lambdas)": out[A.CUDA],
                "CUB DeviceReduce::Sum + DeviceScan::ExclusiveSum (CUDA 12.9 CCCL)": out[A.CUB],
                "speedup_vs_Kokkos::Cuda": {"reduce": out[A.CUDA]["reduce_ms"] / out[A.B200]["reduce_ms"], "scan": out[A.CUDA]["scan_ms"] / out[A.B200]["scan_ms"],
                                            "std_exclusive_scan": out[A.CUDA]["std_exclusive_scan_ms"] / out[A.B200]["std_exclusive_scan_ms"]},
                "speedup_vs_CUB": {"reduce": out[A.CUB]["reduce_ms"] / out[A.B200]["reduce_ms"], "scan": out[A.CUB]["scan_ms"] / out[A.B200]["scan_ms"]}}
        return api, comp

    # ---- the other BASELINE.json configs at this N ------------------------------------------------------------------
    def configs(self):
        torch, np, sp, A = self.torch, self.np, self.sp, getattr(self, "A", None)
        w, rank, dev = self.world, self.rank, self.dev
        res = {}

        def entry(name, nbytes_per_gpu, ms, extra=None, arms_ms=None):
            e = {"ms": ms, "GBs_per_gpu": nbytes_per_gpu / ms / 1e6, "GBs_total": nbytes_per_gpu * w / ms / 1e6,
                 "frac_of_measured_peak": nbytes_per_gpu / ms / 1e6 / self.peak, "algorithmic_bytes_per_gpu": nbytes_per_gpu}
            if extra:
                e.update(extra)
            if arms_ms:
                for k, v in arms_ms.items():
                    e[k] = {"ms": v, "GBs_per_gpu": nbytes_per_gpu / v / 1e6}
            res[name] = e

        # ---- C2: benchmarks/stream copy / triad, 2^28 doubles per GPU (contiguous shards, no exchange)
        n = 1 << 28
        a = torch.empty(n, dtype=torch.float64, device=dev)
        idx = torch.arange(rank * n, (rank + 1) * n, dtype=torch.int64, device=dev)
        b = (self.hashed(idx) % 1000).double() * 0.001
        c = (self.hashed(idx + 12345) % 777).double() * 0.37
        del idx
        va, vb, vc = self.wrap(a), self.wrap(b), self.wrap(c)
        ms = self.timed(lambda: sp.stream_copy(vb, va))
        assert torch.equal(a, b), "stream copy mismatch"
        am = None
        if self.arms:
            am = {}
            for arm, nm in ((A.B200, "kokkos_api_B200"), (A.CUDA, "Kokkos::Cuda")):
                a.zero_()
                am[nm] = self.timed(lambda: self.arms.stream_copy(arm, b.data_ptr(), a.data_ptr(), n))
                assert torch.equal(a, b), f"stream copy mismatch ({nm})"
        entry("C2_stream_copy_2^28", 16.0 * n, ms, arms_ms=am)
        exp = b + 3.0 * c  # torch: separately rounded multiply and add, as the OpenMP reference without contraction
        ms = self.timed(lambda: sp.stream_triad(va, vb, vc, 3.0))
        assert torch.equal(a, exp), "stream triad mismatch"
        am = None
        if self.arms:
            am = {}
            for arm, nm in ((A.B200, "kokkos_api_B200"), (A.CUDA, "Kokkos::Cuda")):
                a.zero_()
                am[nm] = self.timed(lambda: self.arms.stream_triad(arm, a.data_ptr(), b.data_ptr(), c.data_ptr(), 3.0, n))
                assert torch.equal(a, exp), f"stream triad mismatch ({nm})"
        entry("C2_stream_triad_2^28", 24.0 * n, ms, arms_ms=am)
        del a, b, c, exp, va, vb, vc

        # ---- C4: MDRangePolicy<Rank<3>> 7-point stencil + MinMaxLoc, 512^3 per GPU, k-slabs of a 512 x 512 x (512 N) field
        n0 = n1 = 512
        nk = 512
        n2g = nk * w
        k_lo = max(0, rank * nk - 1)
        k_hi = min(n2g, (rank + 1) * nk + 1)
        n2l = k_hi - k_lo
        ii = torch.arange(n0, dtype=torch.float64, device=dev)[None, None, :]
        jj = torch.arange(n1, dtype=torch.float64, device=dev)[None, :, None]
        kk = torch.arange(k_lo, k_hi, dtype=torch.float64, device=dev)[:, None, None]
        # memory order (k, j, i) of a contiguous torch tensor == LayoutLeft (i fastest) of the Kokkos View
        u = (torch.sin(0.011 * ii + 0.3) * torch.cos(0.017 * jj) + 0.5 * torch.sin(0.013 * kk + 0.1 * torch.sin(0.02 * ii))).contiguous()
        # one planted unique maximum and minimum in the global field (owned by the first / last rank)
        pmax, pmin = (37, 211, 5), (400, 17, n2g - 7)
        for (pi, pj, pk), val in ((pmax, 64.0), (pmin, -64.0)):
            if k_lo <= pk < k_hi:
                u[pk - k_lo, pj, pi] = val
        c0, c1 = 0.5, 0.125
        s = u[1:-1, 1:-1, :-2] + u[1:-1, 1:-1, 2:]
        s = s + u[1:-1, :-2, 1:-1]
        s = s + u[1:-1, 2:, 1:-1]
        s = s + u[:-2, 1:-1, 1:-1]
        s = s + u[2:, 1:-1, 1:-1]
        v = c0 * u[1:-1, 1:-1, 1:-1] + c1 * s
        del s
        vmin, vmax = v.min(), v.max()

        def gloc(flat_index):  # flat index into v (k, j, i interior) -> global location (i*n1 + j)*n2 + k
            kq, rem = divmod(int(flat_index), (n1 - 2) * (n0 - 2))
            jq, iq = divmod(rem, n0 - 2)
            return ((iq + 1) * n1 + (jq + 1)) * n2g + (kq + 1 + k_lo)
        mine = torch.tensor([float(vmin), float(vmax)], dtype=torch.float64, device=dev)
        locs = torch.tensor([gloc(v.argmin()), gloc(v.argmax())], dtype=torch.int64, device=dev)
        del v
        if self.distributed:
            allv = torch.empty(2 * w, dtype=torch.float64, device=dev); alll = torch.empty(2 * w, dtype=torch.int64, device=dev)
            self.dist.all_gather_into_tensor(allv, mine); self.dist.all_gather_into_tensor(alll, locs)
            allv, alll = allv.view(w, 2), alll.view(w, 2)
            qmin, qmax = int(allv[:, 0].argmin()), int(allv[:, 1].argmax())  # planted extrema are unique
            exp4 = (float(allv[qmin, 0]), float(allv[qmax, 1]), int(alll[qmin, 0]), int(alll[qmax, 1]))
        else:
            exp4 = (float(mine[0]), float(mine[1]), int(locs[0]), int(locs[1]))
        vu = self.wrap(u.view(-1))
        out = torch.zeros(4, dtype=torch.float64, device=dev)
        ms = self.timed(lambda: sp.stencil7_minmaxloc_async(vu, n0, n1, n2l, n2g, k_lo, c0, c1, out))
        h = out.cpu()
        got4 = (float(h[0]), float(h[1]), int(h[2:].view(torch.int64)[0]), int(h[2:].view(torch.int64)[1]))
        assert got4 == exp4, ("stencil MinMaxLoc mismatch", got4, exp4)
        am = None
        if self.arms and not self.distributed:
            am = {}
            for arm, nm in ((A.B200, "kokkos_api_B200"), (A.CUDA, "Kokkos::Cuda")):
                out.zero_()
                am[nm] = self.timed(lambda: self.arms.stencil7_minmaxloc(arm, u.data_ptr(), n0, n1, n2l, c0, c1, out.data_ptr()))
                h = out.cpu()
                g4 = (float(h[0]), float(h[1]), int(h[2:].view(torch.int64)[0]), int(h[2:].view(torch.int64)[1]))
                assert g4 == exp4, (f"stencil MinMaxLoc mismatch ({nm})", g4, exp4)
        entry("C4_stencil7_minmaxloc_512^3", 8.0 * n0 * n1 * n2l, ms, {"result": {"min": exp4[0], "max": exp4[1], "min_loc": exp4[2], "max_loc": exp4[3]},
                                                                       "sharding": "k-slabs with one halo plane per neighbour; rank partials joined on the device (b200_allreduce_minmaxloc_f64)" if w > 1 else "single GPU"}, am)
        del u, vu

        # ---- C5a: GUPS atomic_add, 2^30-entry int64 table PER GPU (index-range shards of a 2^30 N table), 2^26 updates per GPU
        tl, m = 1 << 30, 1 << 26
        table = torch.zeros(tl, dtype=torch.int64, device=dev)
        gi = torch.arange(rank * m, (rank + 1) * m, dtype=torch.int64, device=dev)
        # counter-based generator: the 32-bit mixer applied twice and folded to 30 bits (uniform over the shard)
        idx = (self.hashed(gi) ^ (self.hashed(gi + 0x51ED27) << 17)) & (tl - 1)
        del gi
        vt, vi = self.wrap(table), self.wrap(idx)
        calls = self.warmup + self.steps
        ms = self.timed(lambda: sp.gups(vt, vi, 3))
        exp_t = torch.zeros(tl, dtype=torch.int64, device=dev)
        exp_t.index_add_(0, idx, torch.full((m,), 3 * calls, dtype=torch.int64, device=dev))
        assert torch.equal(table, exp_t), "GUPS table mismatch"
        am = None
        if self.arms:
            am = {}
            for arm, nm in ((A.B200, "kokkos_api_B200"), (A.CUDA, "Kokkos::Cuda")):
                table.zero_()
                am[nm] = self.timed(lambda: self.arms.gups_add(arm, table.data_ptr(), tl, idx.data_ptr(), m, 3))
                assert torch.equal(table, exp_t), f"GUPS table mismatch ({nm})"
        entry("C5a_gups_atomic_add_table2^30_M2^26", 16.0 * m, ms, {"GUPS_per_gpu": m / ms / 1e6, "GUPS_total": m * w / ms / 1e6,
                                                                     "sector_model_GBs_per_gpu": 72.0 * m / ms / 1e6,
                                                                     "sharding": "table sharded by index range, updates generated per owner: no exchange"}, am)
        if am:
            for nm, v_ in am.items():
                res["C5a_gups_atomic_add_table2^30_M2^26"][nm]["GUPS_per_gpu"] = m / v_ / 1e6
        del table, exp_t, idx, vt, vi
        torch.cuda.empty_cache()

        # ---- C5b: TeamPolicy nested-reduce CRS SpMV, 2^22 rows x 32 nnz PER GPU (row shards; x replicated)
        R, K = 1 << 22, 32
        nnz = R * K
        ncols = R * w
        row_map = torch.arange(0, nnz + 1, K, dtype=torch.int64, device=dev)
        e = torch.arange(rank * nnz, (rank + 1) * nnz, dtype=torch.int64, device=dev)
        rows_g = torch.div(e, K, rounding_mode="floor")
        q = e % K
        band = (rows_g + q - K // 4) % ncols
        rnd = (self.hashed(e) ^ (self.hashed(e + 99991) << 13)) % ncols
        col = torch.where(q < K // 2, band, rnd).to(torch.int32)
        val = ((self.hashed(e + 7) % 17) - 8).double()      # integer-valued: any summation order gives the same bits
        del e, rows_g, q, band, rnd
        xg = ((self.hashed(torch.arange(ncols, dtype=torch.int64, device=dev) + 3) % 13) - 6).double()
        y = torch.zeros(R, dtype=torch.float64, device=dev)
        exp_y = (val * xg[col.long()]).view(R, K).sum(1)
        nbytes = nnz * 12.0 + R * 16.0 + 8.0 * R
        vrm, vcol, vval, vx, vy = self.wrap(row_map), self.wrap(col), self.wrap(val), self.wrap(xg), self.wrap(y)
        ms = self.timed(lambda: sp.spmv_rows(vrm, vcol, vval, vx, vy))
        assert torch.equal(y, exp_y), "SpMV mismatch"
        am = None
        if self.arms:
            am = {}
            for arm, nm in ((A.B200, "kokkos_api_B200"), (A.CUDA, "Kokkos::Cuda")):
                y.zero_()
                am[nm] = self.timed(lambda: self.arms.spmv(arm, R, row_map.data_ptr(), col.data_ptr(), val.data_ptr(), xg.data_ptr(), y.data_ptr(), nnz, ncols))
                assert torch.equal(y, exp_y), f"SpMV mismatch ({nm})"
        entry("C5b_spmv_crs_2^22rows_x32", nbytes, ms, {"sharding": "rows sharded, x replicated: no exchange"}, am)
        del row_map, col, val, xg, y, exp_y

        # ---- launch latency (benchmarks/launch_latency): Kokkos user calls, host clock, 1000 launches per figure
        if self.arms and rank == 0:
            scratch = torch.zeros((1 << 16) + 8, dtype=torch.float64, device=dev)
            lat = {}
            for arm, nm in ((A.B200, "Kokkos::B200"), (A.CUDA, "Kokkos::Cuda")):
                lat[nm] = {f"{opn} n={n}": round(self.arms.launch_latency_us(arm, op, n, 1000, scratch.data_ptr()), 2)
                           for op, opn in ((0, "parallel_for"), (1, "parallel_reduce->View"), (2, "parallel_reduce->scalar")) for n in (1, 1 << 16)}
            res["launch_latency_us"] = lat
        return res

    # ---- e2e: the same step through the host-buffer C-ABI calls -------------------------------------------------------
    def bind_to_gpu_numa_node(self):
        """N > 1 only: run this rank's host threads on the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned buffers of the
        e2e leg are allocated, so that first-touch places them in that node's memory (otherwise all ranks pin node-0 memory and the
        copies of the far GPUs cross the socket link: VERDICT r1 weak 8).  Best effort: returns the node, or None when the topology
        is not visible (containers) or KB200_NUMA=0."""
        try:
            if os.environ.get("KB200_NUMA", "1") == "0":
                return None
            pr = self.torch.cuda.get_device_properties(self.local_rank)
            bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
                node = int(f.read().strip())
            if node < 0:
                return None
            cpus = set()
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                for part in f.read().strip().split(","):
                    lo, _, hi = part.partition("-")
                    cpus.update(range(int(lo), int(hi or lo) + 1))
            allowed = os.sched_getaffinity(0) & cpus
            if len(allowed) < 2:  # the leg runs two host threads
                return None
            os.sched_setaffinity(0, allowed)
            return node
        except Exception:
            return None

    def e2e(self):
        torch = self.torch
        n = self.n
        nl = min(self.n_local, n)
        numa_node = self.bind_to_gpu_numa_node() if self.world > 1 else None
        try:
            hx = torch.empty(n, dtype=torch.float64, pin_memory=True)
            hi = torch.empty(nl, dtype=torch.int64, pin_memory=True)
            hy = torch.empty(nl, dtype=torch.int64, pin_memory=True)
            hx.copy_(self.xd); hi.copy_(self.xi[:nl])
            torch.cuda.synchronize(self.dev)
            e2e_steps = max(2, min(self.steps, 3))

            # the two host-buffer calls of a step run on two execution-space instances from two host threads: the scan's
            # D2H traffic overlaps the reduce's H2D traffic (PCIe is full duplex); each call is blocking, as a scalar-result
            # parallel_reduce / a parallel_scan followed by deep_copy-to-host is in the reference
            from concurrent.futures import ThreadPoolExecutor
            space2 = self.kb.B200(self.local_rank)
            pool = ThreadPoolExecutor(2)

            def e2e_step():
                fr = pool.submit(space2.parallel_reduce_sum_host, hx.data_ptr(), n)
                ft = pool.submit(self.space.parallel_scan_host, hi.data_ptr(), hy.data_ptr(), nl, 0)
                return fr.result(), ft.result()
            e2e_step()  # warm-up (allocates the staging buffers)
            self.sync_all()
            w0 = time.perf_counter()
            for _ in range(e2e_steps):
                r, t = e2e_step()
            torch.cuda.synchronize(self.dev)
            wall = self.max_over_ranks(time.perf_counter() - w0)
            xs = self.xi[:nl]
            assert r == float(self.xd.sum().item()) and t == int(xs.sum().item())
            assert torch.equal(hy[: 1 << 20].to(self.dev), torch.cumsum(xs[: 1 << 20], 0) - xs[: 1 << 20])
            return {"value": (8.0 * n + 16.0 * nl) * self.world / (wall / e2e_steps) / 1e9, "unit": "GB/s",
                    "h2d_bytes_per_step": 8 * n + 8 * nl, "d2h_bytes_per_step": 8 * nl + 16, "steps": e2e_steps,
                    "api": "b200_reduce_sum_f64_host + b200_scan_excl_i64_host (pinned host buffers, 64 MiB chunks, double-buffered), issued from two host "
                           "threads on two instances so that H2D and D2H overlap",
                    "pcie_bound_GBs_per_gpu": "24 B of metric per 16 B of H2D: <= 1.5 x the H2D rate (~55 GB/s) = ~83",
                    "numa_node": numa_node,
                    "note": "per-GPU shards are independent on this leg (no cross-GPU seed): PCIe-bound; at N > 1 each rank's host threads and pinned "
                            "buffers are bound to its GPU's NUMA node (numa_node, null = topology not visible)"}
        except Exception as ex:  # e.g. not enough pinnable host memory
            return {"value": None, "unit": "GB/s", "error": repr(ex)[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=LOG2N_DEFAULT, help="elements per GPU and per View = 2^log2n")
    ap.add_argument("--overlap", type=int, default=-1, help="1: reduce and scan of a step run concurrently on two execution-space instances; 0: back to back; "
                                                            "-1 (default): both are timed for 3 steps after the warm-up and the faster one is used")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C2/C4/C5 legs")
    ap.add_argument("--no-arms", action="store_true", help="skip the Kokkos-user-code / comparator legs")
    args = ap.parse_args()
    rank = env_int("RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank)
        return

    B = Bench(args)
    B.make_headline()
    B.headline()
    roofline = B.roofline()
    api = comp = cfgs = None
    api, comp = B.kokkos_api_and_comparators()
    if not args.no_configs:
        cfgs = B.configs()
    e2e = None if args.no_e2e else B.e2e()
    cpu_baseline = None
    if B.rank == 0 and B.world == 1 and not args.no_cpu:
        # a separate process: the reference's OpenMP build and the CUDA build inside benchlib are two copies of Kokkos, and
        # the OpenMP thread count must not be inherited from anything this process did
        try:
            env = {k: v for k, v in os.environ.items() if not k.startswith("OMP_")}
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "5", "--warmup", "1"],
                               capture_output=True, text=True, timeout=600, env=env)
            cpu_baseline = json.loads(p.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as ex:
            cpu_baseline = {"value": None, "error": repr(ex)[:200]}

    if B.rank == 0:
        world, n = B.world, B.n
        par = f"x{world}: reduce View sharded contiguously + NCCL all-reduce; scan View block-cyclic (block {B.block} elements), one fused kernel per rank, " \
              "round aggregates over peer-mapped NVLink mailboxes (b200_comm_scan_excl_i64)" if B.distributed else "single GPU"
        line = {"metric": "parallel_reduce+parallel_scan HBM throughput", "value": B.value, "unit": "GB/s", "n_gpus": world,
                "steps": args.steps, "warmup": B.warmup, "ms_per_step": B.ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64+int64", "data": "synthetic",
                "config": {"workload": workload_text(args.log2n), "policy": "RangePolicy", "elements_per_gpu_per_view": n,
                           "scan_elements_on_rank0": B.n_local, "algorithmic_bytes_per_step_per_gpu": 24 * n, "parallelism": par,
                           "step": ("the reduce and the scan of a step are issued on two execution-space instances (two streams) and run concurrently; "
                                    "a step ends when both are done" if B.overlap else "reduce then scan, one stream") +
                                   (" (chosen by timing 3 steps of each structure after the warm-up)" if args.overlap < 0 else ""),
                           "api": "typed C-ABI entry points (b200_reduce_sum_f64, b200_scan_excl_i64 / b200_comm_scan_excl_i64): the same calls at every N; "
                                  "the Kokkos-lambda form of the same step is `kokkos_api`",
                           "l2": "inputs (8 GiB per View) are far larger than the 126 MB L2; no flush needed",
                           "timing": "CUDA events on the launching stream, barrier+synchronize both sides, max over ranks",
                           "checks": "reduce == torch sum; EVERY scan output element == torch cumsum with gathered block offsets (bit-exact), before the timed region"},
                "frac_of_measured_peak": B.value / (B.peak * world), "frac_of_8TBs_nominal": B.value / (8000.0 * world),
                "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": B.launches_per_step * args.steps,
                "clocks": B.clocks, "kokkos_api": api, "comparators": comp,
                "breakdown": {"scan_GBs_per_gpu": roofline["achieved"], "scan_ms": B.scan_ms, "scan_ms_inside_step": B.scan_ms_in_step,
                              "reduce_plus_collectives_ms": (B.reduce_ms_alone if B.overlap else B.ms_per_step - B.scan_ms),
                              "reduce_GBs_per_gpu": 8.0 * n / ((B.reduce_ms_alone if B.overlap else B.ms_per_step - B.scan_ms) * 1e-3) / 1e9,
                              "step_structure": ("reduce (+ NCCL all-reduce) on instance B || scan on instance A, fork/join by events every step" if B.overlap
                                                 else "reduce then scan on one instance"),
                              "step_structure_probe_ms": getattr(B, "overlap_probe_ms", None),
                              "configs": cfgs}}
        if getattr(B, "arms_error", None):
            line["arms_error"] = B.arms_error
        print(json.dumps(line), flush=True)
    if B.arms is not None:
        B.torch.cuda.synchronize(B.dev)
        B.arms.finalize()
    if B.distributed:
        B.dist.barrier()
        if B.comm is not None:
            B.comm.finalize()
        B.dist.destroy_process_group()


if __name__ == "__main__":
    main()
