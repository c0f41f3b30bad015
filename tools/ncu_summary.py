"""tools/ncu_summary.py <report.ncu-rep> [kernel-substring] -- condensed per-kernel summary of an ncu report (run here, no GPU):
duration, DRAM bytes/throughput, pipe utilisation, issue rate, occupancy, top warp-stall reasons.  Output is what gets
committed under profiles/."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("time_us", "gpu__time_duration.sum"),
    ("dram_read_GB", "dram__bytes_read.sum"),
    ("dram_write_GB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1tex_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue_pct", "sm__inst_executed.sum.pct_of_peak_sustained_elapsed"),
    ("ipc", "sm__inst_executed.avg.per_cycle_active"),
    ("fp64_pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    ("alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("fma_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("smem_wavefronts_pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("dyn_smem", "launch__shared_mem_per_block_dynamic"),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
]


def main():
    rep = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    col = {k: i for i, k in enumerate(h)}
    for row in rows[2:]:
        name = row[col["Kernel Name"]]
        if filt and filt not in name:
            continue
        print("==", name[:150])
        for label, key in KEYS:
            if key in col:
                print(f"   {label:22s} {row[col[key]]:>16s} {units[col[key]]}")
        stalls = []
        for k, i in col.items():
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") or \
               (k.startswith("smsp__average_warp_latency_issue_stalled_") and k.endswith(".ratio")):
                try:
                    stalls.append((float(row[i]), k.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "")
                                   .replace("_per_issue_active.ratio", "").replace(".ratio", "")))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("   stalls (warp-cycles per issued instruction):", ", ".join(f"{n}={v:.2f}" for v, n in stalls[:8]))


if __name__ == "__main__":
    main()
