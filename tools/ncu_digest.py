#!/usr/bin/env python
"""Digest of an .ncu-rep (read on the CPU box): headline counters + top stall reasons + hottest source lines."""
import csv, io, subprocess, sys

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    return r[0], r[1], r[2:]

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct",
        "sm__pipe_fmaheavy", "sm__inst_executed_pipe_tc", "sm__pipe_tensor", "sm__pipe_tc", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_lsu.avg.pct", "sm__inst_executed_pipe_xu.avg.pct", "sm__inst_executed_pipe_alu.avg.pct",
        "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_uniform.avg.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "smsp__cycles_active.avg", "lts__t_sector_hit_rate.pct",
        "sm__instruction_throughput.avg.pct", "l1tex__throughput.avg.pct", "lts__throughput.avg.pct"]

def main():
    rep = sys.argv[1]
    h, u, rows = raw(rep)
    for row in rows:
        print("kernel:", row[h.index("Kernel Name")][:80])
        for i, n in enumerate(h):
            if any(n.startswith(k) for k in KEEP):
                print(f"  {n:75s} {row[i]:>16s} {u[i]}")
        st = [(float(row[i].replace(",", "") or 0), n) for i, n in enumerate(h)
              if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio")]
        if not st:
            st = [(float(row[i].replace(",", "") or 0), n) for i, n in enumerate(h)
                  if n.startswith("smsp__average_warp") and "stalled" in n]
        for v, n in sorted(st, reverse=True)[:8]:
            print(f"  stall {n:80s} {v:8.3f}")
    if len(sys.argv) > 2:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        r = list(csv.reader(io.StringIO(out)))
        hh = r[0]
        print(hh)
if __name__ == "__main__":
    main()
