#!/usr/bin/env python3
"""ncu raw page (ncu -i x.ncu-rep --page raw --csv) -> compact per-launch CSV with the metrics DESIGN.md / profiles/ cite.
usage: python tools/summarize_ncu.py raw.csv out.csv"""
import csv
import sys

KEEP = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__warps_eligible.avg.per_cycle_active",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    cols = [h for h in KEEP if h in hdr] + [h for h in hdr if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio")]
    idx = [hdr.index(h) for h in cols]
    with open(sys.argv[2], "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([h.replace(STALLS, "stall_").replace("_per_issue_active.ratio", "") for h in cols])
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for i in idx])


if __name__ == "__main__":
    main()
