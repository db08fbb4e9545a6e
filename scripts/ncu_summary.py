"""Summarise an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`) per kernel launch:
duration, DRAM bytes and throughput, occupancy, issue rate, FP64 pipe, registers, and the stall cycles per issued
instruction by reason.  Usage: ncu -i rep --page raw --csv > raw.csv ; python scripts/ncu_summary.py raw.csv"""
import csv
import re
import sys

KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("smsp__inst_executed.sum", "warp instructions"), ("launch__registers_per_thread", "registers"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("launch__occupancy_limit_registers", "occ limit regs (CTAs/SM)"),
        ("launch__occupancy_limit_shared_mem", "occ limit smem (CTAs/SM)")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    stall = [(i, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for i, h in enumerate(hdr)
             if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    seen = {}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        seen[name] = seen.get(name, 0) + 1
        if seen[name] > 2:  # two launches per kernel are enough
            continue
        print(f"## {name}  (launch {seen[name]})")
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {label:28s} {r[i]} {units[i]}")
        st = sorted(((float(r[i].replace(',', '')), n) for i, n in stall if r[i]), reverse=True)[:7]
        print("  stall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in st))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
