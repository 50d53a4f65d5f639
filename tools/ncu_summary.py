#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` exports: one block of key metrics per captured launch."""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed_op_shared_atom.sum", "shared atomics"),
    ("sm__cycles_elapsed.avg", "sm cycles"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"### {name[:110]}")
        for key, label in WANT:
            if key in hdr:
                i = hdr.index(key)
                print(f"  {label:24s} {r[i]:>18s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("per_warp_active.pct") is False and "_not_issued" not in h and h.endswith(".ratio"):
                try:
                    stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            print("  top stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in stalls[:5]))
        print()


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
