"""Compact per-kernel summary of an `ncu --set full` report (read here, no GPU needed).

usage: ncu_summary.py report.ncu-rep [report2.ncu-rep ...] > profiles/rNN_xxx.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block", "smem/blk"),
]


def fmt(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    if x >= 1e6 and not unit:
        return f"{x:.3e}"
    return f"{x:.4g}" + (f" {unit}" if unit else "")


def main():
    print("| kernel | " + " | ".join(k[1] for k in KEYS) + " |")
    print("|---|" + "---|" * len(KEYS))
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        seen = set()
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
            if name in seen:
                continue
            seen.add(name)
            cells = []
            for k, _ in KEYS:
                cells.append(fmt(r[idx[k]], units[idx[k]]) if k in idx else "-")
            print(f"| {name} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
