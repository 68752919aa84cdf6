#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one line per profiled launch with the counters the
roofline argument needs (duration, DRAM bytes and %, FP64 pipe %, occupancy, registers, L1/L2 hit)."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur_us", 1e-3),
    ("dram__bytes_read.sum", "rd_MB", 1e-6),
    ("dram__bytes_write.sum", "wr_MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%", 1),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64c%", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("l1tex__t_sector_hit_rate.pct", "l1hit%", 1),
    ("lts__t_sector_hit_rate.pct", "l2hit%", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%", 1),
    ("smsp__inst_executed.sum", "inst_M", 1e-6),
    ("local_load", "lmem", 1),
]


def main(path):
    rows = list(csv.reader(open(path)))
    # header row = first row whose first cell is "ID"
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[hi], rows[hi + 1]
    col = {n: i for i, n in enumerate(hdr)}
    out = []
    names = [k for k, _, _ in KEYS if k in col]
    print("kernel".ljust(44), " ".join(lbl.rjust(8) for k, lbl, _ in KEYS if k in col))
    for r in rows[hi + 2:]:
        if len(r) < len(hdr):
            continue
        kn = r[col["Kernel Name"]].split("(")[0][-42:]
        vals = []
        for k, lbl, sc in KEYS:
            if k not in col:
                continue
            try:
                v = float(r[col[k]].replace(",", ""))
                u = units[col[k]]
                if lbl == "dur_us":
                    v = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
                elif lbl in ("rd_MB", "wr_MB"):
                    v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1e-6)
                elif lbl == "inst_M":
                    v = v * 1e-6
                vals.append(f"{v:8.1f}")
            except ValueError:
                vals.append(r[col[k]][:8].rjust(8))
        print(kn.ljust(44), " ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
