#!/usr/bin/env python
"""Condenses an .ncu-rep (read here, no GPU needed) into one compact table: one row per profiled launch.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__grid_size", "grid"), ("launch__block_size", "blk"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel")]


def to_unit(val, unit, short):
    try:
        v = float(val.replace(",", ""))
    except ValueError:
        return val
    if short in ("rdMB", "wrMB"):
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
        return "%.2f" % (v * scale)
    if short == "us":
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        return "%.1f" % (v * scale)
    return "%.2f" % v if v != int(v) else "%d" % v


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(s for _, s in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("<unnamed>::", "")
        cells = [to_unit(r[idx[m]], units[idx[m]], s) if m in idx else "-" for m, s in COLS]
        print("| " + name + " | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
