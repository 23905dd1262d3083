#!/usr/bin/env python
"""Per-kernel summary of an ncu report (run where `ncu` is on PATH): one CSV line per captured launch with its
duration, DRAM bytes, achieved DRAM bandwidth, issue-slot and FP64-pipe utilisation, occupancy, registers.

    python scripts/ncu_summary.py report.ncu-rep > profiles/<name>.csv
"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__sass_average_branch_targets_threads_uniform.pct"]


def main():
    txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    kn = col.get("Kernel Name")
    out = csv.writer(sys.stdout)
    names = [w for w in WANT if w in col]
    out.writerow(["kernel", "dram_GBps"] + ["%s [%s]" % (w, units[col[w]]) for w in names])
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    tscale = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}
    for r in rows[2:]:
        try:
            by = sum(float(r[col[k]].replace(",", "")) * scale.get(units[col[k]], 1.0)
                     for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            t = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * tscale.get(units[col["gpu__time_duration.sum"]], 1.0)
            bw = "%.1f" % (by / t * 1e-9)
        except Exception:
            bw = ""
        out.writerow([r[kn][:70], bw] + [r[col[w]] for w in names])


if __name__ == "__main__":
    main()
