#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, no GPU): per launch the time, DRAM bytes, DRAM/L2/SM throughput %,
L2 hit rate, achieved occupancy, IPC, registers -- the evidence profiles/*.md cite."""
import csv, io, subprocess, sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
]

def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = r[col["Kernel Name"]][:48]
        vals = []
        for m in METRICS:
            if m in col:
                v = r[col[m]].replace(",", "")
                u = units[col[m]]
                vals.append(f"{m.split('.')[0].replace('launch__','').replace('__','.')}={v}{u if u not in ('', '%') else ''}")
        print(name, "|", "  ".join(vals))

if __name__ == "__main__":
    main(sys.argv[1])
