#!/usr/bin/env python
"""Markdown table + traffic json of every launch in an .ncu-rep (read here, no GPU)."""
import csv, io, json, subprocess, sys
COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"), ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid")]
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
def main(path, workload=None, out_json=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(n for _, n in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    traffic = {}
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").strip()
        cells = []
        for m, _ in COLS:
            v, u = r[col[m]], units[col[m]]
            try: f = float(v.replace(",", ""))
            except ValueError: cells.append(v); continue
            cells.append(f"{f:.4g} {u}".strip() if u not in ("%",) else f"{f:.1f}")
        print(f"| `{name}` | " + " | ".join(cells) + " |")
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        traffic[name] = {"dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr)}
    if out_json:
        with open(out_json, "w") as f:
            json.dump({"workload": workload, "source": path, "kernels": traffic}, f, indent=1)
if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else None)
