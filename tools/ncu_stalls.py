#!/usr/bin/env python
"""Kernel-wide warp-stall reason shares from the source page of an .ncu-rep (no GPU needed)."""
import csv, io, subprocess, sys, collections
def main(path, kernel):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kernel}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None; tot = collections.Counter()
    for r in rows:
        if hdr is None:
            if "Instructions Executed" in r: hdr = r
            continue
        if len(r) < len(hdr): continue
        for i, h in enumerate(hdr):
            if h.startswith("stall_") and "Not Issued" not in h:
                try: tot[h] += int(r[i])
                except ValueError: pass
    s = sum(tot.values()) or 1
    print(kernel, " ".join(f"{k[6:]}={100*v/s:.1f}%" for k, v in tot.most_common(12)))
if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
