#!/usr/bin/env python
"""Executed warp-instruction histogram by SASS opcode for one kernel of an .ncu-rep (no GPU needed)."""
import csv, io, subprocess, sys, collections
def main(path, kernel, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kernel}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None; tot = collections.Counter()
    for r in rows:
        if hdr is None:
            if "Instructions Executed" in r:
                hdr = r; ci = hdr.index("Instructions Executed"); cs = hdr.index("Source")
            continue
        if len(r) < len(hdr): continue
        try: n = int(r[ci])
        except ValueError: continue
        src = r[cs].strip()
        parts = src.split()
        if not parts: continue
        op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
        tot[op.split(".")[0]] += n
    s = sum(tot.values()) or 1
    print(kernel, "total warp instructions", s)
    for k, v in tot.most_common(top): print(f"{100*v/s:5.1f}%  {v:>12}  {k}")
if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
