#!/usr/bin/env python
"""Per-CUDA-source-line instruction and stall-sample shares of one kernel in an .ncu-rep
(ncu --page source --print-source cuda,sass), read here without a GPU.
usage: ncu_lines.py REP KERNEL [TOP] [inst|stall]"""
import csv, io, re, subprocess, sys, collections

def main(path, kernel, top=40, key="inst"):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", f"regex:{kernel}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    per_line = collections.defaultdict(lambda: [0, 0, ""])
    cur = None
    for r in rows:
        if hdr is None:
            if "Instructions Executed" in r:
                hdr = r
                ci, cst = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
                csrc = hdr.index("Source")
            continue
        if len(r) < len(hdr):
            continue
        src = r[csrc]
        try:
            inst, st = int(r[ci]), int(r[cst])
        except ValueError:
            continue
        first = r[0]
        if first and not first.startswith("0x"):      # a CUDA source line row: "line number"
            cur = (first, src.strip()[:100])
            continue
        if cur is not None:
            per_line[cur][0] += inst
            per_line[cur][1] += st
    tot = sum(v[0] for v in per_line.values()) or 1
    tots = sum(v[1] for v in per_line.values()) or 1
    print(f"kernel {kernel}: {tot} warp instructions, {tots} stall samples")
    idx = 0 if key == "inst" else 1
    for (line, text), (inst, st, _) in sorted(per_line.items(), key=lambda kv: -kv[1][idx])[:top]:
        print(f"{100*inst/tot:5.1f}% inst {100*st/tots:5.1f}% stall  L{line:>4} {text}")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40, sys.argv[4] if len(sys.argv) > 4 else "inst")
