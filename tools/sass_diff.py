#!/usr/bin/env python3
"""Per-kernel SASS comparison of two builds of libosp_b200.so (cuobjdump -sass): which kernels are new, gone or changed.

Used when a change must not touch kernels that were already verified on a B200 -- e.g. the tests/cusim shims and the
opt-in long-row sweep were checked this way against the last GPU-verified commit (only k_merge_xl differs: one more
comparison in its row filter):

    git archive <rev> outerspace_b200/csrc include | tar -x -C /tmp/old && (cd /tmp/old/outerspace_b200/csrc && make)
    python tools/sass_diff.py /tmp/old/outerspace_b200/libosp_b200.so outerspace_b200/libosp_b200.so
"""
import re
import subprocess
import sys


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    table, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            table[cur] = []
        elif cur is not None:
            table[cur].append(line)
    return table


def main():
    old, new = functions(sys.argv[1]), functions(sys.argv[2])
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()[:140]
    same = 0
    for name in sorted(set(old) | set(new)):
        if name not in old:
            print("NEW    ", demangle(name))
        elif name not in new:
            print("GONE   ", demangle(name))
        elif old[name] != new[name]:
            print("CHANGED", demangle(name))
        else:
            same += 1
    print(f"{same} kernels identical, {len(old)} before, {len(new)} after")


if __name__ == "__main__":
    main()
