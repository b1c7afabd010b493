#!/usr/bin/env python
"""Per-FUNCTION view of an ncu report, on top of tools/ncu_lines.py:

    python tools/ncu_groups.py <rep> <kernel> [mangled-substring]

Every profiled source line is attributed to the function whose definition encloses it in the current sources (found by
scanning upwards for a definition line at column 0), so the table follows the code as it is refactored."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.environ.get("ISX_CSRC") or os.path.join(ROOT, "marl-traffic-intersection_b200", "csrc")   # ISX_CSRC: sources of the profiled build
rep, kern = sys.argv[1], sys.argv[2]
mangled = sys.argv[3] if len(sys.argv) > 3 else kern
txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, kern, "100000", mangled],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
DEF = re.compile(r"^(?:ISX_HD\w*|__device__|__global__|template|static|inline|struct)[^;]*?\b(\w+)\s*(?:\(|\{)")
KERN = re.compile(r"^(k_\w+)\(")
sources = {}


def function_of(fname, line):
    path = os.path.join(CSRC, fname)
    if fname not in sources:
        sources[fname] = open(path, errors="ignore").read().split("\n") if os.path.exists(path) else None
    src = sources[fname]
    if src is None:
        return fname
    for i in range(min(line, len(src)) - 1, -1, -1):
        m = KERN.match(src[i]) or (DEF.match(src[i]) if not src[i].startswith(" ") else None)
        if m:
            return m.group(1)
    return fname


inst, samp, lanes = collections.Counter(), collections.Counter(), collections.Counter()
for ln in txt.splitlines():
    m = re.match(r"(\S+):(\d+)\s+([\d,]+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)", ln)
    if not m:
        continue
    f, l, c, thr, sp = m.group(1), int(m.group(2)), int(m.group(3).replace(",", "")), float(m.group(5)), float(m.group(6))
    k = function_of(f, l)
    inst[k] += c
    samp[k] += sp
    lanes[k] += c * thr
tot = sum(inst.values()) or 1
print(txt.splitlines()[0] if txt else "no data")
print(f"{'function':34s} {'%inst':>7s} {'%samples':>9s} {'lanes/inst':>10s}")
for k, v in inst.most_common(24):
    print(f"{k:34s} {100 * v / tot:7.2f} {samp[k]:9.2f} {lanes[k] / max(v, 1):10.1f}")
