#!/usr/bin/env python
"""Per-source-line view of an ncu report without a GUI (no GPU needed).

  python tools/ncu_lines.py gpurun_out/prof_lidar.ncu-rep k_lidar_obs [top_n]

Joins `ncu --page source --print-source sass --csv` (per-SASS-instruction executed counts and stall samples) with the
line table `nvdisasm -g` prints for the cubin inside libisx_b200.so, and aggregates by file:line.  Instruction order
is assumed identical between the two listings (same cubin), which is checked by opcode text."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("ISX_LIB") or os.path.join(ROOT, "marl-traffic-intersection_b200", "csrc", "libisx_b200.so")   # ISX_LIB: the profiled build


def sass_lines_with_src(kernel_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cub = [f for f in os.listdir(tmp) if f.startswith("isx_kernels.") and f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], stdout=subprocess.PIPE, text=True).stdout
    out, cur, active = [], ("?", 0), False
    for ln in txt.splitlines():
        if ln.startswith("//--------------------- .text."):
            active = kernel_substr in ln
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return out


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    mangled = sys.argv[4] if len(sys.argv) > 4 else kern      # substring of the mangled name (template instances)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    # the report may hold several launches: keep the first block for this kernel
    blocks = raw.split('"Kernel Name"')
    blk = next(b for b in blocks if kern in b.split("\n", 1)[0])
    rows = list(csv.reader(io.StringIO(blk.split("\n", 1)[1])))
    hdr = rows[0]
    ia, isrc = hdr.index("Address"), hdr.index("Source")
    iex, ith = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    isamp = hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    sass = sass_lines_with_src(mangled)
    body = [r for r in rows[1:] if len(r) > iex and r[iex] != ""]
    if len(body) != len(sass):
        print(f"warning: {len(body)} profiled instructions vs {len(sass)} disassembled", file=sys.stderr)
    per = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
    tot_ex = tot_samp = 0
    for r, s in zip(body, sass):
        key = s[2]
        ex, th, sp = int(float(r[iex])), int(float(r[ith])), int(float(r[isamp] or 0))
        p = per[key]
        p[0] += ex
        p[1] += th
        p[2] += sp
        for i, h in stall_cols:
            if r[i]:
                p[3][h] += int(float(r[i]))
        tot_ex += ex
        tot_samp += sp
    print(f"kernel {kern}: warp-instructions {tot_ex:,}  samples {tot_samp:,}")
    print(f"{'file:line':28s} {'warp-inst':>12s} {'%inst':>6s} {'thr/inst':>8s} {'%samp':>6s}  top stalls")
    for key, p in sorted(per.items(), key=lambda kv: -kv[1][0])[:topn]:
        st = sorted(p[3].items(), key=lambda kv: -kv[1])[:3]
        sts = " ".join(f"{h[6:]}={v}" for h, v in st if v)
        print(f"{key[0] + ':' + str(key[1]):28s} {p[0]:12,d} {100 * p[0] / max(tot_ex, 1):6.2f} {p[1] / max(p[0], 1):8.1f} "
              f"{100 * p[2] / max(tot_samp, 1):6.2f}  {sts}")
    # by file
    byfile = defaultdict(int)
    for key, p in per.items():
        byfile[key[0]] += p[0]
    print("by file:", {k: f"{100 * v / max(tot_ex, 1):.1f}%" for k, v in byfile.items()})


if __name__ == "__main__":
    main()
