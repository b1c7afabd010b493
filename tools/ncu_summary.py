#!/usr/bin/env python
"""Text summary of one ncu report (first kernel in it):

    python tools/ncu_summary.py <rep> [--json key] > profiles/rNN/ncu_<kernel>_summary.txt

Prints the launch shape, issue-slot utilisation (warp-instructions per cycle per SM sub-partition, active lanes per
instruction), pipe utilisation, DRAM / L2 / shared-memory traffic with the achieved GB/s against the B200 peaks, and the
top stall reasons.  With --json KEY it also prints one JSON object (the entry bench.py reads from
profiles/roofline_ncu.json under KEY, e.g. "C5:65536")."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.per_cycle_active",
        "smsp__inst_executed.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U, V = rows[0], rows[1], rows[2]


def num(name):
    if name not in H or not V[H.index(name)]:
        return None
    try:
        return float(V[H.index(name)].replace(",", ""))
    except ValueError:
        return None


def unit(name):
    return U[H.index(name)] if name in H else ""


for w in WANT:
    if w in H:
        i = H.index(w)
        print(f"{w:75s} {V[i]}  [{U[i]}]")
stalls = [(H[i], float(V[i].replace(",", ""))) for i in range(len(H)) if H[i].startswith("smsp__average_warps_issue_stalled") and H[i].endswith("per_issue_active.ratio") and V[i]]
for n, v in sorted(stalls, key=lambda kv: -kv[1])[:8]:
    print(f"{n:94s} {v:.2f}")

# ---- derived rates against the peaks
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0, "s": 1.0}


def scaled(name):
    v = num(name)
    return None if v is None else v * SCALE.get(unit(name), 1.0)


dur = scaled("gpu__time_duration.sum")
peaks = {}
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peaks = json.load(open(pp))
hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
print("---- derived (per launch; under ncu the launch is serialised and cold-cache, use shares not absolutes)")
out = {}
if dur:
    dr, dw = scaled("dram__bytes_read.sum") or 0.0, scaled("dram__bytes_write.sum") or 0.0
    l2 = scaled("lts__t_bytes.sum")
    if l2 is None and num("lts__t_sectors.sum") is not None:
        l2 = num("lts__t_sectors.sum") * 32.0            # a sector is 32 B
    l1 = scaled("l1tex__t_bytes.sum")
    print(f"duration                         {dur * 1e6:10.1f} us")
    print(f"DRAM bytes (read+write)          {(dr + dw) / 1e6:10.1f} MB   -> {(dr + dw) / dur / 1e9:8.1f} GB/s  = {(dr + dw) / dur / 1e9 / hbm_peak * 100:5.2f}% of the measured HBM peak ({hbm_peak:.0f} GB/s)")
    out["dram_bytes_per_launch"] = dr + dw
    out["dram_gbs"] = (dr + dw) / dur / 1e9
    if l2 is not None:
        print(f"L2 bytes (lts__t_bytes)          {l2 / 1e6:10.1f} MB   -> {l2 / dur / 1e9:8.1f} GB/s")
        out["l2_bytes_per_launch"] = l2
        out["l2_gbs"] = l2 / dur / 1e9
    if l1 is not None:
        print(f"L1/TEX bytes (l1tex__t_bytes)    {l1 / 1e6:10.1f} MB   -> {l1 / dur / 1e9:8.1f} GB/s")
    wf = num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
    cyc = num("sm__cycles_active.avg")
    if wf is not None:
        # one shared-memory wavefront moves up to 128 B (32 banks x 4 B) per SM per cycle
        print(f"shared-memory wavefronts         {wf:14.0f}   -> <= {wf * 128 / dur / 1e9:8.1f} GB/s (128 B per wavefront)"
              + (f", {wf / (cyc * 148):.3f} wavefronts/cycle/SM of a possible 1.0" if cyc else ""))
        out["smem_wavefronts_per_launch"] = wf
        if cyc:
            out["smem_wavefronts_per_cycle_per_sm"] = wf / (cyc * 148)
ipc = num("smsp__issue_active.avg.per_cycle_active")
lanes = num("smsp__thread_inst_executed_per_inst_executed.ratio")
if ipc is not None and lanes is not None:
    print(f"issue slots: {ipc:.3f} warp-inst/cycle/SMSP of 1.0, {lanes:.2f} of 32 lanes active -> {ipc * lanes / 32:.3f} of the thread-instruction issue capacity")
    out["issue"] = {"warp_inst_per_cycle_per_smsp": ipc, "ceiling": 1.0, "active_lanes_per_inst": lanes, "lane_ceiling": 32,
                    "frac_of_thread_issue_capacity": ipc * lanes / 32, "inst_executed": num("smsp__inst_executed.sum")}
if "--json" in sys.argv:
    key = sys.argv[sys.argv.index("--json") + 1]
    print("JSON " + json.dumps({key: out}))
