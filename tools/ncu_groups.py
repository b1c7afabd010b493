#!/usr/bin/env python
"""Grouped (per function) view on top of tools/ncu_lines.py:  python tools/ncu_groups.py <rep> <kernel> [mangled]"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern = sys.argv[1], sys.argv[2]
mangled = sys.argv[3] if len(sys.argv) > 3 else kern
txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, kern, "100000", mangled], stdout=subprocess.PIPE, text=True).stdout
def marks(path, names):
    src = open(path).read().splitlines()
    out = []
    for label, needle in names:
        for i, l in enumerate(src):
            if needle in l:
                out.append((label, i + 1)); break
    return sorted(out, key=lambda x: x[1])
C = os.path.join(ROOT, "marl-traffic-intersection_b200", "csrc")
sim = marks(os.path.join(C, "isx_sim.cuh"), [("on_road/geom", "ISX_HD bool on_road"), ("car_update", "ISX_HD void car_update"), ("corners/sat", "ISX_HD void car_corners"),
    ("path_index_update", "ISX_HD int path_index_update"), ("ego_self_status", "ISX_HD int ego_self_status"), ("reward_base", "ISX_HD float reward_base"),
    ("car_pixel_rect", "ISX_HD PixRect car_pixel_rect"), ("ray_pixel", "ISX_HD void ray_pixel"), ("road_bit/skip", "ISX_HD bool road_bit"),
    ("rcp/make_ray/axis_exit", "ISX_HD float approx_rcp"), ("march_init", "ISX_HD void march_init"), ("sample_event", "ISX_HD int sample_event"),
    ("march_step", "ISX_HD void march_step"), ("ray_road_event", "ISX_HD int ray_road_event"), ("ray_rect_first_hit", "ISX_HD int ray_rect_first_hit"),
    ("beam_window", "ISX_HD BeamWindow beam_window"), ("npc_pair_flags", "ISX_HD int npc_pair_flags"), ("npc_front/steer/throttle", "ISX_HD float npc_front_candidate"),
    ("obs features", "ISX_HD void obs_ego_features")])
ker = marks(os.path.join(C, "isx_kernels.cu"), [("warp helpers", "warp_min_f"), ("dyn: load/reset", "k_dynamics(const Dev d"), ("dyn: traffic load+spawn", "traffic flow (TrafficFlow.cpp:317-367)"),
    ("dyn: npc loop", "NPC controller, sequential in list order"), ("dyn: npc collisions", "NPC-NPC collisions (:347-356)"), ("dyn: npc erase", "ordered erase of dead"),
    ("dyn: ego update+pathidx", "egos (IntersectionEnv.cpp:144-370)"), ("dyn: ego car-car", "car-car override (:293-318)"), ("dyn: bonuses..term", "terminal bonuses (:321-326)"),
    ("dyn: writeback+stats", "---- write back"), ("lid: warp_road_event", "int warp_road_event("), ("lid: setup/stage", "k_lidar_obs(const Dev d"),
    ("lid: features", "per ego: candidate set"), ("lid: beams", "beams: one thread per (ego, beam)"), ("small kernels", "small kernels")])
tot, total = {}, 0
for ln in txt.splitlines():
    m = re.match(r"(\S+):(\d+)\s+([\d,]+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)", ln)
    if not m: continue
    f, l, wi, thr, samp = m.group(1), int(m.group(2)), int(m.group(3).replace(",", "")), float(m.group(5)), float(m.group(6))
    total += wi
    key = f
    for name, table in (("isx_sim.cuh", sim), ("isx_kernels.cu", ker)):
        if f == name:
            key = name + ":?"
            for label, start in table:
                if l >= start: key = label
    t = tot.setdefault(key, [0, 0.0, 0.0]); t[0] += wi; t[1] += wi * thr; t[2] += samp
print(txt.splitlines()[0])
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:32s} {v[0]:12,d} {100 * v[0] / total:5.1f}%  thr/inst {v[1] / max(v[0], 1):5.1f}  samples {v[2]:5.1f}%")
