#!/bin/bash
# Turns the files a `tools/gpu_check.sh; tools/gpu_profile.sh <tag> --config C5` call left in gpurun_out/ into the committed
# evidence of profiles/<round>/: summaries, function / line views, launch list, roofline_ncu.json.
#   tools/collect_profiles.sh <tag> <round dir>      e.g. tools/collect_profiles.sh r02g profiles/r02
TAG=$1; OUT=$2
for k in k_lidar_obs k_features k_traffic k_ego; do python tools/ncu_summary.py gpurun_out/prof_${k}_$TAG.ncu-rep --json C5:65536 > $OUT/ncu_${k}_final_65536envs_summary.txt; done
python tools/ncu_groups.py gpurun_out/prof_k_lidar_obs_$TAG.ncu-rep k_lidar_obs k_lidar_obsILi72ELb1 > $OUT/ncu_k_lidar_obs_final_65536envs_functions.txt
python tools/ncu_lines.py gpurun_out/prof_k_lidar_obs_$TAG.ncu-rep k_lidar_obs 60 k_lidar_obsILi72ELb1 > $OUT/ncu_k_lidar_obs_final_65536envs_lines.txt 2>/dev/null
python tools/ncu_groups.py gpurun_out/prof_k_features_$TAG.ncu-rep k_features > $OUT/ncu_k_features_final_65536envs_functions.txt
python tools/ncu_groups.py gpurun_out/prof_k_traffic_$TAG.ncu-rep k_traffic k_trafficILi8 > $OUT/ncu_k_traffic_final_65536envs_functions.txt
python tools/ncu_groups.py gpurun_out/prof_k_ego_$TAG.ncu-rep k_ego k_egoILi8 > $OUT/ncu_k_ego_final_65536envs_functions.txt
cp gpurun_out/launches_$TAG.csv $OUT/ncu_launches_final_65536envs.csv
python - "$TAG" "$OUT" <<'PY'
import json, csv, collections, re, sys
tag, out = sys.argv[1], sys.argv[2]
line = [l for l in open(f'{out}/ncu_k_lidar_obs_final_65536envs_summary.txt') if l.startswith('JSON ')][0][5:]
d = json.loads(line)
d["source"] = ("one `ncu --set full --clock-control none --import-source on` capture of k_lidar_obs<72> at step 410 of `bench.py --config C5` "
               f"(65,536 envs x 8 agents x 72 beams), final build: {out}/ncu_k_lidar_obs_final_65536envs_summary.txt; regenerate with "
               "tools/gpu_profile.sh + tools/collect_profiles.sh")
json.dump(d, open('profiles/roofline_ncu.json', 'w'), indent=1)
rows = list(csv.reader(l for l in open(f'gpurun_out/launches_{tag}.csv') if l.startswith('"')))
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
t = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[iv].replace(",", "")); u = r[iu]
    t[re.sub(r"\(.*", "", r[ik]).replace("void ", "")].append(v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v))
step = {k: v for k, v in t.items() if "isx::" in k}
tot = sum(sum(v) / len(v) for v in step.values())
o = ["ncu --metrics gpu__time_duration.sum --clock-control none -s 2060 -c 192 : python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config C5   (65536 envs x 8 agents, steady state after the 400-step pre-roll; final build)",
     "(the at::FillFunctor launches are bench.py's L2 flush between timed steps, outside the event pairs; per-launch times under ncu are serialised and cold-cache: compare SHARES)"]
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1]) / len(kv[1])):
    m = sum(v) / len(v)
    o.append(f"{k:60s} n={len(v):4d} mean {m:9.2f} us" + (f"   {100 * m / tot:5.1f}% of the step kernels" if k in step else ""))
open(f'{out}/ncu_launches_final_65536envs_summary.txt', 'w').write("\n".join(o) + "\n")
print("\n".join(o[2:7]))
PY
