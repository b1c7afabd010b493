#!/bin/bash
# ncu evidence for the bench workload (1 GPU) in its steady state (400 warm-up steps).  Plain run first (must exit 0),
# then the launch list, then one --set full capture of each step kernel.
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 400 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
tail -1 gpurun_out/plain.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -s 1610 -c 160 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
for k in k_lidar_obs k_traffic k_ego k_features; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 405 -c 1 -o gpurun_out/prof_$k -f $CMD > gpurun_out/ncu_$k.log 2>&1
echo "$k rc=$?"
done
