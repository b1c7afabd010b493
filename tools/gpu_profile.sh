#!/bin/bash
# ncu evidence for the bench workload (1 GPU) in its steady state (400 warm-up steps).  Plain run first (must exit 0),
# then the launch list, then one --set full capture of each step kernel.
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 400 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
tail -1 gpurun_out/plain.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -s 1610 -c 160 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_lidar_obs -s 405 -c 1 -o gpurun_out/prof_lidar -f $CMD > gpurun_out/ncu_lidar.log 2>&1
echo "lidar rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_traffic -s 405 -c 1 -o gpurun_out/prof_dyn -f $CMD > gpurun_out/ncu_dyn.log 2>&1
echo "dyn rc=$?"
ls -la gpurun_out | head -20
ncu --set full --clock-control none --import-source on -k regex:k_ego -s 405 -c 1 -o gpurun_out/prof_ego -f $CMD > gpurun_out/ncu_ego.log 2>&1
echo "ego rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_features -s 405 -c 1 -o gpurun_out/prof_feat -f $CMD > gpurun_out/ncu_feat.log 2>&1
echo "feat rc=$?"
