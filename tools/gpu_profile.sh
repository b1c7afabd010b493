#!/bin/bash
# ncu evidence for the bench workload (1 GPU).  Plain run first (must exit 0), then the launch list, then one
# --set full capture of each step kernel.
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
tail -1 gpurun_out/plain.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_lidar_obs -s 6 -c 2 -o gpurun_out/prof_lidar -f $CMD > gpurun_out/ncu_lidar.log 2>&1
echo "lidar rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_dynamics -s 6 -c 2 -o gpurun_out/prof_dyn -f $CMD > gpurun_out/ncu_dyn.log 2>&1
echo "dyn rc=$?"
ls -la gpurun_out
