#!/bin/bash
# ncu launch list of the bench command in its steady state (plain run first; ONE ncu pass per gpurun call).
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 400 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
tail -1 gpurun_out/plain.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -s 1610 -c 160 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
