#!/bin/bash
# Latency-floor profiling: tiny batch (512 envs) so that each kernel is a single wave and the stall samples show where
# ONE warp's dependent chain spends its time.
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 400 --no-cpu-baseline --envs-per-gpu 512"
$CMD > gpurun_out/plain_small.log 2> gpurun_out/plain_small.err || { echo "plain run failed"; tail -5 gpurun_out/plain_small.err; exit 1; }
for k in k_traffic k_ego k_features; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 405 -c 1 -o gpurun_out/small_$k -f $CMD > gpurun_out/ncu_small_$k.log 2>&1
echo "$k rc=$?"
done
