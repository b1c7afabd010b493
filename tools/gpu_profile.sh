#!/bin/bash
# ncu evidence for the bench workload (1 GPU) in its steady state.  Plain run first (must exit 0), then the launch list,
# then one --set full capture (+ the L2 / shared-memory counters north_star asks for) of each step kernel.
#   tools/gpu_profile.sh [tag] [bench args...]        e.g. tools/gpu_profile.sh r02 --config C5
TAG=${1:-cur}; shift
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline $*"
EXTRA="lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_bytes.sum,smsp__thread_inst_executed.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum"
$CMD > gpurun_out/plain_$TAG.log 2> gpurun_out/plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.err; exit 1; }
tail -1 gpurun_out/plain_$TAG.log | cut -c1-300
# bench launches per step: k_traffic_order + 4 kernels (+ the L2 flush fill in the timed region); warm-up 5 + pre-roll 400 steps come first
ncu --metrics gpu__time_duration.sum --clock-control none -s 2060 -c 192 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
for k in k_lidar_obs k_traffic k_ego k_features; do
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k "regex:$k(<|\$)" -s 410 -c 1 -o gpurun_out/prof_${k}_$TAG -f $CMD > gpurun_out/ncu_${k}_$TAG.log 2>&1
echo "$k rc=$?"
done
