#!/bin/bash
# Host-buffer step at the bench size under different host-thread counts / pipeline plans / chunk counts (tuning aid).
E=${1:-65536}
for th in 14 12 10 8; do
echo "== ISX_HOST_THREADS=$th"; ISX_HOST_THREADS=$th python tools/e2e_probe.py $E 2>&1 | grep -E "step_host|isx_step_pinned"
done
for cfg in "1,2,3,4,6:4" "1,2,3,4,6:8" "1,2,3,4,6:2" "1,1,2,3,4,5:4" "1,2,4,4,5:4" "1,2,3,5,8:4" "1,2,3,4,5,6,7:4" "2,3,4,5,6:4" "1,1,1,1:4" "1,1,2,2,3,3,4:2"; do
plan=${cfg%%:*}; ch=${cfg##*:}
echo "== ISX_PIPE_PLAN=$plan ISX_PIPE_CHUNKS=$ch (threads 12)"; ISX_HOST_THREADS=12 ISX_PIPE_PLAN=$plan ISX_PIPE_CHUNKS=$ch python tools/e2e_probe.py $E 2>&1 | grep -E "step_host|isx_step_pinned"
done
