#!/bin/bash
# Host-buffer step at the bench size under different host-thread counts / store kinds / pipeline plans (tuning aid).
E=${1:-65536}
for nt in 1 0; do for th in 16 8 4; do
echo "== ISX_HOST_THREADS=$th ISX_EXPAND_NT=$nt"; ISX_HOST_THREADS=$th ISX_EXPAND_NT=$nt python tools/e2e_probe.py $E 2>&1 | grep -E "step_host|isx_step_pinned"
done; done
for plan in "1,1,1,1" "1,1,1,1,1,1,1,1,1,1,1,1" "1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1"; do
echo "== ISX_PIPE_PLAN=$plan"; ISX_PIPE_PLAN=$plan python tools/e2e_probe.py $E 2>&1 | grep -E "step_host|isx_step_pinned|range"
done
