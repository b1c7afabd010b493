#!/bin/bash
# A/B of tuning builds: tools/ab.sh <variant .so> ...   (steady state: 300 warm steps, then 2 x 200 timed)
for so in "$@"; do echo "== $so"; ISX_LIB=$PWD/$so WARM=300 python tools/perf_timeline.py 8192 2 200 2>&1 | grep -E "steps|mean:"; done
