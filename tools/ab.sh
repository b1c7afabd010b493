#!/bin/bash
# A/B of tuning builds in the steady state of the bench workload:  tools/ab.sh [envs] <variant .so> ...
# ("default" = the in-tree library).  400 warm steps, then 2 x 100 timed steps with an event pair around every kernel.
E=${ENVS:-65536}
for so in "$@"; do
  echo "== $so"
  if [ "$so" = default ]; then WARM=400 python tools/perf_timeline.py $E 2 100 2>&1 | grep -E "steps|mean:|Error|error"
  else ISX_LIB=$PWD/$so WARM=400 python tools/perf_timeline.py $E 2 100 2>&1 | grep -E "steps|mean:|Error|error"; fi
done
