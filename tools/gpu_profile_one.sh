mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 400 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$1 -s 405 -c 1 -o gpurun_out/prof_$2 -f $CMD > gpurun_out/ncu_$2.log 2>&1
echo "$2 rc=$?"
