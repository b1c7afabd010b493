mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_round2.py -x -q > gpurun_out/parity_b.log 2>&1; tail -3 gpurun_out/parity_b.log
bash tools/ab.sh default > gpurun_out/ab_b.log 2>&1; grep -E "==|mean:" gpurun_out/ab_b.log
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config C5"
for k in k_features k_traffic k_ego; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 410 -c 1 -o gpurun_out/prof_${k}_r02b -f $CMD > gpurun_out/ncu_${k}_r02b.log 2>&1
echo "$k rc=$?"
done
