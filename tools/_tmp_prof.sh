mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q > gpurun_out/parity_c.log 2>&1; tail -3 gpurun_out/parity_c.log
bash tools/ab.sh default > gpurun_out/ab_c.log 2>&1; grep -E "==|mean:|rror" gpurun_out/ab_c.log
