#!/bin/bash
# GPU regression script: smoke, GPU tests, bench.  Everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/gpu.txt
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout ${PYTEST_TIMEOUT:-2400} python -m pytest tests -m gpu -x -q --durations=12 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.log | cut -c1-1500; tail -5 gpurun_out/bench.err
