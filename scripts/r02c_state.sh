#!/bin/bash
# Round-2 state check on one B200: GPU tests, then the bench line of every BASELINE config (one code state).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02c_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -3 gpurun_out/r02c_pytest.log
for c in book1 cornell teapot instanced walkthrough; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_$c.log 2>&1
  echo "$c rc=$?"; tail -1 gpurun_out/r02c_bench_$c.log | cut -c1-400
done
