#!/bin/bash
# Round-2 final state on one B200 (the GPU tests and smoke() of the same code ran in scripts/r02w_ncu.sh): the driver's two bench
# arms and the bench line of every BASELINE config, one box.
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r02x_bench_book1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r02x_bench_book1.log | cut -c1-200
timeout 300 python bench.py --impl reference > gpurun_out/r02x_bench_reference.log 2>&1; echo "reference rc=$?"; tail -1 gpurun_out/r02x_bench_reference.log | cut -c1-200
for c in cornell teapot instanced walkthrough; do
  timeout 300 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02x_bench_$c.log 2>&1
  echo "$c rc=$?"; tail -1 gpurun_out/r02x_bench_$c.log | cut -c1-200
done
