#!/bin/bash
# Round-2 final state on one B200: GPU tests, the driver's two bench arms, the bench line of every BASELINE config.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
tail -3 gpurun_out/r02h_pytest.log
timeout 600 python bench.py > gpurun_out/r02h_bench_book1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/r02h_bench_book1.log | cut -c1-300
timeout 600 python bench.py --impl reference > gpurun_out/r02h_bench_reference.log 2>&1; echo "reference rc=$?"; tail -1 gpurun_out/r02h_bench_reference.log | cut -c1-300
for c in cornell teapot instanced walkthrough; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_bench_$c.log 2>&1
  echo "$c rc=$?"; tail -1 gpurun_out/r02h_bench_$c.log | cut -c1-300
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02h_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02h_smoke.log
