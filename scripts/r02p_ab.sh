#!/bin/bash
# Round 2p: path-record layout (closest hit in the material queue, miss shader reads one 64 B half) against the previous build,
# both libraries on one box, alternating; then the GPU tests on the new one.
mkdir -p gpurun_out
cp crucible_b200/libcrucible_b200.so /tmp/lib_new.so
for r in 1 2; do
  for f in base hitq; do
    cp variants/$f.so crucible_b200/libcrucible_b200.so
    echo "== round $r $f"
    CONFIGS=book1,cornell,teapot,instanced RENDERS=3 timeout 300 python scripts/ab_fast.py "" 2>&1 | tail -4
  done
done > gpurun_out/r02p_ab_hitq.log 2>&1
cat gpurun_out/r02p_ab_hitq.log
cp /tmp/lib_new.so crucible_b200/libcrucible_b200.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02p_pytest.log
