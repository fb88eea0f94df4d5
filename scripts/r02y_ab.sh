#!/bin/bash
mkdir -p gpurun_out
cp crucible_b200/libcrucible_b200.so /tmp/lib_new.so
{
for r in 1 2; do
  for f in prequad quadpad; do
    cp variants/$f.so crucible_b200/libcrucible_b200.so
    echo "== round $r $f"; CONFIGS=cornell RENDERS=3 timeout 200 python scripts/ab_fast.py "" 2>&1 | tail -1
  done
done
} > gpurun_out/r02y_ab_quadpad.log 2>&1
cp /tmp/lib_new.so crucible_b200/libcrucible_b200.so
cat gpurun_out/r02y_ab_quadpad.log
timeout 300 python -m pytest tests/test_render_parity.py tests/test_trace_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
