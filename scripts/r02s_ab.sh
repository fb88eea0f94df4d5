#!/bin/bash
mkdir -p gpurun_out
cp crucible_b200/libcrucible_b200.so /tmp/lib_new.so
for r in 1 2; do
  for f in hitq2 miss2; do
    cp variants/$f.so crucible_b200/libcrucible_b200.so
    echo "== round $r $f"
    CONFIGS=book1,cornell,teapot RENDERS=3 timeout 300 python scripts/ab_fast.py "" 2>&1 | tail -3
  done
done > gpurun_out/r02s_ab_miss.log 2>&1
cp /tmp/lib_new.so crucible_b200/libcrucible_b200.so
cat gpurun_out/r02s_ab_miss.log
