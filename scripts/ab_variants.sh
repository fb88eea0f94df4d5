#!/bin/bash
# A/B of library builds under variants/ within ONE box (numbers from different gpurun calls differ by 1-2 %):
# per-scene f64 trace / total milliseconds through scripts/sweep_free_pass.py.   usage: ab_variants.sh rounds lib_a.so lib_b.so ...
ROUNDS="$1"; shift
cp crucible_b200/libcrucible_b200.so /tmp/lib_orig.so
for r in $(seq 1 "$ROUNDS"); do
  for f in "$@"; do
    cp variants/$f crucible_b200/libcrucible_b200.so
    echo "== round $r $f"
    VALUES="31" timeout 120 python scripts/sweep_free_pass.py 2>&1 | tail -4
  done
done
cp /tmp/lib_orig.so crucible_b200/libcrucible_b200.so
