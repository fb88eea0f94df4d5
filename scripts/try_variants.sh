#!/bin/bash
# Runs the bench once per library variant under variants/ (kernel-tuning helper; the box copy is scratch).
# usage: try_variants.sh "<bench args>" lib_a.so lib_b.so ...
ARGS="$1"; shift
cp crucible_b200/libcrucible_b200.so /tmp/lib_orig.so
for f in "$@"; do
  cp variants/$f crucible_b200/libcrucible_b200.so
  echo "== $f"
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline $ARGS 2>&1 | tail -1 | python scripts/benchline.py
done
cp /tmp/lib_orig.so crucible_b200/libcrucible_b200.so
