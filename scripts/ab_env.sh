#!/bin/bash
# A/B on ONE box: each argument is "lib.so [ENV=value ...]"; runs scripts/sweep_free_pass.py (VALUES=31) with that
# library from variants/ and that environment.   usage: ab_env.sh "lib_b.so CRB_MINB=10" "lib_g.so CRB_PARK_MIN=8" ...
cp crucible_b200/libcrucible_b200.so /tmp/lib_orig.so
for spec in "$@"; do
  set -- $spec
  lib="$1"; shift
  cp variants/$lib crucible_b200/libcrucible_b200.so
  echo "== $lib $*"
  env VALUES=31 "$@" timeout 120 python scripts/sweep_free_pass.py 2>&1 | tail -4
done
cp /tmp/lib_orig.so crucible_b200/libcrucible_b200.so
