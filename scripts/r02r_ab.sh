#!/bin/bash
mkdir -p gpurun_out
cp crucible_b200/libcrucible_b200.so /tmp/lib_new.so
{
cp variants/base.so crucible_b200/libcrucible_b200.so
echo "== base"; CONFIGS=book1,cornell,teapot,instanced RENDERS=3 timeout 300 python scripts/ab_fast.py "" 2>&1 | tail -4
cp /tmp/lib_new.so crucible_b200/libcrucible_b200.so
echo "== new"; CONFIGS=book1,cornell,teapot,instanced RENDERS=3 timeout 300 python scripts/ab_fast.py "CRB_REC_BYPASS=0" "CRB_REC_BYPASS=1" "CRB_REC_BYPASS=0" "CRB_REC_BYPASS=1" 2>&1 | tail -16
} > gpurun_out/r02r_ab_bypass.log 2>&1
cat gpurun_out/r02r_ab_bypass.log
