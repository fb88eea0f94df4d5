#!/bin/bash
mkdir -p gpurun_out
cp crucible_b200/libcrucible_b200.so /tmp/lib_new.so
{
for r in 1 2; do
cp variants/miss2.so crucible_b200/libcrucible_b200.so
echo "== round $r miss2"; CONFIGS=book1,cornell,teapot RENDERS=3 timeout 300 python scripts/ab_fast.py "" 2>&1 | tail -3
cp variants/pf.so crucible_b200/libcrucible_b200.so
echo "== round $r pf"; CONFIGS=book1,cornell,teapot RENDERS=3 timeout 300 python scripts/ab_fast.py "CRB_SHADE_PF=0" "CRB_SHADE_PF=1" 2>&1 | tail -6
done
} > gpurun_out/r02t_ab_pf.log 2>&1
cp /tmp/lib_new.so crucible_b200/libcrucible_b200.so
cat gpurun_out/r02t_ab_pf.log
