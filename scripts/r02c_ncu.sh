#!/bin/bash
# ncu captures of the round-2 engines (each command first runs plain; gpurun does that itself for ncu commands).
mkdir -p gpurun_out
export CRB_TRAVERSAL=f
B1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$B1 > gpurun_out/r02c_plain_book1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_fast -s 5 -c 1 -f -o gpurun_out/trace_fast_book1_r02c $B1 > gpurun_out/r02c_ncu1.log 2>&1
echo "book1 trace rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_shade_ -s 24 -c 4 -f -o gpurun_out/shade_book1_r02c $B1 > gpurun_out/r02c_ncu2.log 2>&1
echo "book1 shade rc=$?"
B4="python bench.py --config instanced --steps 1 --warmup 3 --no-cpu-baseline"
$B4 > gpurun_out/r02c_plain_instanced.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_fast -s 3 -c 1 -f -o gpurun_out/trace_fast_cfg4_r02c $B4 > gpurun_out/r02c_ncu3.log 2>&1
echo "instanced trace rc=$?"
unset CRB_TRAVERSAL
B2="python bench.py --config cornell --samples 100 --steps 1 --warmup 3 --no-cpu-baseline"
$B2 > gpurun_out/r02c_plain_cornell.log 2>&1 && tail -1 gpurun_out/r02c_plain_cornell.log | grep -o '"trace_engine": "[^"]*"'
CRB_TRAVERSAL=f $B2 > gpurun_out/r02c_plain_cornell_fast.log 2>&1; tail -1 gpurun_out/r02c_plain_cornell_fast.log | grep -o '"ms_trace": [0-9.]*'
CRB_TRAVERSAL=r CRB_MINB=8 $B2 > gpurun_out/r02c_plain_cornell_r8.log 2>&1; tail -1 gpurun_out/r02c_plain_cornell_r8.log | grep -o '"ms_trace": [0-9.]*'
CRB_TRAVERSAL=r CRB_MINB=10 $B2 > gpurun_out/r02c_plain_cornell_r10.log 2>&1; tail -1 gpurun_out/r02c_plain_cornell_r10.log | grep -o '"ms_trace": [0-9.]*'
for c in book1 teapot; do
CRB_TRAVERSAL=r CRB_MINB=10 python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_plain_${c}_r10.log 2>&1; tail -1 gpurun_out/r02c_plain_${c}_r10.log | grep -o '"ms_trace": [0-9.]*'
done
