#!/bin/bash
# ncu evidence of the final round-2 code (every ncu command runs only after the same command has exited 0 without ncu:
# gpurun does that itself, and the plain runs below keep their logs).
mkdir -p gpurun_out
M=$(python -c "import sys; sys.path.insert(0,'scripts'); import summarize_ncu as s; print(','.join(s.METRICS))")
B1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
B1S="python bench.py --samples 16 --steps 1 --warmup 3 --no-cpu-baseline"
$B1S > gpurun_out/r02w_plain_book1_s16.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02w.csv $B1S > gpurun_out/r02w_ncu0.log 2>&1
echo "launch list rc=$?"
$B1 > gpurun_out/r02w_plain_book1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_fast -s 5 -c 1 -f -o gpurun_out/trace_book1_r02w $B1 > gpurun_out/r02w_ncu1.log 2>&1
echo "book1 trace rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_shade_ -s 24 -c 4 -f -o gpurun_out/shade_book1_r02w $B1 > gpurun_out/r02w_ncu2.log 2>&1
echo "book1 shade rc=$?"
B4="python bench.py --config instanced --steps 1 --warmup 3 --no-cpu-baseline"
$B4 > gpurun_out/r02w_plain_instanced.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_fast -s 3 -c 1 -f -o gpurun_out/trace_cfg4_r02w $B4 > gpurun_out/r02w_ncu3.log 2>&1
echo "instanced trace rc=$?"
PA="python scripts/profile_all.py"
$PA > gpurun_out/r02w_plain_profile_all.log 2>&1 &&
ncu --metrics $M --clock-control none --csv --page raw --log-file gpurun_out/all_kernels_r02w.csv $PA > gpurun_out/r02w_ncu4.log 2>&1
echo "all kernels rc=$?"; tail -2 gpurun_out/r02w_plain_profile_all.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02w_pytest.log; tail -3 gpurun_out/r02w_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02w_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02w_smoke.log
