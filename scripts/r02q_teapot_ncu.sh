#!/bin/bash
# teapot trace: the record-layout change costs 8 % there (and gains 3-5 % on book1 / the 10 M-triangle scene): what differs?
mkdir -p gpurun_out
M="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"
cp crucible_b200/libcrucible_b200.so /tmp/lib_new.so
B="python bench.py --config teapot --steps 1 --warmup 3 --no-cpu-baseline"
for f in base hitq; do
  cp variants/$f.so crucible_b200/libcrucible_b200.so
  $B > gpurun_out/r02q_plain_$f.log 2>&1 && ncu --metrics $M --clock-control none -k regex:k_trace_fast -s 4 -c 3 --csv --log-file gpurun_out/r02q_teapot_$f.csv $B > gpurun_out/r02q_ncu_$f.log 2>&1
  echo "$f rc=$?"
done
cp /tmp/lib_new.so crucible_b200/libcrucible_b200.so
python - <<'PY'
import csv
for f in ("base", "hitq"):
    rows = [r for r in csv.reader(open(f"gpurun_out/r02q_teapot_{f}.csv")) if len(r) > 10]
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    by = {}
    for r in rows[1:]:
        by.setdefault(r[ix["ID"]], {})[r[ix["Metric Name"]]] = r[ix["Metric Value"]]
    for k, v in by.items():
        print(f, k, {a.split("__")[-1][:44]: b for a, b in v.items()})
PY
