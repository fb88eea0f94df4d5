#!/usr/bin/env python
"""Summarise ncu output into small text files that can be committed under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/launches_r01.md
  python scripts/summarize_ncu.py kernel   gpurun_out/prof.ncu-rep  profiles/trace_r01.md
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
           "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum",
           "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("crb::", "")
        v = float(r[vi].replace(",", ""))
        if r[ui] == "ns":
            v /= 1e3
        elif r[ui] == "ms":
            v *= 1e3
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"ncu launch list: {src} (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write("| kernel | launches | total us | share | avg us | max us |\n|---|---:|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {a[0]} | {a[1]:.1f} | {a[1] / tot:.1%} | {a[1] / a[0]:.1f} | {a[2]:.1f} |\n")
        f.write(f"\ntotal {tot:.1f} us over {sum(a[0] for a in agg.values())} launches\n")
    print(open(dst).read())


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    with open(dst, "w") as f:
        f.write(f"ncu --set full capture: {src}\n\n")
        names = [r[h.index("Kernel Name")] for r in rows[2:]]
        f.write("kernels: " + "; ".join(n.split("(")[0] for n in names) + "\n\n| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(names))) + " |\n")
        f.write("|---|---|" + "---:|" * len(names) + "\n")
        for m in METRICS:
            if m in h:
                i = h.index(m)
                f.write(f"| {m} | {rows[1][i]} | " + " | ".join(r[i] for r in rows[2:]) + " |\n")
    print(open(dst).read())


def bykernel(src, dst):
    """One column per DISTINCT kernel of an ncu --csv --page raw log (the launch that ran longest), for the all-kernels pass
    of scripts/profile_all.py."""
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    ki, di = h.index("Kernel Name"), h.index("gpu__time_duration.sum")
    best, count, total = collections.OrderedDict(), collections.Counter(), collections.Counter()
    for r in data:
        if len(r) != len(h):
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("crb::", "")
        t = float(r[di].replace(",", ""))
        count[name] += 1
        total[name] += t
        if name not in best or t > float(best[name][di].replace(",", "")):
            best[name] = r
    with open(dst, "w") as f:
        f.write(f"ncu pass over every kernel: {src} (python scripts/profile_all.py; per kernel the launch that ran longest)\n\n")
        short = [m for m in METRICS if m in h]
        f.write("| kernel | launches | total " + units[di] + " | " + " | ".join(m.split(".")[0].replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active", "") for m in short) + " |\n")
        f.write("|---|---:|---:|" + "---:|" * len(short) + "\n")
        for name, r in best.items():
            f.write(f"| {name} | {count[name]} | {total[name]:.0f} | " + " | ".join(r[h.index(m)] for m in short) + " |\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "bykernel": bykernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
