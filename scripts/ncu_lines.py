#!/usr/bin/env python
"""Per-source-line view of an ncu capture (needs -lineinfo, the SAME build of the library, and nvdisasm here):
  python scripts/ncu_lines.py capture.ncu-rep crucible_b200/libcrucible_b200.so k_trace_fast [top]
Joins the SASS page of the capture (warp stall samples, instructions, L1 wavefronts per instruction) with nvdisasm's line
table of the kernel and prints the hottest source lines."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, lib, pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# the first launch whose name matches `pat` (a report may hold several kernels)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--kernel-name", "regex:" + pat, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
kname = rows[0][1]
hdr = rows[1]
body = rows[2:]
for i, r in enumerate(body):  # a report with several matching launches repeats the table: keep the first
    if r and r[0] == "Kernel Name":
        body = body[:i]
        break
data = [dict(zip(hdr, r)) for r in body if len(r) == len(hdr)]
base = int(data[0]["Address"], 16)
mangled_hint = re.sub(r"[^A-Za-z0-9_]", "", kname.split("(")[0].split("::")[-1].split("<")[0])
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
    line_of = None
    for f in sorted(os.listdir(td)):
        if not f.endswith(".cubin"):
            continue
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, f)], capture_output=True, text=True).stdout
        # split per function
        cur_fn, cur_line, table, n_ins = None, None, {}, {}
        for ln in dis.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                cur_fn = m.group(1)
                table[cur_fn] = {}
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
            if m and cur_fn:
                table[cur_fn][int(m.group(1), 16)] = cur_line
        for fn, t in table.items():
            if pat in fn and len(t) == len(data):
                line_of = t
                print("matched", fn, "in", f, len(t), "instructions")
                break
        if line_of:
            break
if not line_of:
    sys.exit(f"no function containing {pat!r} with {len(data)} instructions found (was the library rebuilt since the capture?)")
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
for d in data:
    off = int(d["Address"], 16) - base
    key = line_of.get(off)
    a = agg[key]
    a[0] += int(d.get("# Samples", "0") or 0)
    a[1] += int(d.get("Instructions Executed", "0") or 0)
    a[2] += int(d.get("Thread Instructions Executed", "0") or 0)
    a[3] += int(d.get("L1 Wavefronts Shared", "0") or 0)
    a[4] += int(d.get("L2 Theoretical Sectors Local", "0") or 0) + int(d.get("L2 Theoretical Sectors Global", "0") or 0)
ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"kernel {kname[:80]}  samples {ts}  warp instructions {ti}")
src_cache = {}
def src(key):
    if not key:
        return ""
    path = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", key[0])
    if path not in src_cache:
        try:
            src_cache[path] = open(path).read().splitlines()
        except Exception:
            src_cache[path] = []
    l = src_cache[path]
    return l[key[1] - 1].strip()[:100] if 0 < key[1] <= len(l) else ""
print("samples%  instr%  lanes  file:line  source")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    lanes = a[2] / a[1] if a[1] else 0
    print(f"{a[0] / ts * 100:6.2f}  {a[1] / ti * 100:6.2f}  {lanes:5.1f}  {key[0] if key else '?'}:{key[1] if key else 0}  {src(key)}")
