#!/bin/bash
# Round-end check on the GPU box: bench line (clock samples, traffic source) + every BASELINE config at full size.
timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_check.log 2>&1
tail -1 gpurun_out/bench_check.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["value"], d["clocks"], d["roofline"]["traffic"], d["roofline"]["traffic_source"][:40])'
timeout 250 python scripts/run_configs.py "$@" > gpurun_out/configs_r01d.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/configs_r01d.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["config"], d["precision"], round(d["msamples_per_s"], 1), round(d["mrays_per_s"], 1), round(d["ms_total"], 1), round(d["ms_trace"], 1), d["host_commit_s"])
PY
