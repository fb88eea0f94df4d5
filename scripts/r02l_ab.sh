#!/bin/bash
# Round 2l: shared-memory node stride 64 vs 80 B (A/B on one box, one process), parity of the engines, teapot e2e with the pooled image upload
mkdir -p gpurun_out
CONFIGS=book1,cornell RENDERS=4 timeout 600 python scripts/ab_fast.py "CRB_NODE_STRIDE=64" "CRB_NODE_STRIDE=80" "CRB_NODE_STRIDE=64" "CRB_NODE_STRIDE=80" > gpurun_out/r02l_ab_stride.log 2>&1; cat gpurun_out/r02l_ab_stride.log
timeout 600 python -m pytest tests/test_render_parity.py tests/test_trace_parity.py tests/test_golden.py -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02l_pytest.log
timeout 300 python bench.py --config teapot --no-cpu-baseline > gpurun_out/r02l_bench_teapot.log 2>&1; echo "bench rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r02l_bench_teapot.log'):
    if l.startswith('{'):
        d = json.loads(l); print('teapot', d['value'], d['e2e'], d['ms_per_step'])
PY
