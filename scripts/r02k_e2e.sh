#!/bin/bash
# Round 2k: host staging cache + cr_scene_reserve + staged H2D.  GPU tests, then the end-to-end breakdown of the 10 M-triangle
# config with the staged copy on and off, then the bench lines of the two configs whose e2e was staging bound.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02k_pytest.log
CONFIG=instanced SPP=4 CRB_STAGED_H2D=1 timeout 300 python scripts/e2e_breakdown.py > gpurun_out/r02k_e2e_instanced_staged.log 2>&1; tail -3 gpurun_out/r02k_e2e_instanced_staged.log
CONFIG=instanced SPP=4 CRB_STAGED_H2D=0 timeout 300 python scripts/e2e_breakdown.py > gpurun_out/r02k_e2e_instanced_plain.log 2>&1; tail -3 gpurun_out/r02k_e2e_instanced_plain.log
timeout 600 python bench.py --config instanced > gpurun_out/r02k_bench_instanced.log 2>&1; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02k_bench_instanced.log
