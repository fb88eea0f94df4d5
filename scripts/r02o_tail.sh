#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_render_parity.py tests/test_golden.py tests/test_nested_elements.py tests/test_object_animation.py -m gpu -x -q > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02o_pytest.log
timeout 400 python scripts/tail_round_sweep.py > gpurun_out/r02o_tail_sweep.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/r02o_tail_sweep.log | tail -30
