#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_dropin.py 2>&1 | grep -v Warning | grep -v "out\[name\]" > gpurun_out/r2d_debug_dropin.log; tail -25 gpurun_out/r2d_debug_dropin.log
timeout 600 python -m pytest tests/test_dropin.py tests/test_nongrid.py tests/test_training.py -m gpu -q -s 2>&1 | grep -v Warning > gpurun_out/r2d_pytest.log; echo "pytest exit $?" >> gpurun_out/r2d_pytest.log
grep -E "worst|passed|failed|FAILED|Error" gpurun_out/r2d_pytest.log | tail -20
