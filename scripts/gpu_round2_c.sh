#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_dropin.py > gpurun_out/r2c_debug_dropin.log 2>&1; echo "exit $?" >> gpurun_out/r2c_debug_dropin.log
cat gpurun_out/r2c_debug_dropin.log | tail -40
timeout 300 python -m pytest tests/test_nongrid.py tests/test_training.py tests/test_real_scenes.py tests/test_data_io.py -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_pytest.log
tail -n 15 gpurun_out/r2c_pytest.log
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_ref.log 2>&1; echo "ref exit $?" >> gpurun_out/r2c_ref.log
tail -n 2 gpurun_out/r2c_ref.log | cut -c1-1600
