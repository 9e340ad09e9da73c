#!/bin/bash
# compute-sanitizer memcheck over the kernels added late in round 2 (non-grid pools, mma.sync backward kernels, pool_prepare)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_nongrid.py -m gpu -q -x -k "plug or forward_matches" > gpurun_out/r2_memcheck_nongrid.log 2>&1; echo "nongrid exit $?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/r2_memcheck_nongrid.log | tail -3
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_training.py -m gpu -q -x -k "social" > gpurun_out/r2_memcheck_training.log 2>&1; echo "training exit $?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/r2_memcheck_training.log | tail -3
