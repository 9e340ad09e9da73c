#!/bin/bash
# full GPU check of the round-2 layer-1 kernel: parity suite, bench, per-kernel shares
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -n 5 gpurun_out/r2a_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/r2a_bench.log
tail -n 3 gpurun_out/r2a_bench.log | cut -c1-1500
TB2_SPARSE=tc timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_tc.log 2>&1
tail -n 2 gpurun_out/r2a_bench_tc.log | cut -c1-400
