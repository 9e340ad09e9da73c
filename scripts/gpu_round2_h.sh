#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_classical.py -m gpu -q 2>&1 | tail -3
timeout 300 python scripts/classical_bench.py 6250 1000 > gpurun_out/r2h_classical.log 2>&1; tail -4 gpurun_out/r2h_classical.log | cut -c1-700
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench.log 2>&1; tail -1 gpurun_out/r2h_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4g ms %.3f' % (d['value'], d['ms_per_step'])); print('e2e', d['e2e']); print('train', d['train'])"
