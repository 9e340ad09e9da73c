#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scripts/debug_dropin.py 2>&1 | grep -v Warning | grep -v "out\[name\]" > gpurun_out/r2e_debug_dropin.log; grep -E "^social|hidden_dim_encoding.weight" gpurun_out/r2e_debug_dropin.log | tail -40
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4g ms %.3f e2e' % (d['value'], d['ms_per_step']), d['e2e'])"
