#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 2>gpurun_out/p_bench.err | tail -1 > gpurun_out/p_bench.json
python -c "import sys,json; d=json.load(open('gpurun_out/p_bench.json')); print('value %.4g ms %.3f e2e %.4g (median %.3f) frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_median'], d['roofline']['frac']), {k: round(v['avg_us'],1) for k,v in d['roofline']['kernels'].items()})"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_real_scenes.py tests/test_dropin.py tests/test_training.py -m gpu -q -x 2>&1 | tail -3
