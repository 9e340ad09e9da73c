#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r_bench.err | tail -1 > gpurun_out/r_bench.json
python -c "import sys,json; d=json.load(open('gpurun_out/r_bench.json')); print('value %.4g ms %.3f e2e %.4g (median %.3f p95 %.3f) frac %.3f train %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_median'], d['e2e']['ms_p95'], d['roofline']['frac'], d['train']['ms_per_step']))"
timeout 200 python scripts/train_bench_social.py 2>&1 | tail -1 | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_training.py tests/test_sgan.py tests/test_vae.py -m gpu -q -x 2>&1 | tail -2
