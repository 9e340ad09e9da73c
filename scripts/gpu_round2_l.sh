#!/bin/bash
mkdir -p gpurun_out
TB2_PREP_DEBUG=1 bash scripts/gpu_l1.sh ts 2>&1 | grep -i "prep\|call\|small"
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4g ms %.3f e2e %.4g frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']), {k: round(v['avg_us'],1) for k,v in d['roofline']['kernels'].items()})"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_real_scenes.py tests/test_training.py tests/test_dropin.py tests/test_sgan.py tests/test_vae.py -m gpu -q -x 2>&1 | tail -3
