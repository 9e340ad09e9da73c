#!/bin/bash
# TS default vs SS pair: kernel debug counters + bench A/B + parity suite
mkdir -p gpurun_out
TB2_L1_DEBUG=1 bash scripts/gpu_l1.sh ts | grep -v "^  unit" | tail -6
for v in "" "TB2_SPARSE=pair"; do
  env $v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v value %.4g ms %.3f e2e %.4g frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']), {k: round(v['avg_us'],1) for k,v in d['roofline']['kernels'].items()})"
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_real_scenes.py tests/test_dropin.py tests/test_training.py -m gpu -q -x 2>&1 | tail -3
