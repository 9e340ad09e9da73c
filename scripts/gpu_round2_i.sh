#!/bin/bash
mkdir -p gpurun_out
for v in "" "TB2_NO_FUSE2=1"; do
  env $v timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), {k: round(v['avg_us'],1) for k,v in d['roofline']['kernels'].items()})"
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_real_scenes.py -m gpu -q -x 2>&1 | tail -3
