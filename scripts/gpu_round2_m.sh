#!/bin/bash
mkdir -p gpurun_out
for v in 1 2 0; do
  echo "== TB2_CUBLAS=$v"
  TB2_CUBLAS=$v timeout 300 python scripts/profile_train.py social 2>&1 | grep -v "^$" | tail -22 | head -12
  TB2_CUBLAS=$v timeout 200 python scripts/train_bench_social.py 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_training.py tests/test_dropin.py -m gpu -q -x 2>&1 | tail -3
