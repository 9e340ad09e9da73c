#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do
  echo "== TB2_BWD_TC=$v"
  TB2_BWD_TC=$v timeout 300 python scripts/profile_train.py social 2>&1 | grep -v "^$" | tail -24 | head -16
  TB2_BWD_TC=$v timeout 200 python scripts/train_bench_social.py 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_training.py tests/test_dropin.py -m gpu -q -x 2>&1 | tail -3
