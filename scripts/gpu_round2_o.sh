#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/profile_train.py social 2>&1 | grep -v "^$" | tail -24 | head -18
timeout 200 python scripts/train_bench_social.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_training.py tests/test_dropin.py tests/test_parallel.py -m gpu -q -x 2>&1 | tail -3
