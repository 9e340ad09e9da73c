#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do echo "TB2_GRID_TC=$v"; TB2_GRID_TC=$v timeout 280 python scripts/configs_bench.py 256 2>&1 | grep "^{" | head -3 | cut -c1-330; done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_real_scenes.py tests/test_training.py tests/test_dropin.py -m gpu -q -x 2>&1 | tail -3
timeout 200 python scripts/train_bench.py 2>&1 | tail -1 | cut -c1-250
