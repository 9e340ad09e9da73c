#!/bin/bash
# A/B of the layer-1 kernels on one B200 (each mode in its own process under a timeout: a deadlocked
# tcgen05 pipeline must not take the whole call down)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/l1_gpu.txt 2>&1
for mode in "$@"; do
  timeout 150 python scripts/l1_check.py --mode $mode > gpurun_out/l1_$mode.log 2>&1
  echo "mode $mode exit $?" >> gpurun_out/l1_$mode.log
  grep -v "^  unit" gpurun_out/l1_$mode.log | tail -n 30
done
