#!/bin/bash
# A/B of the layer-1 kernels on one B200 (each mode in its own process under a timeout: a deadlocked
# tcgen05 pipeline must not take the whole call down).  A mode may carry an environment prefix: "TB2_NO_FUSE2=1:pair"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/l1_gpu.txt 2>&1
for spec in "$@"; do
  mode=${spec##*:}; envs=""; [ "$spec" != "$mode" ] && envs=${spec%:*}
  tag=$(echo "$spec" | tr ':=' '__')
  env $envs timeout 150 python scripts/l1_check.py --mode $mode > gpurun_out/l1_$tag.log 2>&1
  echo "mode $spec exit $?" >> gpurun_out/l1_$tag.log
  grep -v "^  unit" gpurun_out/l1_$tag.log | tail -n 30
done
