#!/bin/bash
# full ncu capture of pool_prepare_kernel (warp-state statistics: where do its 10 us go?)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train"
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:pool_prepare -s 60 -c 2 -o gpurun_out/r2_prof_prep $CMD > gpurun_out/r2_ncu_prep.log 2>&1
echo "exit $?"; ls -la gpurun_out/r2_prof_prep.ncu-rep
