#!/bin/bash
mkdir -p gpurun_out
timeout 500 python scripts/debug_social_grads2.py 2>&1 | grep -v Warning | grep -v "out\[name\]" > gpurun_out/r2f_debug.log; tail -45 gpurun_out/r2f_debug.log
