#!/bin/bash
mkdir -p gpurun_out
for v in "" "TB2_SPARSE=tc" "TB2_DISABLE_TC=1"; do
  env $v timeout 300 python scripts/debug_dropin.py 2>&1 | grep -v Warning | grep -v "out\[name\]" | grep -E "^env|^social|pool\." 
done > gpurun_out/r2g_debug.log
cat gpurun_out/r2g_debug.log
