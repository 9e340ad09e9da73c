#!/bin/bash
# ncu evidence for round 2 (final default path): launch list (shares) + one full capture of the dominant kernel + sanitizer runs
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/r2_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sparse_layer1_pair -s 60 -c 2 -o gpurun_out/r2_prof_pair_ts $CMD > gpurun_out/r2_ncu_full.log 2>&1
echo "full capture exit $?"; ls -la gpurun_out/r2_prof_pair_ts.ncu-rep
timeout 600 compute-sanitizer --tool memcheck python scripts/l1_check.py --mode ts --iters 2 > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/r2_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python scripts/l1_check.py --mode ts --iters 2 > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck exit $?"; tail -4 gpurun_out/r2_racecheck.log
