#!/bin/bash
# One GPU-box session: tcgen05 canary first (short timeout), then the full GPU suite, then bench.
# Usage (under gpurun): bash scripts/gpu_check.sh <tag>
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
echo "== canary: tcgen05 dense layer ==" > $out/${tag}_canary.log
timeout 180 python -m pytest tests/test_gpu_parity.py -q -x --timeout 120 \
    -k "single_step_matches_oracle or obs_length_two" >> $out/${tag}_canary.log 2>&1
rc=$?
echo "canary rc=$rc" >> $out/${tag}_canary.log
if [ $rc -ne 0 ]; then
    echo "tcgen05 canary failed -> rest of the session runs with TB2_DISABLE_TC=1" >> $out/${tag}_canary.log
    export TB2_DISABLE_TC=1
fi
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
if [ -z "$TB2_DISABLE_TC" ]; then
    TB2_DISABLE_TC=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 -k "social" \
        > $out/${tag}_pytest_notc.log 2>&1
    echo "pytest(no tc) rc=$?" >> $out/${tag}_pytest_notc.log
fi
timeout 300 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.log 2>&1
timeout 200 python scripts/train_bench.py > $out/${tag}_train_bench.log 2>&1
echo "bench rc=$?" >> $out/${tag}_bench.log
timeout 120 python scripts/profile_e2e.py > $out/${tag}_e2e_profile.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sparse_layer1_tc|lstm_gates_tc|dense_layer_tc|pool_prepare" \
    -s 40 -c 8 -o $out/${tag}_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_full.log 2>&1
timeout 600 python scripts/classical_bench.py > $out/${tag}_classical_bench.log 2>&1
timeout 300 python scripts/configs_bench.py > $out/${tag}_configs_bench.log 2>&1
timeout 300 python scripts/train_bench_social.py > $out/${tag}_train_social.log 2>&1
timeout 300 python scripts/evaluator_bench.py social > $out/${tag}_evaluator.log 2>&1
timeout 300 python scripts/sgan_bench.py social > $out/${tag}_sgan.log 2>&1
tail -4 $out/${tag}_canary.log
grep -E "passed|failed|FAILED|ADE mean|teacher-forced" $out/${tag}_pytest.log | tail -30
[ -f $out/${tag}_pytest_notc.log ] && tail -3 $out/${tag}_pytest_notc.log
tail -2 $out/${tag}_bench.log
cat $out/parity_report.txt 2>/dev/null | tail -20; head -3 $out/${tag}_e2e_profile.log; tail -6 $out/${tag}_e2e_profile.log

tail -2 $out/${tag}_train_bench.log
cat $out/${tag}_classical_bench.log
tail -4 $out/${tag}_configs_bench.log | cut -c1-160
tail -1 $out/${tag}_train_social.log; tail -1 $out/${tag}_evaluator.log; tail -1 $out/${tag}_sgan.log
