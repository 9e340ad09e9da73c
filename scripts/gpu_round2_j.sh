#!/bin/bash
# full validation of the default path: GPU suite, smoke, bench (both arms), records for profiles/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/j_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 > gpurun_out/j_smoke.txt
timeout 400 python bench.py 2>gpurun_out/j_bench.err | tail -1 > gpurun_out/j_bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/j_ref.err | tail -1 > gpurun_out/j_bench_ref.json
timeout 300 python scripts/configs_bench.py 2>&1 | tail -8 > gpurun_out/j_configs.txt
timeout 300 python scripts/classical_bench.py 2>&1 | tail -4 > gpurun_out/j_classical.txt
cat gpurun_out/j_pytest.txt gpurun_out/j_smoke.txt; head -c 600 gpurun_out/j_bench.json; echo; head -c 400 gpurun_out/j_bench_ref.json
timeout 200 python scripts/train_bench_social.py 2>&1 | tail -1 > gpurun_out/j_train_social.json
timeout 200 python scripts/train_bench.py 2>&1 | tail -1 > gpurun_out/j_train_dlstm.json
cat gpurun_out/j_train_social.json gpurun_out/j_train_dlstm.json | cut -c1-300
