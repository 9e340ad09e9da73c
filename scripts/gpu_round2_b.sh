#!/bin/bash
# drop-in tests + both bench arms on one GPU; with 2 GPUs also the torchrun launch (training all-reduce)
mkdir -p gpurun_out
N=${1:-1}
timeout 600 python -m pytest tests/test_dropin.py tests/test_loss.py tests/test_training.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
tail -n 6 gpurun_out/r2b_pytest.log
if [ "$N" = "1" ]; then
  timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/r2b_bench.log
  tail -n 2 gpurun_out/r2b_bench.log | cut -c1-3000
  timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2b_ref.log 2>&1; echo "ref exit $?" >> gpurun_out/r2b_ref.log
  tail -n 2 gpurun_out/r2b_ref.log | cut -c1-1500
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2b_bench_${N}gpu.log 2>&1; echo "bench exit $?" >> gpurun_out/r2b_bench_${N}gpu.log
  tail -n 2 gpurun_out/r2b_bench_${N}gpu.log | cut -c1-3000
fi
