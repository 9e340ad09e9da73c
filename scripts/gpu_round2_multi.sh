#!/bin/bash
# multi-GPU records: bench.py (inference sharded, training sub-record with the NCCL all-reduce) and the classical launcher
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}gpu.log 2>&1; echo "bench exit $?"
grep '^{' gpurun_out/r2_bench_${N}gpu.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.4g ms %.3f n_gpus %d' % (d['value'], d['ms_per_step'], d['n_gpus'])); print('e2e', d['e2e']['value'], d['e2e']['ms_median']); print('train', json.dumps(d['train']))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/classical_bench.py $((6250 * N)) 1000 > gpurun_out/r2_classical_${N}gpu.log 2>&1; echo "classical exit $?"
grep '^{' gpurun_out/r2_classical_${N}gpu.log | cut -c1-400
