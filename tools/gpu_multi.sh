#!/bin/bash
# multi-GPU check: bench (both arms) and the cfg5 sweep under torchrun with N ranks (N = $1),
# sweep size $2 (default 2^25 draws)
N=${1:-2}
DRAWS=${2:-33554432}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"
cat gpurun_out/bench_${N}gpu.json; tail -5 gpurun_out/bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/bench_configs.py --only cfg5 --sweep-draws $DRAWS > gpurun_out/sweep_${N}gpu.json 2> gpurun_out/sweep_${N}gpu.err; echo "rc=$?"
cat gpurun_out/sweep_${N}gpu.json; tail -5 gpurun_out/sweep_${N}gpu.err
