#!/bin/bash
# multi-GPU check: bench under torchrun with N ranks (N = first argument)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"
cat gpurun_out/bench_${N}gpu.json; tail -5 gpurun_out/bench_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_${N}gpu.json 2> gpurun_out/bench_ref_${N}gpu.err; echo "rc=$?"
cat gpurun_out/bench_ref_${N}gpu.json
