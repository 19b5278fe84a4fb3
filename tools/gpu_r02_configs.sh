#!/bin/bash
# round 2: every BASELINE configuration, the occupation families, small batches / latency on one GPU
mkdir -p gpurun_out
timeout 900 python tools/bench_configs.py --sweep-draws 16777216 > gpurun_out/configs_r02.jsonl 2> gpurun_out/configs_r02.err; echo "configs rc=$?"
cat gpurun_out/configs_r02.jsonl; tail -3 gpurun_out/configs_r02.err
timeout 600 python tools/bench_families.py > gpurun_out/families_r02.jsonl 2> gpurun_out/families_r02.err; echo "families rc=$?"
cat gpurun_out/families_r02.jsonl; tail -3 gpurun_out/families_r02.err
timeout 600 python tools/bench_small_batches.py > gpurun_out/small_batches_r02.jsonl 2> gpurun_out/small_batches_r02.err; echo "small rc=$?"
cat gpurun_out/small_batches_r02.jsonl | cut -c1-400
